TQ_KERNEL=mma timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
TQ_KERNEL=mma python gpurun_smoke.py 2>&1 | tail -8
TQ_KERNEL=ws python gpurun_smoke.py 2>&1 | tail -6
