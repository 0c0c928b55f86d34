export TQ_KERNEL=ws
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "micro4 or tiny8" 2>&1 | tail -3 || echo "TIMEOUT small"
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 || echo "TIMEOUT suite"
timeout 60 python profiles/prof_resolve.py 230300 2>&1 | grep -E "resolve ms" | tail -1
