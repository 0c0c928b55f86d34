timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python profiles/prof_resolve.py 230300 2>&1 | grep -E "resolve ms" | tail -1
python gpurun_smoke.py 2>&1 | tail -4
