timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python profiles/time_resolve.py C2 both 2>&1 | tail -2
