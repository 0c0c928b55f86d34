TQ_KERNEL=mma timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -c 2500 gpurun_out/bench_r1g.json
python profiles/time_resolve.py C2 both 2>&1 | tail -2
python profiles/time_resolve.py C3 both 2>&1 | tail -2
