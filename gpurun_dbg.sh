timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err || { tail -5 gpurun_out/bench_r1i.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_r1i_plain.json 2> gpurun_out/bench_r1i_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_r1i_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resolve_mma_kernel -s 4 -c 1 -o gpurun_out/prof_r1i_bench -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_r1i_bench.log 2>&1
python bench.py --config C3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1i_C3.json 2> gpurun_out/bench_r1i_C3.err
python -c "import json;[print(f, json.load(open('gpurun_out/'+f))['value'], json.load(open('gpurun_out/'+f))['e2e']['value']) for f in ['bench_r1i.json','bench_r1i_plain.json','bench_r1i_C3.json']]"
