python -m pytest tests -m gpu -x -q > gpurun_out/pytest10.log 2>&1; echo pytest rc=$?; tail -25 gpurun_out/pytest10.log
python bench.py --dtype f32 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks'], d['e2e'])"
