python -m pytest tests -q -x -m gpu > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
python scripts/small_grids.py --cases 1440x720d,1440x720,360x180 --variants march,resident --steps 1000 2>/dev/null | cut -c1-120
python bench.py --grid 7200x3600 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-small 2>&1 | cut -c1-330
python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['l2_resident_configs'])"
