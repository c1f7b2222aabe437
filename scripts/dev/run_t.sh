python -m pytest tests -x -q -m gpu > gpurun_out/pytest_t.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_t.log
tail -3 gpurun_out/pytest_t.log
for g in "16384x8192" "7200x3600" "7200x3600 --diffusion"; do
python bench.py --grid $g --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-small 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$g', round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3))"
done
python scripts/small_grids.py --cases 360x180,1440x720,1440x720d --variants resident --steps 1000 2>/dev/null | cut -c1-90
