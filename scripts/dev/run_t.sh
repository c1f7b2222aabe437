for g in 7200x3600 7200x902 7200x452; do
for ch in 64 48 52 56 60 68 72 76 80 88 96 100; do
SWB_CHUNK=$ch python bench.py --grid $g --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-small 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$g chunk $ch', round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3))"
done; done
