python -m pytest tests -m gpu -x -q > gpurun_out/pytest4.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest4.log
for v in 0 1 2 3 4 5; do for ch in 64 128; do echo "variant $v chunk $ch"; SWB_CHUNK=$ch SWB_TMA_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['clocks'])"; done; done
SWB_KERNEL=ldg timeout 300 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
