# the CFL estimate in fp32: exactness tests, then the step time against the fp64 estimate of the previous build
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "cfl or headline or bench_horizon or nan or past_final or streaming" > gpurun_out/r3_pytest_cfl.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/r3_pytest_cfl.log | cut -c1-200
for rep in 1 2; do
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-small --no-fp32 > gpurun_out/r3_cfl_bench$rep.json 2> gpurun_out/r3_cfl_bench$rep.err; echo rc=$?
python - $rep <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r3_cfl_bench{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), "sustained", round(d["roofline"]["sustained"]["ms_per_step"], 4), round(d["roofline"]["sustained"]["frac"], 4), d["roofline"]["sustained"]["clocks"]["sm_mhz"])
PY
done
python scripts/ab_kernels.py --grids 7200x3600 --envs "SWB_RESIDENT=0" --steps 200 2>&1 | cut -c1-250
