# A/B on one box: CFL estimate in fp32 (default library) against the fp64 estimate (-DSWB_CFL_EST_F32=0), interleaved
mkdir -p gpurun_out
for rep in 1 2 3; do for lib in libswb200 libswb200_est64; do
SWB_LIB=sw_solver_b200/$lib.so python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained > gpurun_out/r3_ab.json 2> gpurun_out/r3_ab.err
python - $lib <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r3_ab.json").read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4))
PY
done; done
for lib in libswb200 libswb200_est64; do SWB_LIB=sw_solver_b200/$lib.so python scripts/ab_kernels.py --grids 7200x3600,16384x8192 --envs "SWB_RESIDENT=0" --steps 100 --sustain 2 2>&1 | cut -c1-330; done
