# dev probe: the step kernel's memory traffic without its arithmetic (library built with -DSWB_COPY_ONLY: every row stores the row above)
mkdir -p gpurun_out
for dt in f64 f32; do for lib in libswb200 libswb200_copy; do
SWB_LIB=sw_solver_b200/$lib.so python bench.py --dtype $dt --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained > gpurun_out/r3_copy_${dt}_$lib.json 2> gpurun_out/r3_copy_${dt}_$lib.err; echo "$dt $lib rc=$?"
python - "$dt" "$lib" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r3_copy_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1:], "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[1:], "failed", e); print(open(f"gpurun_out/r3_copy_{sys.argv[1]}_{sys.argv[2]}.err").read()[-600:])
PY
done; done
