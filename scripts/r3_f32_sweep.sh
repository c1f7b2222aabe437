# fp32 step kernel on the headline grid: marching length and guided-tail sweep (the defaults were tuned on the fp64 kernel, two blocks per SM)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --dtype f32 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained"
for c in 48 64 96; do for t in 256 512 1024; do for tc in 16 32; do
SWB_CHUNK=$c SWB_TAIL_ROWS=$t SWB_TAIL_CHUNK=$tc $B > gpurun_out/r3_f32s.json 2> gpurun_out/r3_f32s.err
python - "$c" "$t" "$tc" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r3_f32s.json").read().strip().splitlines()[-1])
    print("chunk", sys.argv[1], "tail rows", sys.argv[2], "tail chunk", sys.argv[3], "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[1:], "failed", e)
PY
done; done; done
