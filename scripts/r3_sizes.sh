# how the step kernel's fraction of HBM depends on the length of the launch (same 16384 longitudes, more or fewer latitudes)
mkdir -p gpurun_out
for g in 16384x2048 16384x8192 16384x32768; do
python bench.py --grid $g --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained > gpurun_out/r3_size_$g.json 2> gpurun_out/r3_size_$g.err; echo "$g rc=$?"
python - "$g" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r3_size_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), d["clocks"])
PY
done
