# round 2, second session: tests and the bench line with its new sub-records on one GPU
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r3_pytest_final.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/r3_pytest_final.log | cut -c1-200
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3_bench_final.json 2> gpurun_out/r3_bench_final.err; echo bench rc=$?
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3_bench_final.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"] / 1e9)
print(d["config3_7200x3600_one_gpu"]); print(d["l2_resident_configs"][2])
PY
python scripts/probe_solve_save.py 2>&1 | head -3
