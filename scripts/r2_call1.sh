# round 2, GPU call 1 (one GPU): full GPU test suite, bench-horizon error growth, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r2_pytest1.log
timeout 600 python scripts/horizon_parity.py --out gpurun_out/r2_horizon.json > gpurun_out/r2_horizon.log 2>&1; echo horizon rc=$?
tail -40 gpurun_out/r2_horizon.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo bench rc=$?
cut -c1-2500 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err; echo ref rc=$?
cut -c1-1500 gpurun_out/r2_ref1.json
