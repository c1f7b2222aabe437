# round 2: what the driver runs at round end on one GPU (tests, smoke, bench both arms) + the round-2 ncu evidence
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_final.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/r2_pytest_final.log | cut -c1-200
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r2_smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_ref_final.json 2> gpurun_out/r2_ref_final.err; echo ref rc=$?
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?
cut -c1-1200 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained"
$CMD > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1; echo ncu1 rc=$?
$CMD > gpurun_out/r2_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 2 -o gpurun_out/r2_prof_headline -f $CMD > gpurun_out/r2_ncu_headline.log 2>&1; echo ncu2 rc=$?
ls -la gpurun_out/launches_r2.csv gpurun_out/r2_prof_headline.ncu-rep
