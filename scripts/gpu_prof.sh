set -x
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-small"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 2 -o gpurun_out/prof3 -f $CMD > gpurun_out/ncu3.log 2>&1
echo ncu rc=$?
tail -3 gpurun_out/plain.log | cut -c1-300
