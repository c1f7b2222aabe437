CMD="python bench.py --grid 7200x3600 --diffusion --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-small"
$CMD > gpurun_out/plain_diff.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 1 -o gpurun_out/prof_diff -f $CMD > gpurun_out/ncu_diff.log 2>&1
echo ncu rc=$?
