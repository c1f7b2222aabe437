# ncu of the fp32 step kernel on the headline grid (round 2)
set -x
CMD="python bench.py --steps 5 --warmup 3 --dtype f32 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained"
$CMD > gpurun_out/r2_plain_f32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 1 -o gpurun_out/r2_prof_f32 -f $CMD > gpurun_out/r2_ncu_f32.log 2>&1; echo ncu rc=$?
tail -2 gpurun_out/r2_plain_f32.log | cut -c1-300
