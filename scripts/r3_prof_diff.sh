# ncu --set full of the step kernel WITH diffusion on 7200 x 3600 (round 2 library): report + raw page
set -x
mkdir -p gpurun_out
CMD="python bench.py --grid 7200x3600 --diffusion --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained"
$CMD > gpurun_out/r3_plain_diff.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 1 -o gpurun_out/r3_prof_diff -f $CMD > gpurun_out/r3_ncu_diff.log 2>&1; echo ncu rc=$?
ncu -i gpurun_out/r3_prof_diff.ncu-rep --page raw --csv > gpurun_out/r3_prof_diff_raw.csv 2>/dev/null
tail -1 gpurun_out/r3_plain_diff.log | cut -c1-300
