# ncu capture of resident_kernel on BASELINE config 2 (1440 x 720, diffusion on)
CMD="python scripts/small_grids.py --cases 1440x720d --variants resident --steps 200 --warmup 50"
$CMD > gpurun_out/plain_small_diff.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resident_kernel -s 1 -c 1 -o gpurun_out/prof_res_diff -f $CMD > gpurun_out/ncu_res_diff.log 2>&1
echo ncu rc=$?
