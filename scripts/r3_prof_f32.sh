# ncu --set full of the packed fp32 step kernel: ring variant 1 (report kept), variant 2 (raw page only: two reports exceed the pull limit)
set -x
CMD="python bench.py --steps 5 --warmup 3 --dtype f32 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained"
for v in 1 2; do
export SWB_TMA_VARIANT_F32=$v
$CMD > gpurun_out/r3_plain_f32_v$v.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 1 -o gpurun_out/r3_prof_f32_v$v -f $CMD > gpurun_out/r3_ncu_f32_v$v.log 2>&1; echo ncu rc=$?
ncu -i gpurun_out/r3_prof_f32_v$v.ncu-rep --page raw --csv > gpurun_out/r3_prof_f32_v${v}_raw.csv 2>/dev/null
done
rm -f gpurun_out/r3_prof_f32_v2.ncu-rep
ls -la gpurun_out
