# L2-resident configurations: phase breakdown (timing build), step times, ncu capture of resident_kernel
set -x
for c in 360x180 1440x720 1440x720d; do
SWB_LIB=sw_solver_b200/libswb200_timing.so SWB_RES_TIMING_FILE=gpurun_out/stamps_$c.bin python scripts/small_grids.py --cases $c --variants resident --steps 500 >> gpurun_out/small_timing.log 2>> gpurun_out/small_timing.err
done
grep res-timing gpurun_out/small_timing.err | awk '{$1="";print}' | sort -u -k5,6
python scripts/small_grids.py --cases 360x180,720x360,1440x720,1440x720d,2000x1000 --variants march,resident --steps 2000 > gpurun_out/small_final.log 2> gpurun_out/small_final.err
cut -c1-170 gpurun_out/small_final.log
CMD="python scripts/small_grids.py --cases 1440x720 --variants resident --steps 200 --warmup 50"
$CMD > gpurun_out/plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resident_kernel -s 1 -c 1 -o gpurun_out/prof_res -f $CMD > gpurun_out/ncu_res.log 2>&1
echo ncu rc=$?
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json | cut -c1-3000
