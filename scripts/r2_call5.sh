set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "streaming or large_grids" > gpurun_out/r2_pytest5.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/r2_pytest5.log | cut -c1-300
CH="python scripts/ab_kernels.py --grid 7200x3600 --steps 3 --warmup 2"
SWB_STREAM=1 SWB_STREAM_UNIT=64 $CH > gpurun_out/r2_plain5.log 2>&1 && \
SWB_STREAM=1 SWB_STREAM_UNIT=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident_kernel -c 2 -o gpurun_out/r2_prof_stream -f $CH > gpurun_out/r2_ncu5a.log 2>&1; echo ncu rc=$?
SWB_RESIDENT=0 SWB_CFL_FILTER=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 2 -o gpurun_out/r2_prof_march -f $CH > gpurun_out/r2_ncu5b.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/*.ncu-rep
