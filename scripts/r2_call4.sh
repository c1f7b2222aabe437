set -x
mkdir -p gpurun_out
export SWB_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "streaming or resident or large_grids" > gpurun_out/r2_pytest4.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/r2_pytest4.log | cut -c1-300
timeout 900 python scripts/ab_kernels.py --grids 7200x3600,16384x8192 --envs "SWB_RESIDENT=0;SWB_STREAM=1 SWB_STREAM_MAP=1;SWB_STREAM=1 SWB_STREAM_MAP=1 SWB_STREAM_UNIT=64;SWB_STREAM=1 SWB_STREAM_MAP=0" > gpurun_out/r2_ab4.log 2>&1; cat gpurun_out/r2_ab4.log | cut -c1-400
timeout 600 python scripts/ab_kernels.py --grids 1440x720 --envs "SWB_STREAM=0;SWB_STREAM=1 SWB_STREAM_MAP=1;SWB_STREAM=1 SWB_STREAM_MAP=1 SWB_STREAM_UNIT=16;SWB_STREAM=1 SWB_STREAM_MAP=0" > gpurun_out/r2_ab4s.log 2>&1; cat gpurun_out/r2_ab4s.log | cut -c1-400
grep -h "swb\]" gpurun_out/r2_ab4*.log | sort | uniq | head
