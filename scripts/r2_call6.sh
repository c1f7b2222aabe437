set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "streaming or large_grids or cfl_filter" > gpurun_out/r2_pytest6.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/r2_pytest6.log | cut -c1-300
export SWB_VERBOSE=1
timeout 900 python scripts/ab_kernels.py --grids 7200x3600,16384x8192 --sustain 1.5 --envs "SWB_RESIDENT=0;SWB_STREAM=1 SWB_CFL_FILTER=0;SWB_STREAM=1 SWB_CFL_FILTER=1;SWB_STREAM=1 SWB_CFL_FILTER=1 SWB_STREAM_UNIT=64;SWB_STREAM=1 SWB_STREAM_MAP=0" > gpurun_out/r2_ab6.log 2>&1; cat gpurun_out/r2_ab6.log | cut -c1-420
timeout 600 python scripts/ab_kernels.py --grids 7200x3600,2880x1440 --diffusion --envs "SWB_STREAM=0;SWB_STREAM=1" > gpurun_out/r2_ab6d.log 2>&1; cat gpurun_out/r2_ab6d.log | cut -c1-420
timeout 600 python scripts/ab_kernels.py --grids 2880x1440,4000x2000,1440x720 --envs "SWB_STREAM=0;SWB_STREAM=1" > gpurun_out/r2_ab6s.log 2>&1; cat gpurun_out/r2_ab6s.log | cut -c1-420
