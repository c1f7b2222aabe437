set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "streaming or resident or large_grids or diffusion_star" > gpurun_out/r2_pytest3.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/r2_pytest3.log
timeout 900 python scripts/ab_kernels.py --grids 7200x3600,16384x8192 --envs "SWB_RESIDENT=0;SWB_STREAM=1" --sustain 1.5 > gpurun_out/r2_ab3.log 2>&1; cat gpurun_out/r2_ab3.log
timeout 600 python scripts/ab_kernels.py --grids 7200x3600,2880x1440,1440x720 --diffusion --envs "SWB_STREAM=0;SWB_STREAM=1" > gpurun_out/r2_ab3d.log 2>&1; cat gpurun_out/r2_ab3d.log
timeout 600 python scripts/ab_kernels.py --grids 2880x1440,4000x2000,1440x720 --envs "SWB_STREAM=0;SWB_STREAM=1" > gpurun_out/r2_ab3s.log 2>&1; cat gpurun_out/r2_ab3s.log
