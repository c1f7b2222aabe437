set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "resident or diffusion_star or guided or separable" > gpurun_out/r2_pytest16.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r2_pytest16.log | cut -c1-300
export SWB_VERBOSE=1
timeout 900 python scripts/ab_kernels.py --grids 1440x720,2000x1000,2880x1440,4000x2000 --steps 2000 --envs "SWB_RES_MULT=4;SWB_RES_MULT=1" 2>&1 | grep -v "^\[swb" | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 1440x720,2000x1000,2880x1440 --diffusion --steps 2000 --envs "SWB_RES_MULT=4;SWB_RES_MULT=1" 2>&1 | grep -v "^\[swb" | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 360x180,720x360 --steps 4000 --envs "SWB_RES_MULT=4;SWB_RES_MULT=1" 2>&1 | grep -v "^\[swb" | cut -c1-260
