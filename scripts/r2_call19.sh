set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest19.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/r2_pytest19.log | cut -c1-300
timeout 900 python scripts/ab_kernels.py --grids 16384x8192,7200x3600 --steps 200 --sustain 1.0 --envs "SWB_CHUNK=64;SWB_PDL=1;SWB_CHUNK=64;SWB_PDL=1" 2>&1 | cut -c1-330
timeout 900 python scripts/ab_kernels.py --grids 7200x3600 --diffusion --steps 200 --envs "SWB_PDL=1" 2>&1 | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 16384x8192 --dtype float32 --steps 200 --envs "SWB_PDL=1" 2>&1 | cut -c1-260
