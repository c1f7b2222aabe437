set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest8.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/r2_pytest8.log | cut -c1-300
timeout 600 python scripts/ab_kernels.py --grids 7200x3600,2880x1440,1440x720 --diffusion --steps 300 --envs "SWB_PDL=1" > gpurun_out/r2_ab8d.log 2>&1; cat gpurun_out/r2_ab8d.log | cut -c1-300
