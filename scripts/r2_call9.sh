set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest9.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/r2_pytest9.log | cut -c1-300
timeout 600 python scripts/ab_kernels.py --grids 1440x720,7200x3600 --steps 300 --envs "SWB_PDL=1" > gpurun_out/r2_ab9rh.log 2>&1; cat gpurun_out/r2_ab9rh.log | cut -c1-300
timeout 600 python scripts/ab_kernels.py --ic zgf --grids 1440x720,7200x3600 --steps 300 --envs "SWB_FSEP=1;SWB_FSEP=0" > gpurun_out/r2_ab9zgf.log 2>&1; cat gpurun_out/r2_ab9zgf.log | cut -c1-300
timeout 600 python scripts/ab_kernels.py --grids 7200x3600,2880x1440,1440x720 --diffusion --steps 300 --envs "SWB_PDL=1" > gpurun_out/r2_ab9d.log 2>&1; cat gpurun_out/r2_ab9d.log | cut -c1-300
timeout 600 python scripts/ab_kernels.py --ic zgf --grids 1440x720 --diffusion --steps 300 --envs "SWB_FSEP=1;SWB_FSEP=0" > gpurun_out/r2_ab9zd.log 2>&1; cat gpurun_out/r2_ab9zd.log | cut -c1-300
