set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest7.log 2>&1; echo pytest rc=$?
tail -8 gpurun_out/r2_pytest7.log | cut -c1-300
timeout 900 python scripts/ab_kernels.py --grids 16384x8192,7200x3600 --steps 200 --envs "SWB_PDL=0;SWB_PDL=1;SWB_PDL=1" > gpurun_out/r2_ab7.log 2>&1; cat gpurun_out/r2_ab7.log | cut -c1-300
timeout 900 python scripts/ab_kernels.py --grids 7200x3600 --steps 200 --envs "SWB_CHUNK=56;SWB_CHUNK=64;SWB_CHUNK=72;SWB_CHUNK=76;SWB_CHUNK=80;SWB_CHUNK=96;SWB_CHUNK=100" > gpurun_out/r2_ab7c.log 2>&1; cat gpurun_out/r2_ab7c.log | cut -c1-300
timeout 900 python scripts/ab_kernels.py --grids 16384x8192 --steps 100 --envs "SWB_CHUNK=64;SWB_CHUNK=72;SWB_CHUNK=96" > gpurun_out/r2_ab7d.log 2>&1; cat gpurun_out/r2_ab7d.log | cut -c1-300
