set -x
mkdir -p gpurun_out
timeout 900 python scripts/ab_kernels.py --grids 7200x3600 --steps 300 --envs "SWB_TAIL_ROWS=0;SWB_PDL=1;SWB_TAIL_ROWS=960;SWB_TAIL_ROWS=1280;SWB_TAIL_ROWS=960 SWB_TAIL_CHUNK=24" 2>&1 | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 16384x8192 --steps 100 --sustain 1.5 --envs "SWB_TAIL_ROWS=0;SWB_PDL=1;SWB_TAIL_ROWS=512" 2>&1 | cut -c1-330
timeout 900 python scripts/ab_kernels.py --grids 7200x3600 --diffusion --steps 200 --envs "SWB_TAIL_ROWS=0;SWB_PDL=1" 2>&1 | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 16384x8192 --dtype float32 --steps 100 --envs "SWB_TAIL_ROWS=0;SWB_PDL=1" 2>&1 | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 4000x2000,2880x1440 --steps 300 --envs "SWB_RESIDENT=0 SWB_TAIL_ROWS=0;SWB_RESIDENT=0;SWB_RESIDENT=1" 2>&1 | cut -c1-260
