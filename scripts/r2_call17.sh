set -x
timeout 900 python scripts/ab_kernels.py --grids 16384x8192 --steps 100 --envs "SWB_CHUNK=64;SWB_CHUNK=96;SWB_CHUNK=80;SWB_CHUNK=64;SWB_CHUNK=96 SWB_TAIL_ROWS=256;SWB_CHUNK=48" 2>&1 | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 7200x3600 --steps 300 --envs "SWB_CHUNK=64;SWB_CHUNK=96;SWB_CHUNK=80;SWB_CHUNK=64;SWB_CHUNK=96 SWB_TAIL_ROWS=640;SWB_CHUNK=48" 2>&1 | cut -c1-260
