set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bands.py -x -q -k "guided or bands_are or against_oracle or cfl_filter_is_exact or contexts_leave" > gpurun_out/r2_pytest15.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r2_pytest15.log | cut -c1-300
timeout 900 python scripts/ab_kernels.py --grids 7200x3600 --steps 300 --envs "SWB_TAIL2_ROWS=0;SWB_PDL=1;SWB_TAIL2_ROWS=80;SWB_TAIL2_ROWS=160;SWB_TAIL_ROWS=0" 2>&1 | cut -c1-260
timeout 900 python scripts/ab_kernels.py --grids 16384x8192 --steps 100 --envs "SWB_TAIL2_ROWS=0;SWB_PDL=1;SWB_TAIL2_ROWS=48;SWB_TAIL_ROWS=0" 2>&1 | cut -c1-260
