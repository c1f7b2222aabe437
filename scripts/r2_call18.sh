set -x
for lib in libswb200 libswb200_EARLY libswb200_PRE libswb200_BOTH libswb200; do
  echo "== $lib"
  SWB_LIB=$PWD/sw_solver_b200/$lib.so timeout 600 python scripts/ab_kernels.py --grids 16384x8192,7200x3600 --steps 200 --envs "SWB_PDL=1" 2>&1 | cut -c1-260
  SWB_LIB=$PWD/sw_solver_b200/$lib.so timeout 600 python scripts/ab_kernels.py --grids 16384x8192 --steps 200 --dtype float32 --envs "SWB_PDL=1" 2>&1 | cut -c1-260
  SWB_LIB=$PWD/sw_solver_b200/$lib.so timeout 600 python scripts/ab_kernels.py --grids 7200x3600 --diffusion --steps 200 --envs "SWB_PDL=1" 2>&1 | cut -c1-260
done
SWB_LIB=$PWD/sw_solver_b200/libswb200_BOTH.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bands.py -x -q -k "against_oracle or bands_are or cfl_filter_is_exact or guided or steps_past or nan_state or fp32_mode_against or known_answer" > gpurun_out/r2_pytest18.log 2>&1; echo pytest-both rc=$?; tail -3 gpurun_out/r2_pytest18.log
