set -x
mkdir -p gpurun_out
for lib in libswb200 libswb200_u2; do
  echo "== $lib"
  SWB_LIB=$PWD/sw_solver_b200/$lib.so timeout 600 python scripts/ab_kernels.py --grids 16384x8192,7200x3600 --steps 200 --envs "SWB_PDL=1" 2>&1 | cut -c1-260
  SWB_LIB=$PWD/sw_solver_b200/$lib.so timeout 600 python scripts/ab_kernels.py --grids 16384x8192 --steps 200 --dtype float32 --envs "SWB_PDL=1" 2>&1 | cut -c1-260
done
