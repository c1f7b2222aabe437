set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "save_data or known_answer or print_interval or driver" > gpurun_out/r2_pytest11.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r2_pytest11.log | cut -c1-300
for lib in libswb200 libswb200_sp; do
  echo "== $lib"
  SWB_LIB=$PWD/sw_solver_b200/$lib.so timeout 600 python scripts/ab_kernels.py --grids 7200x3600,2880x1440,1440x720 --diffusion --steps 300 --envs "SWB_PDL=1;SWB_RESIDENT=0" 2>&1 | cut -c1-260
done
SWB_LIB=$PWD/sw_solver_b200/libswb200_sp.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "diffusion_star or resident_kernel_is_bit or against_oracle" > gpurun_out/r2_pytest11b.log 2>&1; echo pytest-sp rc=$?; tail -3 gpurun_out/r2_pytest11b.log
python - <<'PY'
import time, sys
sys.path.insert(0, '.')
from sw_solver_b200.solver import solve
from sw_solver_b200.ic import ICType
solve(360, 180, ICType.RossbyHaurwitzWave, 0.01, 0.5, False)
for kw in ({}, {"save_data": {"interval": 10}}, {"save_data": {"interval": 100}}):
    t0 = time.perf_counter(); solve(360, 180, ICType.RossbyHaurwitzWave, 1.0, 0.5, False, **kw); print(kw.get("save_data", {}).get("interval"), "wall %.3f s" % (time.perf_counter() - t0))
PY
