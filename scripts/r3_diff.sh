# two-phase diffusion tiles: parity (ring-fed star against the direct-load kernels, resident kernels, bands), then step times
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bands.py -x -q -m gpu -k "diffusion or resident or star or streaming or bands or multi_step or oracle" > gpurun_out/r3_pytest_diff.log 2>&1; echo pytest rc=$?
tail -6 gpurun_out/r3_pytest_diff.log | cut -c1-300
python scripts/ab_kernels.py --grids 7200x3600,4000x2000,16384x8192 --diffusion --envs "SWB_RESIDENT=0" --steps 100 > gpurun_out/r3_ab_diff.log 2>&1; cat gpurun_out/r3_ab_diff.log | cut -c1-300
python scripts/small_grids.py --steps 2000 --cases 1440x720d,2000x1000d,1000x500d > gpurun_out/r3_small_diff.log 2>&1; grep -v "^$" gpurun_out/r3_small_diff.log | cut -c1-400 | tail -8
