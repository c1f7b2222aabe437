# resident kernel: tiles that hold a pole / halo row run only THAT row through the boundary body -- bit-identity tests, then small-grid step times
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bands.py -x -q -m gpu -k "resident or streaming or bands or small or diffusion_star or separable" > gpurun_out/r3_pytest_res.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/r3_pytest_res.log | cut -c1-200
python scripts/small_grids.py --steps 2000 --cases 360x180,720x360,1440x720,1440x720d,2000x1000,2000x1000d,2880x1440 > gpurun_out/r3_small_res.log 2>&1; grep resident\"\:\ true gpurun_out/r3_small_res.log | cut -c1-260
