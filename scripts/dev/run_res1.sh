set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident or large_grids" > gpurun_out/pytest_res1.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_res1.log
tail -30 gpurun_out/pytest_res1.log
python scripts/small_grids.py --cases 360x180,1440x720,1440x720d --variants march,resident,resident-ldg > gpurun_out/small1.log 2>&1
cat gpurun_out/small1.log
