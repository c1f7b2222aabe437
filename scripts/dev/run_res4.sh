python -m pytest tests -q -m gpu > gpurun_out/pytest_res4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_res4.log
tail -4 gpurun_out/pytest_res4.log
python scripts/small_grids.py --cases 360x180,1440x720,1440x720d,720x360d --variants march,resident --steps 1000 > gpurun_out/small4.log 2> gpurun_out/small4.err
cut -c1-150 gpurun_out/small4.log
