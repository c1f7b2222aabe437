python -m pytest tests -x -q -m gpu > gpurun_out/pytest_diff1.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_diff1.log
tail -5 gpurun_out/pytest_diff1.log
python scripts/small_grids.py --cases 1440x720d,720x360d,360x180d --variants march,resident,resident-tma --steps 1000 > gpurun_out/small5.log 2> gpurun_out/small5.err
cut -c1-170 gpurun_out/small5.log; tail -3 gpurun_out/small5.err
