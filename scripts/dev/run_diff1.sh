python -m pytest tests -x -q -m gpu > gpurun_out/pytest_diff1.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_diff1.log
tail -15 gpurun_out/pytest_diff1.log
python scripts/small_grids.py --cases 1440x720d,720x360d --variants march,resident,resident-tma,resident-ldg --steps 1000 > gpurun_out/small5.log 2> gpurun_out/small5.err
cut -c1-170 gpurun_out/small5.log; tail -3 gpurun_out/small5.err
python bench.py --grid 7200x3600 --diffusion --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-small 2>&1 | cut -c1-330
python bench.py --grid 7200x3600 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-small 2>&1 | cut -c1-330
