python -m pytest tests -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
python __graft_entry__.py --smoke 2>&1 | tail -3
python scripts/small_grids.py --cases 1440x720d,360x180 --variants march,resident --steps 1000 2>/dev/null | cut -c1-120
