python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident or large_grids" > gpurun_out/pytest_res2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_res2.log
tail -5 gpurun_out/pytest_res2.log
python scripts/small_grids.py --cases 360x180,1440x720,1440x720d --variants march,resident,resident-ldg --steps 500 > gpurun_out/small2.log 2> gpurun_out/small2.err
cut -c1-200 gpurun_out/small2.log; grep res-timing gpurun_out/small2.err | sort | uniq -c | sort -rn | awk '{$1="";print}' | sort -u -k5,6 | head -40
