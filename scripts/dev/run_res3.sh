python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident or cfl_filter" > gpurun_out/pytest_res3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_res3.log
tail -3 gpurun_out/pytest_res3.log
for c in 360x180 1440x720 1440x720d; do
SWB_RES_TIMING_FILE=gpurun_out/stamps_$c.bin python scripts/small_grids.py --cases $c --variants resident --steps 500 >> gpurun_out/small3.log 2>> gpurun_out/small3.err
done
cut -c1-120 gpurun_out/small3.log; grep res-timing gpurun_out/small3.err | awk '{$1="";print}' | sort -u -k5,6 | head
