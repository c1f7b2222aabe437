python -m pytest tests -x -q -m gpu > gpurun_out/pytest_t.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_t.log
tail -6 gpurun_out/pytest_t.log
