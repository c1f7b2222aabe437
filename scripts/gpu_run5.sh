python -m pytest tests -m gpu -x -q > gpurun_out/pytest5.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest5.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo bench rc=$?; tail -3 gpurun_out/bench5.err; cat gpurun_out/bench5.json | cut -c1-330
