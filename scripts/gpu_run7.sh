G=${1:-8}
export SWB_BENCH_VERBOSE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $G --steps 100 --warmup 5 --no-e2e > gpurun_out/bench7_${G}gpu.json 2> gpurun_out/bench7_${G}gpu.err; echo bench rc=$?; grep "\[bench\]" gpurun_out/bench7_${G}gpu.err; cat gpurun_out/bench7_${G}gpu.json | cut -c1-300
