# multi-GPU checks: parity of the band path (bit-identical to one GPU) + scaling bench
G=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 scripts/check_bands_multi.py > gpurun_out/bands_multi_$G.log 2>&1; echo check rc=$?; grep -E "bands x|solve\(\) x|Error|error|Traceback" gpurun_out/bands_multi_$G.log | head -20
if [ "${2:-}" = "bench" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $G --steps 100 --warmup 5 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err; echo bench rc=$?; cat gpurun_out/bench_${G}gpu.json | cut -c1-330
fi
