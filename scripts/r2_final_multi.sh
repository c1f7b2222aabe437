# round 2, multi-GPU validation with the final library: band / solve() parity on real IPC peers, the bench line as the
# driver runs it (band parity in front of the timed region, strong-scaling sub-record), and a 100-step weak-scaling point
G=${1:-8}
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29551 scripts/check_bands_multi.py > gpurun_out/r2_bands_multi_$G.log 2>&1; echo check rc=$?
grep -E "bands x|solve\(\) x|Error|error|Traceback" gpurun_out/r2_bands_multi_$G.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/r2_final_${G}gpu.json 2> gpurun_out/r2_final_${G}gpu.err; echo bench rc=$?
cut -c1-600 gpurun_out/r2_final_${G}gpu.json; tail -3 gpurun_out/r2_final_${G}gpu.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $G --steps 100 --warmup 10 --no-e2e --no-parity --no-strong > gpurun_out/r2_final_${G}gpu_100.json 2> gpurun_out/r2_final_${G}gpu_100.err; echo bench100 rc=$?
cut -c1-400 gpurun_out/r2_final_${G}gpu_100.json
if [ "$G" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "contexts_leave" 2>&1 | tail -3; fi
