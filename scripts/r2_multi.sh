# round 2, multi-GPU call: (1) the launch-per-step band path with globaltimer stamps (-DSWB_BAND_TIMING build),
# (2) the bench line as the driver runs it (band parity in front of the timed region, strong-scaling sub-record)
G=${1:-2}
set -x
mkdir -p gpurun_out
[ -f sw_solver_b200/libswb200_bt.so ] || SWB_NVCC_FLAGS="-DSWB_BAND_TIMING" python -m sw_solver_b200.build --out=$PWD/sw_solver_b200/libswb200_bt.so
SWB_LIB=$PWD/sw_solver_b200/libswb200_bt.so SWB_BAND_TIMING_FILE=$PWD/gpurun_out/bt$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --steps 200 --warmup 10 --no-e2e --no-parity --no-strong --no-sustained > gpurun_out/r2_bt_${G}gpu.json 2> gpurun_out/r2_bt_${G}gpu.err; echo bt rc=$?
cut -c1-400 gpurun_out/r2_bt_${G}gpu.json; tail -3 gpurun_out/r2_bt_${G}gpu.err
ls -la gpurun_out/bt$G* | head -20
python scripts/band_timing.py gpurun_out/bt$G.16384x$((8192*G)) --first 12 --last 25; python scripts/band_timing.py gpurun_out/bt$G.16384x$((8192*G)) --last 60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/r2_bench_${G}gpu.json 2> gpurun_out/r2_bench_${G}gpu.err; echo bench rc=$?
cut -c1-3000 gpurun_out/r2_bench_${G}gpu.json; tail -5 gpurun_out/r2_bench_${G}gpu.err
if [ "$G" = "2" ]; then python scripts/probe_copy.py > gpurun_out/r2_probe_copy.log 2>&1; cat gpurun_out/r2_probe_copy.log | head -60; fi
