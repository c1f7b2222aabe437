# strong scaling of BASELINE configs 2 and 3 over G GPUs of one box (latitude bands)
G=${1:-2}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $G --strong --no-e2e --no-cpu-baseline --steps 400 --warmup 20 "$@" 2>gpurun_out/strong.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], '|', d['roofline']['kernel'], '|', round(d['ms_per_step']*1e3,2), 'us/step |', round(d['value']/1e9,1), 'G updates/s')"; }
run --grid 7200x3600
run --grid 1440x720 --diffusion
