# two independent single-GPU runs at the same time (same box, same power / thermal conditions as a 2-band run)
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-small > gpurun_out/solo0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-small > gpurun_out/solo1.json 2>/dev/null &
wait
python -c "
import json
for f in ('solo0','solo1'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['ms_per_step'],4), d['clocks'])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bands x2', round(d['ms_per_step'],4), d['clocks'])"
