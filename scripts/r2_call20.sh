set -x
mkdir -p gpurun_out
bash scripts/r2_final_multi.sh 2 2>&1 | grep -v "^+" | cut -c1-400
for mult in 4 1 4 1; do
  SWB_RES_MULT=$mult timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --strong --grid 2880x1440 --steps 1000 --warmup 50 --no-e2e --no-parity --no-strong --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mult $mult', d['roofline']['kernel'], round(d['ms_per_step']*1e3,2), 'us/step')"
done
for mult in 4 1; do
  SWB_RES_MULT=$mult timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --strong --grid 1440x720 --diffusion --steps 2000 --warmup 50 --no-e2e --no-parity --no-strong --no-sustained 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1440x720+diff mult $mult', d['roofline']['kernel'], round(d['ms_per_step']*1e3,2), 'us/step')"
done
