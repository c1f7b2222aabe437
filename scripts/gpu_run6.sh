export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench6_2gpu.json 2> gpurun_out/bench6_2gpu.err; echo rc=$?; tail -20 gpurun_out/bench6_2gpu.err; cat gpurun_out/bench6_2gpu.json | cut -c1-1500
