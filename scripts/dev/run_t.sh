bash scripts/gpu_multi.sh 8
bash scripts/gpu_strong.sh 8
