python scripts/small_grids.py --cases 360x180,1440x720,1440x720d,720x360,2880x1440 --variants march,resident --steps 1000 > gpurun_out/small4.log 2> gpurun_out/small4.err
cut -c1-150 gpurun_out/small4.log
