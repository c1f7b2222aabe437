# round 2, session 2: the default library (scalar fp32 rows) and the packed-fp32 A/B library, parity then step times; e2e leg with the prefetched first field
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fp32 or pinned or guards_select" > gpurun_out/r3_pytest_f32.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/r3_pytest_f32.log | cut -c1-300
SWB_LIB=sw_solver_b200/libswb200_pk.so SWB_TEST_F32_PACKED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fp32 or guards_select" > gpurun_out/r3_pytest_f32_pk.log 2>&1; echo pytest-pk rc=$?
tail -5 gpurun_out/r3_pytest_f32_pk.log | cut -c1-300
B="python bench.py --steps 20 --warmup 5 --dtype f32 --no-e2e --no-cpu-baseline --no-small --no-fp32 --no-sustained"
for lib in libswb200 libswb200_pk; do for v in 2 1; do SWB_LIB=sw_solver_b200/$lib.so SWB_TMA_VARIANT_F32=$v timeout 300 $B > gpurun_out/r3_f32_${lib}_v$v.json 2> gpurun_out/r3_f32_${lib}_v$v.err; echo "$lib variant $v rc=$?"; cut -c1-230 gpurun_out/r3_f32_${lib}_v$v.json; done; done
