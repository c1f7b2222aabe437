// dev probe: one step kernel alone (a second to compile; SASS counts with scripts/sass_hist.py)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I include -I sw_solver_b200/csrc -cubin
//        [-DPROBE_F64] [-DPROBE_MINB=3] [-DPROBE_DIFF=true] -o x.cubin scripts/probes/f32_inst.cu
#include "swb_kernels.cuh"
#ifndef PROBE_MINB
#define PROBE_MINB 4
#endif
#ifndef PROBE_DIFF
#define PROBE_DIFF false
#endif
#ifndef PROBE_R
#define PROBE_R 4
#endif
#ifndef PROBE_S
#define PROBE_S 3
#endif
#ifdef PROBE_F64
// the headline fp64 kernel: march_kernel<double, 0, 0, 0, 0, 1, 4, 4, 0, 2, 4, 1> (or with diffusion: S = 3, no CFL filter)
template __global__ void swb::march_kernel<double, false, PROBE_DIFF, 0, false, true, 4, (PROBE_DIFF ? 3 : 4), false, 2, 4, !PROBE_DIFF>(
    const __grid_constant__ swb::Params, const __grid_constant__ swb::TmaMaps);
#else
template __global__ void swb::march_kernel<float, false, PROBE_DIFF, 0, false, true, PROBE_R, PROBE_S, false, PROBE_MINB, 4, false>(
    const __grid_constant__ swb::Params, const __grid_constant__ swb::TmaMaps);
#endif
