// Are Blackwell's packed fp32 instructions (FFMA2 / FADD2 / FMUL2) rounded element for element like
// the scalar ones?  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -o f32x2_probe f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ float lo(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ float hi(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ unsigned rnd(unsigned &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ float rf(unsigned &s, int mode)
{
    if (mode == 0) return __uint_as_float((rnd(s) & 0x007fffffu) | 0x3f800000u) * ((rnd(s) & 1) ? 1.f : -1.f);  // [1,2) with sign
    unsigned e = 100 + rnd(s) % 56;                                                                              // wide exponents
    return __uint_as_float((rnd(s) & 0x807fffffu) | (e << 23));
}
__global__ void k(unsigned long long *bad, int mode)
{
    unsigned s = 1234567u + blockIdx.x * 7919u + threadIdx.x * 104729u + mode;
    unsigned long long nb = 0;
    for (int it = 0; it < 20000; ++it) {
        float a0 = rf(s, mode), a1 = rf(s, mode), b0 = rf(s, mode), b1 = rf(s, mode), c0 = rf(s, mode), c1 = rf(s, mode);
        unsigned long long A = pk(a0, a1), B = pk(b0, b1), C = pk(c0, c1), r;
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(A), "l"(B), "l"(C));
        nb += (__float_as_uint(lo(r)) != __float_as_uint(__fmaf_rn(a0, b0, c0))) + (__float_as_uint(hi(r)) != __float_as_uint(__fmaf_rn(a1, b1, c1)));
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(A), "l"(B));
        nb += (__float_as_uint(lo(r)) != __float_as_uint(__fadd_rn(a0, b0))) + (__float_as_uint(hi(r)) != __float_as_uint(__fadd_rn(a1, b1)));
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(A), "l"(B));
        nb += (__float_as_uint(lo(r)) != __float_as_uint(__fsub_rn(a0, b0))) + (__float_as_uint(hi(r)) != __float_as_uint(__fsub_rn(a1, b1)));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(A), "l"(B));
        nb += (__float_as_uint(lo(r)) != __float_as_uint(__fmul_rn(a0, b0))) + (__float_as_uint(hi(r)) != __float_as_uint(__fmul_rn(a1, b1)));
    }
    if (nb) atomicAdd(bad, nb);
}
int main()
{
    unsigned long long *d, h = 0;
    cudaMalloc(&d, 8);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(d, 0, 8);
        k<<<296, 256>>>(d, mode);
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("mode %d: %llu mismatching elements out of %llu packed results\n", mode, h, 296ull * 256 * 20000 * 8);
    }
    return 0;
}
