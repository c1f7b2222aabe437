// dev probe: where does cvt.f32.f64 (F2F.F32.F64) execute on a B200 SM -- its issue rate alone and next to DFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f2f_rate scripts/probes/f2f_rate.cu && ./f2f_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, long long *cyc, int n)
{
    constexpr int ILP = 8;
    double a[ILP];
    float f[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-3 + i + 1.0; f[i] = 0.f; }
    const double c = 1.0001, d = 0.5;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0 || MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(c), "d"(d));
            if (MODE == 1 || MODE == 2) { float t; asm volatile("cvt.rz.f32.f64 %0, %1;" : "=f"(t) : "d"(a[i])); f[i] += t; }
            if (MODE == 3) {  // the bit-select conversion on the integer pipe
                unsigned hi, lo; asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(a[i]));
                unsigned t = __funnelshift_l(lo, hi, 3);
                t = (t & ~0x40000000u) | (hi & 0x40000000u);
                f[i] += fabsf(__uint_as_float(t));
                a[i] = __hiloint2double(hi + it, lo);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char *name, int warps)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int n = 4096;
    k<MODE><<<148, warps * 32>>>(out, cyc, n);
    k<MODE><<<148, warps * 32>>>(out, cyc, n);
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-22s warps/SM %2d: %.2f cycles per loop trip element per warp (8 elements per trip)\n", name, warps, (double)h[0] / ((double)n * 8));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {8, 16}) {
        if (w == 8) { run<0>("DFMA", 8); run<1>("F2F + FADD", 8); run<2>("DFMA + F2F + FADD", 8); run<3>("bit-select + FADD", 8); }
        else { run<0>("DFMA", 16); run<1>("F2F + FADD", 16); run<2>("DFMA + F2F + FADD", 16); run<3>("bit-select + FADD", 16); }
    }
    return 0;
}
