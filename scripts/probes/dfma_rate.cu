// dev probe: FP64 pipe of a B200 SM sub-partition -- DFMA issue rate against the number of independent chains per warp and warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rate scripts/probes/dfma_rate.cu && ./dfma_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int n)
{
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double c = 1.0001, d = 0.5;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(c), "d"(d));
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int warps)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int n = 4096;
    k<ILP><<<148, warps * 32>>>(out, cyc, n);
    k<ILP><<<148, warps * 32>>>(out, cyc, n);
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / ((double)n * ILP);
    printf("DFMA warps/SM %2d (%d per sub-partition) chains/warp %d: %.2f cycles per instruction per warp; %.3f warp-instructions per cycle per sub-partition\n", warps,
           warps / 4, ILP, per, (warps / 4.0) / per);
    cudaFree(out); cudaFree(cyc);
}
// IEEE division (div.rn.f64: what STRICT pays 13 times per cell) and the reciprocal seed + cubic Newton step of FAST
template <int MODE>
__global__ void kd(double *out, long long *cyc, int n)
{
    constexpr int ILP = 4;
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    const double c = 1.0000001;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = __ddiv_rn(c, a[i]) + 1.0;
            if (MODE == 1) {
                double r;
                asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i]));
                double e = fma(-a[i], r, 1.0);
                e = fma(e, e, e);
                a[i] = fma(r, e, r) + 1.0;
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void rund(const char *name, int warps)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int n = 1024;
    kd<MODE><<<148, warps * 32>>>(out, cyc, n);
    kd<MODE><<<148, warps * 32>>>(out, cyc, n);
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SM %2d: %.1f cycles per result per sub-partition (4 chains per warp; each result + one DADD)\n", name, warps,
           (double)h[0] / ((double)n * 4) / (warps / 4.0));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    rund<0>("div.rn.f64", 8); rund<0>("div.rn.f64", 16); rund<1>("rcp seed + cubic Newton (FAST)", 8); rund<1>("rcp seed + cubic Newton (FAST)", 16);
    for (int w : {4, 8, 12, 16}) { run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); }
    return 0;
}
