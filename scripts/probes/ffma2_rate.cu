// dev probe: issue rate of Blackwell's two-wide FP32 instructions against the scalar ones (per SM sub-partition).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate scripts/probes/ffma2_rate.cu && ./ffma2_rate
// Each warp runs ILP independent dependency chains of N instructions; WARPS warps per block, one block per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int ILP>
__global__ void k(float *out, long long *cyc, int n)
{
    float a[ILP], b[ILP];
    unsigned long long p[ILP];
    unsigned q[ILP], q2[ILP];
    for (int i = 0; i < ILP; ++i) { q[i] = threadIdx.x + i; q2[i] = threadIdx.x * 3 + i; }
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = a[i] + 1.0f; asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i])); }
    const float c = 1.0001f, d = 0.5f;
    unsigned long long pc, pd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(c));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pd) : "f"(d));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d)); }
            if (MODE == 1) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc), "l"(pd)); }
            if (MODE == 2) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pd)); }
            if (MODE == 3) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc)); }
            if (MODE == 4) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(d)); }
            if (MODE == 6) {  // packed + an integer-pipe instruction: is the second cycle of a packed instruction an issue slot?
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc), "l"(pd));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[i]) : "r"(it));
            }
            if (MODE == 7) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[i]) : "r"(it));
            }
            if (MODE == 8) {  // packed + two integer-pipe instructions
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc), "l"(pd));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[i]) : "r"(it));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(q2[i]) : "r"(it));
            }
            if (MODE == 5) {  // scalar and packed interleaved
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc), "l"(pd));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi + (float)(q[i] ^ q2[i]); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char *name, int warps)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int n = 4096;
    k<MODE, ILP><<<148, warps * 32>>>(out, cyc, n);
    k<MODE, ILP><<<148, warps * 32>>>(out, cyc, n);
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / ((double)n * ILP * (MODE == 8 ? 3 : (MODE >= 5 ? 2 : 1)));
    // instructions per cycle per sub-partition: warps/4 warps each issue n*ILP instructions in h[0] cycles
    printf("%-28s warps/SM %2d ILP %d: %.3f cycles per instruction per warp; %.3f warp-instructions per cycle per sub-partition\n", name, warps, ILP, per,
           (warps / 4.0) / per);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {4, 8, 16}) {
        if (w == 4) { run<0, 8>("FFMA", 4); run<1, 8>("FFMA2", 4); run<2, 8>("FADD2", 4); run<3, 8>("FMUL2", 4); run<4, 8>("FADD", 4); run<5, 8>("FFMA2+FFMA", 4); }
        if (w == 8) { run<0, 8>("FFMA", 8); run<1, 8>("FFMA2", 8); run<2, 8>("FADD2", 8); run<3, 8>("FMUL2", 8); run<4, 8>("FADD", 8); run<5, 8>("FFMA2+FFMA", 8); }
        if (w == 16) { run<0, 8>("FFMA", 16); run<1, 8>("FFMA2", 16); run<2, 8>("FADD2", 16); run<3, 8>("FMUL2", 16); run<4, 8>("FADD", 16); run<5, 8>("FFMA2+FFMA", 16); }
    }
    run<6, 8>("FFMA2+LOP3", 16); run<7, 8>("FFMA+LOP3", 16); run<8, 8>("FFMA2+2xLOP3", 16);
    run<0, 1>("FFMA latency", 4); run<1, 1>("FFMA2 latency", 4); run<2, 1>("FADD2 latency", 4);
    return 0;
}
