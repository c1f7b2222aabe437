// Minimal probe of the fp64 2-D TMA load/store primitives used by step_kernel_tma.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe [param|global]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

struct Maps { CUtensorMap in, out; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ void body(const CUtensorMap *min_, const CUtensorMap *mout, int c0, int r0, int flags, const double *gin, double *gout, int pitch, int esz, int lm, int bw)
{
    __shared__ __align__(1024) double tin[2 * 32];
    __shared__ __align__(1024) double tout[2 * 30 + 4];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x;
    const unsigned bar_a = smem_u32(&bar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (!(flags & 1)) {
        for (int r = 0; r < 2; ++r) tin[r * 32 + lane] = gin[(r0 + r) * pitch + c0 + lane];
        __syncwarp();
    }
    if ((flags & 1) && lane == 0) {
        if (lm != 6) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2 * bw * 8) : "memory");
        if (lm == 0 || lm == 4)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tin)), "l"((unsigned long long)min_), "r"(c0 * esz), "r"(r0), "r"(bar_a) : "memory");
        if (lm == 7)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                         ::"r"(smem_u32(tin)), "l"((unsigned long long)min_), "r"(c0 * esz), "r"(r0), "r"(bar_a), "l"(0x1000000000000000ULL) : "memory");
        if (lm == 8)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tin)), "l"((unsigned long long)min_), "r"(c0 * esz), "r"(r0), "r"(bar_a) : "memory");
        if (lm == 3)
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(tin)), "l"((unsigned long long)min_), "r"(c0 * esz), "r"(r0), "r"(bar_a) : "memory");
        if (lm == 2) {  // two 1-D bulk copies of one row each
            for (int r = 0; r < 2; ++r)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tin + r * 32)), "l"(gin + (r0 + r) * pitch + c0 + 1), "r"(256), "r"(bar_a) : "memory");
        }
        if (lm == 6) {  // plain arrive (no tx), no copy
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
        }
    }
    if ((flags & 1) && lm != 4 && lm != 5)
        asm volatile(
            "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
            ::"r"(bar_a), "r"(0) : "memory");
    if ((flags & 1) && (lm == 4 || lm == 5)) { __nanosleep(200000); __syncwarp(); }
    for (int r = 0; r < 2; ++r)
        if (lane >= 1 && lane <= 30) tout[r * 30 + lane - 1] = tin[r * 32 + lane] + 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (!(flags & 2)) {
        for (int r = 0; r < 2; ++r)
            if (lane >= 1 && lane <= 30) gout[(r0 + r) * pitch + c0 + lane] = tout[r * 30 + lane - 1];
    }
    if ((flags & 2) && lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"((unsigned long long)mout), "r"(smem_u32(tout)), "r"((c0 + 1) * esz), "r"(r0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void k_param(const __grid_constant__ Maps m, int c0, int r0, int flags, const double *gin, double *gout, int pitch, int esz, int lm, int bw) { body(&m.in, &m.out, c0, r0, flags, gin, gout, pitch, esz, lm, bw); }
__global__ void k_global(const Maps *m, int c0, int r0, int flags, const double *gin, double *gout, int pitch, int esz, int lm, int bw) { body(&m->in, &m->out, c0, r0, flags, gin, gout, pitch, esz, lm, bw); }

typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const bool use_param = argc < 2 || strcmp(argv[1], "param") == 0;
    const int flags = argc > 2 ? atoi(argv[2]) : 3;
    const int esz = argc > 3 ? atoi(argv[3]) : 1;  // 1: FLOAT64 elements, 2: pairs of UINT32
    const unsigned bw = argc > 6 ? atoi(argv[6]) : 32;  // load box width (elements)
    const int lm = argc > 5 ? atoi(argv[5]) : 0;
    const int promo = argc > 4 ? atoi(argv[4]) : 3;  // CUtensorMapL2promotion 0..3
    const int rows = 9, cols = 48, pitch = 48;
    std::vector<double> h(rows * pitch);
    for (int i = 0; i < rows * pitch; ++i) h[i] = i;
    double *din, *dout;
    cudaMalloc(&din, sizeof(double) * rows * pitch);
    cudaMalloc(&dout, sizeof(double) * rows * pitch);
    cudaMemcpy(din, h.data(), sizeof(double) * rows * pitch, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, sizeof(double) * rows * pitch);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)fn;
    Maps m;
    cuuint64_t dims[2] = {(cuuint64_t)cols * esz, rows}, strides[1] = {pitch * 8};
    cuuint32_t box_in[2] = {bw * esz, 2}, box_out[2] = {30u * esz, 2}, es[2] = {1, 1};
    const CUtensorMapDataType dt = esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    CUresult r1 = enc(&m.in, dt, 2, din, dims, strides, box_in, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&m.out, dt, 2, dout, dims, strides, box_out, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d %d\n", (int)r1, (int)r2);
    const int c0 = argc > 7 ? atoi(argv[7]) : 5, r0 = 3;
    if (use_param) k_param<<<1, 32>>>(m, c0, r0, flags, din, dout, pitch, esz, lm, (int)bw);
    else {
        Maps *dm;
        cudaMalloc(&dm, sizeof(Maps));
        cudaMemcpy(dm, &m, sizeof(Maps), cudaMemcpyHostToDevice);
        k_global<<<1, 32>>>(dm, c0, r0, flags, din, dout, pitch, esz, lm, (int)bw);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("promo=%d lm=%d bw=%u ", promo, lm, bw); printf("%s flags=%d esz=%d: sync -> %s\n", use_param ? "param" : "global", flags, esz, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<double> o(rows * pitch);
    cudaMemcpy(o.data(), dout, sizeof(double) * rows * pitch, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const bool in = r >= r0 && r < r0 + 2 && c >= c0 + 1 && c <= c0 + 30;
            const double want = in ? h[r * pitch + c] + 1.0 : 0.0;
            if (o[r * pitch + c] != want) ++bad;
        }
    printf("mismatches: %d\n", bad);
    return bad != 0;
}
