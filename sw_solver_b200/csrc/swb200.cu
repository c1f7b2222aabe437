// swb200.cu -- C ABI (include/swb200.h) over the sm_100a kernels in swb_kernels.cuh.
//
// Host side of the drop-in boundary: everything here is plain enqueue logic.  The per-step
// scalar state of the reference's loop (time, num_steps, dt, CFL maxima; numpy.py:494-532)
// lives on the device in three rotating `Ctrl` slots, so enqueueing steps never waits for
// the GPU.
#include "swb_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "swb200.h"

using namespace swb;

namespace {

thread_local std::string g_err;

int fail(int code, const char *what)
{
    g_err = what;
    return code;
}

int cuda_fail(cudaError_t e, const char *where)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    g_err = buf;
    return (int)e;
}

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// Every ABI call works on the context's device and leaves the caller's current device as it found it
// (constructing a Stepper on cuda:1 must not move torch's current device).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
        else if (err == cudaSuccess) prev = -1;  // nothing to restore
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define ON_DEVICE(c)                      \
    DeviceGuard guard__((c)->cfg.device); \
    CK(guard__.err)

int valid_cols(int dtype) { return dtype == SWB_F32 ? kStripValidF32 : kStripValid; }
int nstrips_of(int M, int dtype) { return (M + 1 + valid_cols(dtype) - 1) / valid_cols(dtype); }

int required_pitch(int M, int dtype)
{
    // the last strip reads 64 columns (+4 for the diffusion star / topography differences); round to 16 elements
    const int need = valid_cols(dtype) * (nstrips_of(M, dtype) - 1) + kStripCols + 4;
    return (need + 15) / 16 * 16;
}

}  // namespace

struct swb_ctx {
    swb_config cfg;
    Params prm;
    bool tables_set = false, state_bound = false, cfl_ready = false, linked = false;
    double *d_phi = nullptr, *d_tab = nullptr, *d_log = nullptr, *d_pre = nullptr, *d_pred = nullptr, *d_prec = nullptr;
    cudaStream_t setup_stream = nullptr;  // table upload + prerec_kernel (never the caller's or the default stream)
    Ctrl *d_ctrl = nullptr;
    int *d_err = nullptr;
    unsigned int *d_done = nullptr;
    XBlock *d_xblock = nullptr;
    unsigned long long *d_scratch = nullptr;
    swb_status *h_status = nullptr;  // pinned mirror
    int launch = 0;                  // index of the next step launch (mod-3 cycle)
    long long since_init = 0;        // step launches since swb_cfl_init (its parity: the buffer set an active step reads)
    unsigned long long seq = 1;      // stamp of the next step launch (multi-GPU)
    int64_t launches = 0;
    int sm_count = 148;
    bool use_tma = false;    // march_kernel reads through the TMA ring
    bool tma_ready = false;  // tensor maps built (the placement of the arrays allows it)
    bool res_tma = false;    // resident_kernel reads through the TMA ring
    bool cfl_filter = false;  // march_kernel<..., CFLF> + cfl_fixup_kernel (fp64 FAST, large grids)
    int tma_variant = 0;  // index into kTmaVariants (fp64) / kTmaVariantsF32
    TmaMaps maps;
    // resident_kernel (grids that fit the GPU as one wave of blocks): the whole batch of steps in one launch
    int base_chunk = 2;  // marching length of march_kernel (the resident kernel picks its own)
    bool resident = false;
    bool streaming = false;       // resident_kernel<..., STREAM>: one contiguous range of rows per block
    bool resident_bands = false;  // swb_enable_resident_bands: every band of this globe has a GPU of its own
    const void *res_fn = nullptr;
    int res_smem = 0;
    dim3 res_grid;
};

namespace {

// TMA ring shapes built into the library: {latitudes per box R, stages S}
struct TmaVariant { int R, S, minb, W; };
// (three blocks per SM -- 168 registers -- spill and lose 25-40 %)
constexpr TmaVariant kTmaVariants[] = {{4, 4, 2, 4}, {4, 3, 2, 4}, {2, 4, 2, 4}, {4, 4, 1, 8}};
constexpr int kNumTmaVariants = (int)(sizeof(kTmaVariants) / sizeof(kTmaVariants[0]));
constexpr int kDefaultTmaVariant = 0;  // R = 4, S = 4, two blocks per SM: no spills (1.26 ms/step at 16384 x 8192)

// Launch with programmatic stream serialisation (see pdl_enter in swb_kernels.cuh); SWB_PDL=0 launches plainly.
bool pdl_enabled()
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("SWB_PDL");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on != 0;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t lc = {};
    lc.gridDim = grid;
    lc.blockDim = block;
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kernel, std::forward<Args>(args)...);
}

template <typename T, bool STRICT, bool DIFF, int FM, bool HS, bool TMA, int R, int S, bool GUARD, int MINB, int W = kWarpsPerBlock,
          bool CFLF = false>
cudaError_t launch_inst(const swb_ctx *c, dim3 grid, cudaStream_t st)
{
    using LY = TmaLayout<R, S, T, box_cols<T, DIFF, TMA>()>;
    const int smem = rec_bytes(c->prm.chunk, DIFF) + (TMA ? W * LY::kWarpBytes : 0);
    static int configured[64] = {0};  // per instantiation and device: largest size opted in so far
    auto fn = march_kernel<T, STRICT, DIFF, FM, HS, TMA, R, S, GUARD, MINB, W, CFLF>;
    const int dev = c->cfg.device & 63;
    if (configured[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    return launch_chain(fn, grid, dim3(W * 32), (size_t)smem, st, c->prm, c->maps);
}

// mode: SWB_STRICT, SWB_FAST, SWB_FAST_GUARDED;  fm: 0 latitude-only Coriolis, 1 a 2-D array, 2 separable (tables)
template <bool DIFF, int FM, bool HS>
cudaError_t launch_ldg_m(const swb_ctx *c, int mode, dim3 g, cudaStream_t st)
{
    if (mode == SWB_STRICT) return launch_inst<double, true, DIFF, FM, HS, false, 1, 1, true, 2>(c, g, st);
    if (mode == SWB_FAST) return launch_inst<double, false, DIFF, FM, HS, false, 1, 1, false, 2>(c, g, st);
    return launch_inst<double, false, DIFF, FM, HS, false, 1, 1, true, 2>(c, g, st);
}
template <bool DIFF, int FM>
cudaError_t launch_ldg_h(const swb_ctx *c, int mode, dim3 g, cudaStream_t st, bool hs)
{
    return hs ? launch_ldg_m<DIFF, FM, true>(c, mode, g, st) : launch_ldg_m<DIFF, FM, false>(c, mode, g, st);
}
template <bool DIFF>
cudaError_t launch_ldg_f(const swb_ctx *c, int mode, dim3 g, cudaStream_t st, int fm, bool hs)
{
    if (fm == 2) return launch_ldg_m<DIFF, 2, false>(c, mode, g, st);  // (separable Coriolis: flat topography only, swb_create)
    return fm == 1 ? launch_ldg_h<DIFF, 1>(c, mode, g, st, hs) : launch_ldg_h<DIFF, 0>(c, mode, g, st, hs);
}
cudaError_t launch_ldg(const swb_ctx *c, int mode, dim3 g, cudaStream_t st, bool diff, int fm, bool hs)
{
    return diff ? launch_ldg_f<true>(c, mode, g, st, fm, hs) : launch_ldg_f<false>(c, mode, g, st, fm, hs);
}

// separable Coriolis through the TMA ring (default ring shapes only)
template <bool DIFF>
cudaError_t launch_tma_sep(const swb_ctx *c, int mode, dim3 g, cudaStream_t st)
{
    constexpr int S = DIFF ? 3 : 4;
    if (mode == SWB_STRICT) return launch_inst<double, true, DIFF, 2, false, true, 4, S, true, 2>(c, g, st);
    if (mode == SWB_FAST) return launch_inst<double, false, DIFF, 2, false, true, 4, S, false, 2>(c, g, st);
    return launch_inst<double, false, DIFF, 2, false, true, 4, S, true, 2>(c, g, st);
}

template <int R, int S, int MINB, int W = kWarpsPerBlock>
cudaError_t launch_tma_rs(const swb_ctx *c, int mode, dim3 g, cudaStream_t st)
{
    if (mode == SWB_STRICT) return launch_inst<double, true, false, false, false, true, R, S, true, (W > 4 ? 1 : 2), W>(c, g, st);
    if (mode == SWB_FAST) return launch_inst<double, false, false, false, false, true, R, S, false, MINB, W>(c, g, st);
    return launch_inst<double, false, false, false, false, true, R, S, true, MINB, W>(c, g, st);
}
cudaError_t launch_tma(const swb_ctx *c, int mode, dim3 g, cudaStream_t st)
{
    if (c->cfg.f_sep) return c->cfg.use_diffusion ? launch_tma_sep<true>(c, mode, g, st) : launch_tma_sep<false>(c, mode, g, st);
    if (c->cfl_filter) {  // default ring shape only
        if (mode == SWB_FAST) return launch_inst<double, false, false, false, false, true, 4, 4, false, 2, kWarpsPerBlock, true>(c, g, st);
        return launch_inst<double, false, false, false, false, true, 4, 4, true, 2, kWarpsPerBlock, true>(c, g, st);
    }
    if (c->cfg.use_diffusion) {  // three stages of 68-column boxes: the ring feeds the Lax-Wendroff stream and the diffusion star
        if (mode == SWB_STRICT) return launch_inst<double, true, true, false, false, true, 4, 3, true, 2>(c, g, st);
        if (mode == SWB_FAST) return launch_inst<double, false, true, false, false, true, 4, 3, false, 2>(c, g, st);
        return launch_inst<double, false, true, false, false, true, 4, 3, true, 2>(c, g, st);
    }
    switch (c->tma_variant) {
        case 0: return launch_tma_rs<4, 4, 2>(c, mode, g, st);
        case 1: return launch_tma_rs<4, 3, 2>(c, mode, g, st);
        case 2: return launch_tma_rs<2, 4, 2>(c, mode, g, st);
        default: return launch_tma_rs<4, 4, 1, 8>(c, mode, g, st);
    }
}

// fp32 mode: FAST arithmetic only, latitude-only Coriolis, flat topography
constexpr TmaVariant kTmaVariantsF32[] = {{4, 4, 2, 4}, {4, 3, 3, 4}, {4, 3, 4, 4}, {2, 4, 4, 4}};
constexpr int kNumTmaVariantsF32 = (int)(sizeof(kTmaVariantsF32) / sizeof(kTmaVariantsF32[0]));
constexpr int kDefaultTmaVariantF32 = 2;  // R = 4, S = 3, four blocks per SM (0.68 ms/step at 16384 x 8192)

template <int R, int S, int MINB>
cudaError_t launch_tma_f32_rs(const swb_ctx *c, int mode, dim3 g, cudaStream_t st)
{
    if (mode == SWB_FAST) return launch_inst<float, false, false, false, false, true, R, S, false, MINB>(c, g, st);
    return launch_inst<float, false, false, false, false, true, R, S, true, MINB>(c, g, st);
}
cudaError_t launch_tma_f32(const swb_ctx *c, int mode, dim3 g, cudaStream_t st)
{
    if (c->cfg.use_diffusion) {
        if (mode == SWB_FAST) return launch_inst<float, false, true, false, false, true, 4, 3, false, 2>(c, g, st);
        return launch_inst<float, false, true, false, false, true, 4, 3, true, 2>(c, g, st);
    }
    switch (c->tma_variant) {
        case 0: return launch_tma_f32_rs<4, 4, 2>(c, mode, g, st);
        case 1: return launch_tma_f32_rs<4, 3, 3>(c, mode, g, st);
        case 2: return launch_tma_f32_rs<4, 3, 4>(c, mode, g, st);
        default: return launch_tma_f32_rs<2, 4, 4>(c, mode, g, st);
    }
}
cudaError_t launch_ldg_f32(const swb_ctx *c, int mode, dim3 g, cudaStream_t st, bool diff)
{
    if (diff) {
        if (mode == SWB_FAST) return launch_inst<float, false, true, false, false, false, 1, 1, false, 2>(c, g, st);
        return launch_inst<float, false, true, false, false, false, 1, 1, true, 2>(c, g, st);
    }
    if (mode == SWB_FAST) return launch_inst<float, false, false, false, false, false, 1, 1, false, 2>(c, g, st);
    return launch_inst<float, false, false, false, false, false, 1, 1, true, 2>(c, g, st);
}

int launch_one(swb_ctx *c, cudaStream_t st)
{
    const swb_config &cfg = c->cfg;
    const Params &p = c->prm;
    const int rows = p.jl_last - p.jl_first + 1;
    const bool f32 = cfg.dtype == SWB_F32;
    const int W = (c->use_tma && !f32 && !cfg.use_diffusion && !cfg.f_sep) ? kTmaVariants[c->tma_variant].W : kWarpsPerBlock;
    unsigned gy = (unsigned)((rows + p.chunk - 1) / p.chunk);
    if (p.n_big > 0) gy = (unsigned)(p.n_big + (rows - p.n_big * p.chunk + p.chunk_small - 1) / p.chunk_small);
    dim3 grid((unsigned)((p.nstrips + W - 1) / W), gy);
    if (c->linked && !p.fixed) {  // the maxima of all bands' previous step -> this step's slot
        CK(launch_chain(gather_kernel, dim3(1), dim3(32), 0, st, p));
        c->launches++;
        CK(cudaGetLastError());
    }
    if (f32) {
        if (c->use_tma) CK(launch_tma_f32(c, cfg.mode, grid, st));
        else CK(launch_ldg_f32(c, cfg.mode, grid, st, cfg.use_diffusion));
    } else if (c->use_tma) {
        CK(launch_tma(c, cfg.mode, grid, st));
    } else {
        CK(launch_ldg(c, cfg.mode, grid, st, cfg.use_diffusion, cfg.f_sep ? 2 : (cfg.f_is_2d ? 1 : 0), cfg.has_hs));
    }
    c->launches++;
    CK(cudaGetLastError());
    if (c->cfl_filter && !p.fixed) {  // returns at once unless the filter found no valid maximum
        CK(launch_chain(cfl_fixup_kernel, dim3(c->sm_count * 4), dim3(256), 0, st, p));
        c->launches++;
        CK(cudaGetLastError());
    }
    if (c->linked && !p.fixed) {
        CK(launch_chain(publish_kernel, dim3(1), dim3(32), 0, st, p));
        c->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

// ---- resident_kernel: instantiations and set-up --------------------------------------
// Longest marches: with the dt-independent row records kept in shared memory (direct-load kernels,
// short marches) 32 rows; beyond that the TMA kernels scale the records straight from the
// per-context table, and what two blocks per SM leave of the shared memory next to their rings
// holds the records of 96 rows (128 with the three-stage diffusion ring).
constexpr int kResPreSmemChunk = 32;
constexpr int kResFourStageChunk = 96;  // no diffusion: four ring stages up to here, three beyond
constexpr int res_max_chunk(bool tma, bool /*diff*/) { return !tma ? kResPreSmemChunk : 128; }

struct ResKernel { const void *fn; int smem; };
constexpr int kResMixChunk = 24;  // diffusion, one band, marches up to this long: resident_kernel<..., MIX> (mixed pole tiles, see march_step)
template <bool STRICT, bool DIFF, bool TMA, bool GUARD, int S = (DIFF ? 3 : 4), bool BANDS = false, int FM = 0, bool MIX = false>  // (with diffusion: three stages of wider boxes, as in march_kernel)
ResKernel res_kernel_inst(int chunk)
{
    using LY = TmaLayout<4, S, double, box_cols<double, DIFF, TMA>()>;
    auto fn = resident_kernel<double, STRICT, DIFF, TMA, (TMA ? 4 : 1), (TMA ? S : 1), GUARD, 2, BANDS, kWarpsPerBlock, FM, MIX>;
    return ResKernel{(const void *)fn, (chunk <= kResPreSmemChunk ? 2 : 1) * rec_bytes(chunk, DIFF) +
                                           (TMA ? kWarpsPerBlock * LY::kWarpBytes : kWarpsPerBlock * kFirstRowsBytes)};
}
template <bool BANDS>
ResKernel res_kernel_sel(int mode, bool diff, bool tma, int chunk)
{
    if (tma && diff && !BANDS && chunk <= kResMixChunk)
        return mode == SWB_FAST ? res_kernel_inst<false, true, true, false, 3, false, 0, true>(chunk) : res_kernel_inst<false, true, true, true, 3, false, 0, true>(chunk);
    if (tma && diff)
        return mode == SWB_FAST ? res_kernel_inst<false, true, true, false, 3, BANDS>(chunk) : res_kernel_inst<false, true, true, true, 3, BANDS>(chunk);
    if (tma && chunk > kResFourStageChunk)
        return mode == SWB_FAST ? res_kernel_inst<false, false, true, false, 3, BANDS>(chunk) : res_kernel_inst<false, false, true, true, 3, BANDS>(chunk);
    if (tma) return mode == SWB_FAST ? res_kernel_inst<false, false, true, false, 4, BANDS>(chunk) : res_kernel_inst<false, false, true, true, 4, BANDS>(chunk);
    if (mode == SWB_FAST) return diff ? res_kernel_inst<false, true, false, false, 3, BANDS>(chunk) : res_kernel_inst<false, false, false, false, 4, BANDS>(chunk);
    return diff ? res_kernel_inst<false, true, false, true, 3, BANDS>(chunk) : res_kernel_inst<false, false, false, true, 4, BANDS>(chunk);
}
// streaming variant (FAST modes, TMA ring; the rings carry the row records: no block-shared records)
template <bool DIFF, bool GUARD, bool BANDS, bool CFLF>
ResKernel stream_kernel_inst()
{
    constexpr int S = DIFF ? 3 : 4;
    using LY = TmaLayout<4, S, double, box_cols<double, DIFF, true>(), stage_rec_bytes<4>(true, DIFF)>;
    auto fn = stream_kernel<DIFF, S, GUARD, BANDS, CFLF>;
    return ResKernel{(const void *)fn, kWarpsPerBlock * LY::kWarpBytes};
}
template <bool BANDS>
ResKernel stream_kernel_sel(bool guard, bool diff, bool cflf)
{
    if (diff) return guard ? stream_kernel_inst<true, true, BANDS, false>() : stream_kernel_inst<true, false, BANDS, false>();
    if (cflf) return guard ? stream_kernel_inst<false, true, BANDS, true>() : stream_kernel_inst<false, false, BANDS, true>();
    return guard ? stream_kernel_inst<false, true, BANDS, false>() : stream_kernel_inst<false, false, BANDS, false>();
}
ResKernel stream_kernel_of(int mode, bool diff, bool bands, bool cflf)
{
    return bands ? stream_kernel_sel<true>(mode != SWB_FAST, diff, cflf) : stream_kernel_sel<false>(mode != SWB_FAST, diff, cflf);
}

// separable Coriolis (one band): the same shapes with FM = 2
ResKernel res_kernel_sep(int mode, bool diff, bool tma, int chunk)
{
    if (mode == SWB_STRICT) return diff ? res_kernel_inst<true, true, false, true, 3, false, 2>(chunk) : res_kernel_inst<true, false, false, true, 4, false, 2>(chunk);
    const bool guard = mode != SWB_FAST;
    if (tma && diff) return guard ? res_kernel_inst<false, true, true, true, 3, false, 2>(chunk) : res_kernel_inst<false, true, true, false, 3, false, 2>(chunk);
    if (tma && chunk > kResFourStageChunk)
        return guard ? res_kernel_inst<false, false, true, true, 3, false, 2>(chunk) : res_kernel_inst<false, false, true, false, 3, false, 2>(chunk);
    if (tma) return guard ? res_kernel_inst<false, false, true, true, 4, false, 2>(chunk) : res_kernel_inst<false, false, true, false, 4, false, 2>(chunk);
    if (diff) return guard ? res_kernel_inst<false, true, false, true, 3, false, 2>(chunk) : res_kernel_inst<false, true, false, false, 3, false, 2>(chunk);
    return guard ? res_kernel_inst<false, false, false, true, 4, false, 2>(chunk) : res_kernel_inst<false, false, false, false, 4, false, 2>(chunk);
}

// (bands: the FAST modes; a STRICT band keeps one launch per step)
ResKernel res_kernel_of(int mode, bool diff, bool tma, int chunk, bool bands, bool sep = false)
{
    if (sep) return res_kernel_sep(mode, diff, tma, chunk);
    if (bands) return res_kernel_sel<true>(mode, diff, tma, chunk);
    if (!tma && mode == SWB_STRICT) return diff ? res_kernel_inst<true, true, false, true>(chunk) : res_kernel_inst<true, false, false, true>(chunk);
    return res_kernel_sel<false>(mode, diff, tma, chunk);
}

// Decide whether this context's grid runs resident (fp64, one band, latitude-only Coriolis,
// flat topography, all blocks co-resident with marches of at most res_max_chunk() rows) and
// size the launch.  SWB_RESIDENT=0 keeps the launch-per-step kernel, SWB_RES_CHUNK / SWB_KERNEL
// (ldg | tma) override the choices below (tuning and tests).
int setup_resident(swb_ctx *c)
{
    c->resident = false;
    c->streaming = false;
    const swb_config &cfg = c->cfg;
    Params &p = c->prm;
    if (cfg.dtype != SWB_F64 || cfg.f_is_2d || cfg.has_hs) return 0;
    if (cfg.f_sep && cfg.world != 1) return 0;  // (separable Coriolis: the resident kernels serve one band)
    const bool sep = cfg.f_sep != 0;
    // bands: only once the peers are linked AND the caller vouches that every band runs on a GPU of
    // its own (several resident kernels waiting for each other on one GPU would never all start)
    if (cfg.world != 1 && !(c->linked && c->resident_bands && cfg.mode != SWB_STRICT)) return 0;
    if (const char *e = getenv("SWB_RESIDENT"))
        if (atoi(e) == 0) return 0;
    const int rows = p.jl_last - p.jl_first + 1;
    const int gx = (p.nstrips + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const bool tma_ok = c->res_tma && cfg.mode != SWB_STRICT && (cfg.use_diffusion || c->tma_variant == kDefaultTmaVariant);
    const char *force = getenv("SWB_KERNEL");
    int forced_chunk = 0;
    if (const char *e = getenv("SWB_RES_CHUNK")) forced_chunk = atoi(e);
    auto capacity = [&](bool tma, int *cap) -> int {
        // (the largest footprint of this kernel: at the longest march, or at the longest one that keeps two record sets)
        // (every footprint this context may end up with must leave the same number of blocks per SM)
        const int probe[3] = {kResPreSmemChunk, kResFourStageChunk, res_max_chunk(tma, cfg.use_diffusion)};
        int nb_min = 1 << 30;
        for (int k = 0; k < (tma ? 3 : 1); ++k) {
            const ResKernel rk = res_kernel_of(cfg.mode, cfg.use_diffusion, tma, probe[k], cfg.world > 1, sep);
            CK(cudaFuncSetAttribute(rk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, rk.smem));
            int nb = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk.fn, kWarpsPerBlock * 32, (size_t)rk.smem));
            nb_min = nb < nb_min ? nb : nb_min;
        }
        *cap = nb_min * c->sm_count;
        return 0;
    };
    auto chunk_for = [&](int cap, int mult, int max_chunk) -> int {  // shortest march (a multiple of `mult`) with gx * ceil(rows / chunk) <= cap
        const int ny = cap / gx;
        if (ny < 1) return 0;
        int ch = (rows + ny - 1) / ny;
        if (forced_chunk > 0) ch = forced_chunk;
        ch = (ch + mult - 1) / mult * mult;
        if (ch > max_chunk || (long long)gx * ((rows + ch - 1) / ch) > cap) return 0;
        return ch;
    };
    int cap = 0, chunk = 0;
    bool tma = false;
    int rc = capacity(false, &cap);
    if (rc) return rc;
    chunk = chunk_for(cap, 1, res_max_chunk(false, cfg.use_diffusion));
    // marches of 8 rows and more: the TMA ring with its unrolled four-row tiles (measured)
    if (tma_ok && (chunk >= 8 || chunk == 0 || (force && strcmp(force, "tma") == 0))) {
        int cap_t = 0;
        rc = capacity(true, &cap_t);
        if (rc) return rc;
        // (any marching length: the last tile of a march may be partial -- march_step serves it with the fast body;
        // SWB_RES_MULT=4 restores whole four-row tiles)
        int mult = 1;
        if (const char *e = getenv("SWB_RES_MULT")) mult = atoi(e) > 0 ? atoi(e) : 1;
        const int ch_t = chunk_for(cap_t, mult, res_max_chunk(true, cfg.use_diffusion));
        if (ch_t > 0) { tma = true; chunk = ch_t; cap = cap_t; }
    }
    // Streaming variant (stream_kernel): any grid, also beyond one wave of blocks.
    // Opt-in (SWB_STREAM=1): measured on B200 the streaming kernel is bit-identical but 15-25 % slower per row than
    // march_kernel / the one-tile resident kernel (register pressure: its tile loop spills loop-control values, see
    // profiles/r2_stream_kernel.md), which outweighs the launch chain and the wave tail it removes.
    bool stream = false;
    if (const char *e = getenv("SWB_STREAM")) stream = c->tma_ready && cfg.mode != SWB_STRICT && c->d_prec != nullptr && !sep && atoi(e) != 0;
    if (stream) {
        // the CFL filter (estimate per cell, exact evaluation near the previous maximum, fix-up inside the launch):
        // as for march_kernel on grids whose step takes far longer than the barrier; SWB_CFL_FILTER=0 / 1 overrides
        bool cflf = !cfg.use_diffusion && (long long)(p.M + 1) * rows >= 4000000LL;
        if (const char *e = getenv("SWB_CFL_FILTER")) cflf = !cfg.use_diffusion && atoi(e) != 0;
        const ResKernel rk = stream_kernel_of(cfg.mode, cfg.use_diffusion, cfg.world > 1, cflf);
        CK(cudaFuncSetAttribute(rk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, rk.smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk.fn, kWarpsPerBlock * 32, (size_t)rk.smem));
        const long long capb = (long long)nb * c->sm_count;
        if (capb < 1) return 0;
        int map = 1;
        if (const char *e = getenv("SWB_STREAM_MAP")) map = atoi(e) != 0;
        long long forced_unit = 0;
        if (const char *e = getenv("SWB_STREAM_UNIT")) forced_unit = atoll(e);
        long long unit = 0, nblk = 0, total = 0;
        if (map == 0) {  // one contiguous range per block
            total = (long long)gx * rows;
            unit = forced_unit > 0 ? forced_unit : (total + capb - 1) / capb;
            unit = (unit + 3) / 4 * 4;  // whole four-row tiles
            nblk = (total + unit - 1) / unit;
        } else {
            // tiles of `unit` rows dealt round-robin: the tile height with the fewest row-times per block,
            // rounds x (unit + the rows a march start costs); ties go to the longer march
            auto cost = [&](long long u) {
                const long long ntile = (long long)gx * ((rows + u - 1) / u);
                const long long rounds = (ntile + capb - 1) / capb;
                return rounds * (u + 6);
            };
            unit = 4;
            for (long long u = 4; u <= 2048 && u <= ((rows + 3) / 4 * 4); u += 4)
                if (cost(u) <= cost(unit)) unit = u;
            if (forced_unit > 0) unit = (forced_unit + 3) / 4 * 4;
            total = (long long)gx * ((rows + unit - 1) / unit);
            nblk = total < capb ? total : capb;
        }
        if (nblk > capb || nblk < 1 || unit > (1 << 30)) return 0;
        p.stream_total = total;
        p.stream_unit = (int)unit;
        p.stream_rows = rows;
        p.stream_gx = gx;
        p.stream_map = map;
        p.pre_smem = 0;
        c->res_fn = rk.fn;
        c->res_smem = rk.smem;
        c->res_grid = dim3((unsigned)nblk, 1);
        c->resident = true;
        c->streaming = true;
        if (getenv("SWB_VERBOSE"))
            fprintf(stderr, "[swb] streaming resident kernel: %d x %d rows, %d block columns, map %d, unit %lld rows, %lld tiles/rows, %lld blocks "
                            "(%d per SM x %d SMs), %d B of shared memory per block, CFL filter %d\n",
                    cfg.M, rows, gx, map, unit, total, nblk, nb, c->sm_count, rk.smem, (int)cflf);
        return 0;
    }
    if (chunk <= 0) return 0;
    const ResKernel rk = res_kernel_of(cfg.mode, cfg.use_diffusion, tma, chunk, cfg.world > 1, sep);
    CK(cudaFuncSetAttribute(rk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, rk.smem));  // (it may be none of the kernels probed above)
    {
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rk.fn, kWarpsPerBlock * 32, (size_t)rk.smem));
        if ((long long)nb * c->sm_count < (long long)gx * ((rows + chunk - 1) / chunk)) return 0;  // does not fit after all: launch per step
    }
    c->res_fn = rk.fn;
    c->res_smem = rk.smem;
    c->res_grid = dim3((unsigned)gx, (unsigned)((rows + chunk - 1) / chunk));
    p.chunk = chunk;
    p.pre_smem = chunk <= kResPreSmemChunk ? 1 : 0;
    c->resident = true;
    return 0;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D tensor map over the h, u, v arrays of one buffer set: (cols x rows x 3 fields), `pitch`
// elements per row and `field_stride` bytes between consecutive fields
int make_map(CUtensorMap *out, void *base, uint64_t cols, uint64_t rows, uint64_t pitch, uint64_t field_stride, uint32_t box_cols,
             uint32_t box_rows, bool f32)
{
    static encode_tiled_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn) return fail(SWB_E_NODEVICE, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (encode_tiled_fn)fn;
    }
    const cuuint64_t dims[3] = {cols, rows, 3};
    const cuuint64_t strides[2] = {pitch * (f32 ? sizeof(float) : sizeof(double)), field_stride};
    const cuuint32_t box[3] = {box_cols, box_rows, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d (base %p, pitch %llu)", (int)r, base, (unsigned long long)pitch);
        return fail(SWB_E_INVALID, buf);
    }
    return 0;
}

// Marching length.  A lone warp needs ~0.8 us per latitude (one long dependent chain), and a
// march starts with two extra rows, so: the longest chunk (<= 64 rows: two blocks of records
// + rings fit one SM) that still fills about one wave of blocks (2 per SM); small grids end
// up with 2-row marches, large ones with 64 (measured).
int pick_chunk(int rows, int nstrips, int sm_count, bool diffusion)
{
    // (A wave-quantisation model -- waves x (rows + start-up) -- was tried for grids of several waves and
    // dropped: measured at 7200 x 3600, 56 / 64 / 72 / 76 / 96-row marches take 267.9 / 272.1 / 271.9 / 275.1 /
    // 281.9 us per step; at 16384 x 8192, 64 / 72 / 96 rows 1243 / 1248 / 1240 us.  Shorter marches drain
    // better than fuller last waves pay; 64 stays.)
    const long long want_blocks = (long long)sm_count * 2 * 9 / 10;
    const int gx = (nstrips + kWarpsPerBlock - 1) / kWarpsPerBlock;
    // grids of several waves of blocks: 96-row marches (a march start costs ~5 us of a block slot; the guided tail of
    // swb_bind_state takes care of the end of the launch) -- with the tail, 64 / 80 / 96 rows: 16384 x 8192 1342 / 1326 / 1315 us,
    // 7200 x 3600 270.0 / 270.1 / 266.0 us per step
    // (not with diffusion: 7200 x 3600 423 -> 445 us)
    if (!diffusion && (long long)gx * ((rows + 63) / 64) >= 3 * (long long)sm_count * 2) return 96;
    int chunk = 64;
    while (chunk > 2 && (long long)gx * ((rows + chunk - 1) / chunk) < want_blocks) chunk >>= 1;
    return chunk;
}

}  // namespace

extern "C" {

static int create_resources(swb_ctx *c);

int swb_abi_version(void) { return SWB_ABI_VERSION; }
const char *swb_last_error(void) { return g_err.c_str(); }
int swb_required_pitch(int M, int dtype) { return (M >= 2 && (dtype == SWB_F64 || dtype == SWB_F32)) ? required_pitch(M, dtype) : SWB_E_INVALID; }

int swb_create(swb_ctx **out, const swb_config *cfg)
{
    if (!out || !cfg) return fail(SWB_E_INVALID, "swb_create: null argument");
    *out = nullptr;
    if (cfg->struct_bytes != (int32_t)sizeof(swb_config)) return fail(SWB_E_INVALID, "swb_create: swb_config size mismatch (ABI)");
    if (cfg->M < 2 || cfg->N < 2 || (cfg->N & 1)) return fail(SWB_E_INVALID, "swb_create: need M >= 2 and an even N >= 2");
    if (cfg->dtype != SWB_F64 && cfg->dtype != SWB_F32) return fail(SWB_E_INVALID, "swb_create: bad dtype");
    if (cfg->f_sep && (cfg->f_is_2d || cfg->has_hs || cfg->dtype != SWB_F64))
        return fail(SWB_E_INVALID, "swb_create: a separable Coriolis parameter serves fp64 with flat topography (pass the 2-D array otherwise)");
    if (cfg->dtype == SWB_F32 && (cfg->mode == SWB_STRICT || cfg->f_is_2d || cfg->has_hs))
        return fail(SWB_E_INVALID, "swb_create: the fp32 mode serves FAST arithmetic with latitude-only Coriolis and flat topography");
    if (cfg->mode != SWB_STRICT && cfg->mode != SWB_FAST && cfg->mode != SWB_FAST_GUARDED) return fail(SWB_E_INVALID, "swb_create: bad mode");
    if (cfg->n_local < 3) return fail(SWB_E_INVALID, "swb_create: a band needs at least three latitude rows");
    if (cfg->pitch < required_pitch(cfg->M, cfg->dtype) || (cfg->pitch & 3)) return fail(SWB_E_INVALID, "swb_create: pitch must be a multiple of 4 and >= swb_required_pitch(M, dtype)");
    if (cfg->j_begin < 0 || cfg->j_begin + cfg->n_local > cfg->N) return fail(SWB_E_INVALID, "swb_create: band outside the grid");
    if (cfg->j_first < 1 || cfg->j_last > cfg->N - 2 || cfg->j_first > cfg->j_last) return fail(SWB_E_INVALID, "swb_create: bad j_first / j_last");
    if (cfg->j_first - 1 < cfg->j_begin || cfg->j_last + 1 >= cfg->j_begin + cfg->n_local)
        return fail(SWB_E_INVALID, "swb_create: the band must hold one row beyond each updated edge");
    if (cfg->world < 1 || cfg->world > kMaxRanks || cfg->rank < 0 || cfg->rank >= cfg->world) return fail(SWB_E_INVALID, "swb_create: bad rank / world");
    if (cfg->world > 1) {
        const int need = cfg->use_diffusion ? 2 : 1;
        if (cfg->halo < need || cfg->halo > 2) return fail(SWB_E_INVALID, "swb_create: halo must be 1 (2 with diffusion)");
        if ((cfg->j_begin > 0 && cfg->j_first - cfg->halo < cfg->j_begin) ||
            (cfg->j_begin + cfg->n_local < cfg->N && cfg->j_last + cfg->halo >= cfg->j_begin + cfg->n_local))
            return fail(SWB_E_INVALID, "swb_create: the band does not hold `halo` rows beyond its interior edges");
        if (cfg->j_last - cfg->j_first + 1 < cfg->halo) return fail(SWB_E_INVALID, "swb_create: band narrower than the halo");
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(SWB_E_NODEVICE, "swb_create: no CUDA device (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SWB_E_INVALID, "swb_create: bad device ordinal");
    DeviceGuard guard(cfg->device);
    CK(guard.err);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(SWB_E_NODEVICE, "swb_create: kernels are built for sm_100a (Blackwell B200) only");

    swb_ctx *c = new (std::nothrow) swb_ctx();
    if (!c) return fail(SWB_E_INVALID, "swb_create: out of host memory");
    c->cfg = *cfg;
    c->sm_count = prop.multiProcessorCount;
    int rc = create_resources(c);
    if (rc) {  // nothing allocated so far outlives a failed create
        const std::string msg = g_err;
        swb_destroy(c);
        g_err = msg;
        return rc;
    }
    *out = c;
    return 0;
}

static int create_resources(swb_ctx *c)
{
    const swb_config *cfg = &c->cfg;
    const size_t nloc = (size_t)cfg->n_local;
    const size_t nphi = (size_t)cfg->pitch + 8;  // lanes of the last strip look past M+2
    CK(cudaMalloc(&c->d_phi, sizeof(double) * 2 * nphi));  // phi, then the longitude factor of a separable Coriolis parameter
    CK(cudaMemset(c->d_phi, 0, sizeof(double) * 2 * nphi));
    CK(cudaMalloc(&c->d_tab, sizeof(double) * (size_t)T_COUNT * nloc));
    CK(cudaMemset(c->d_tab, 0, sizeof(double) * (size_t)T_COUNT * nloc));
    CK(cudaMalloc(&c->d_pre, sizeof(double) * (size_t)kRowRec * nloc));
    if (cfg->use_diffusion && cfg->mode != SWB_STRICT) CK(cudaMalloc(&c->d_pred, sizeof(double) * (size_t)kRowDiff * nloc));
    if (cfg->mode != SWB_STRICT && cfg->dtype == SWB_F64)  // the streaming kernel's padded table: row record + diffusion record per latitude
        CK(cudaMalloc(&c->d_prec, sizeof(double) * (size_t)rec_stride(cfg->use_diffusion != 0) * (nloc + 2 * kRecPad)));
    CK(cudaStreamCreateWithFlags(&c->setup_stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&c->d_ctrl, sizeof(Ctrl) * 3));
    CK(cudaMemset(c->d_ctrl, 0, sizeof(Ctrl) * 3));
    CK(cudaMalloc(&c->d_err, sizeof(int)));
    CK(cudaMemset(c->d_err, 0, sizeof(int)));
    CK(cudaMalloc(&c->d_done, sizeof(unsigned int) * 4));
    CK(cudaMemset(c->d_done, 0, sizeof(unsigned int) * 4));
    CK(cudaMalloc(&c->d_xblock, sizeof(XBlock)));
    CK(cudaMemset(c->d_xblock, 0, sizeof(XBlock)));
    CK(cudaMalloc(&c->d_scratch, sizeof(unsigned long long) * 4));
    if (cfg->log_capacity > 0) {
        CK(cudaMalloc(&c->d_log, sizeof(double) * 2 * (size_t)cfg->log_capacity));
        CK(cudaMemset(c->d_log, 0, sizeof(double) * 2 * (size_t)cfg->log_capacity));
    }
    CK(cudaMallocHost(&c->h_status, sizeof(swb_status)));
    memset(c->h_status, 0, sizeof(swb_status));

    Params &p = c->prm;
    memset(&p, 0, sizeof p);
    p.phi = c->d_phi;
    p.fa = c->d_phi + nphi;
    p.fs_scale = cfg->fsep_scale;
    p.fs_sa = cfg->fsep_sa;
    p.tab = c->d_tab;
    p.pre = c->d_pre;
    p.pred = c->d_pred;
    p.prec = c->d_prec ? c->d_prec + (size_t)kRecPad * rec_stride(cfg->use_diffusion != 0) : nullptr;
    p.ctrl = c->d_ctrl;
    p.err = c->d_err;
    p.done_ctr = c->d_done;
    p.log = c->d_log;
    p.log_cap = cfg->log_capacity;
    p.M = cfg->M;
    p.nrows = cfg->n_local;
    p.pitch = cfg->pitch;
    p.jl_first = cfg->j_first - cfg->j_begin;
    p.jl_last = cfg->j_last - cfg->j_begin;
    p.pole_s = (cfg->j_begin == 0);
    p.pole_n = (cfg->j_begin + cfg->n_local == cfg->N);
    p.nstrips = nstrips_of(p.M, cfg->dtype);
    p.chunk = pick_chunk(p.jl_last - p.jl_first + 1, p.nstrips, c->sm_count, cfg->use_diffusion != 0);
    if (const char *ch = getenv("SWB_CHUNK")) {  // tuning knob: 1 .. 100 (the row records of longer marches do not fit two blocks per SM)
        const int v = atoi(ch);
        if (v >= 1 && v <= 100) p.chunk = v;
    }
    c->base_chunk = p.chunk;
    if (const char *fr = getenv("SWB_FORCE_RARE")) p.force_rare = atoi(fr) != 0;
    p.courant = cfg->courant;
    p.final_time = cfg->final_time;
    p.dxmin = cfg->dxmin;
    p.dymin = cfg->dymin;
    p.g = cfg->g;
    p.a = cfg->a;
    p.nu = cfg->nu;
    p.dphi = cfg->dphi;
    p.rank = cfg->rank;
    p.world = 1;  // becomes cfg->world once the peers are linked
    p.halo = cfg->halo;
    p.xlocal = c->d_xblock;
    p.spin_limit = 20000000000LL;  // ~10 s of SM clock ticks; SWB_SPIN_SECONDS / swb_set_spin_limit change it
    if (const char *e = getenv("SWB_SPIN_SECONDS")) {
        const double sec = atof(e);
        if (sec > 0.0) p.spin_limit = (long long)(sec * 2.0e9);
    }
#ifdef SWB_BAND_TIMING
    CK(cudaMalloc(&p.tlog, sizeof(unsigned long long) * kTlogCap * kTlogRow));
    CK(cudaMemset(p.tlog, 0, sizeof(unsigned long long) * kTlogCap * kTlogRow));
#endif
    return 0;
}

int swb_destroy(swb_ctx *c)
{
    if (!c) return 0;
    DeviceGuard guard(c->cfg.device);
#ifdef SWB_BAND_TIMING
    if (const char *path0 = getenv("SWB_BAND_TIMING_FILE")) {  // header {world, rank, cap, row, last stamp} + the raw ring
        if (c->prm.tlog && c->seq > 64) {
            std::vector<unsigned long long> h((size_t)kTlogCap * kTlogRow);
            cudaDeviceSynchronize();
            cudaMemcpy(h.data(), c->prm.tlog, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            const std::string path = std::string(path0) + "." + std::to_string(c->cfg.M) + "x" + std::to_string(c->cfg.N) + ".r" + std::to_string(c->cfg.rank);
            if (FILE *f = fopen(path.c_str(), "wb")) {
                const unsigned long long hdr[8] = {(unsigned long long)c->cfg.world, (unsigned long long)c->cfg.rank, (unsigned long long)kTlogCap,
                                                   (unsigned long long)kTlogRow, c->seq, (unsigned long long)c->resident, 0, 0};
                fwrite(hdr, sizeof(unsigned long long), 8, f);
                fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
                fclose(f);
            }
        }
    }
    cudaFree(c->prm.tlog);
#endif
    cudaFree(c->d_phi);
    cudaFree(c->d_tab);
    cudaFree(c->d_pre);
    cudaFree(c->d_pred);
    cudaFree(c->d_prec);
    if (c->setup_stream) cudaStreamDestroy(c->setup_stream);
    cudaFree(c->d_ctrl);
    cudaFree(c->d_err);
    cudaFree(c->d_done);
    cudaFree(c->d_xblock);
    cudaFree(c->d_scratch);
    cudaFree(c->d_log);
    cudaFree(c->prm.dbg);
    cudaFreeHost(c->h_status);
    delete c;
    return 0;
}

int swb_set_tables(swb_ctx *c, const swb_tables *t)
{
    if (!c || !t) return fail(SWB_E_INVALID, "swb_set_tables: null argument");
    if (!t->phi || !t->c || !t->tg || !t->ac || !t->tgMidy || !t->cMidy || !t->dy || !t->dy1 || !t->dyc || !t->dy1c)
        return fail(SWB_E_INVALID, "swb_set_tables: a required table is null");
    if (!c->cfg.f_is_2d && !c->cfg.f_sep && !t->f_lat) return fail(SWB_E_INVALID, "swb_set_tables: f_lat required when f_is_2d == 0");
    if (c->cfg.f_sep && (!t->fsep_lon || !t->fsep_lat_b || !t->fsep_lat_d)) return fail(SWB_E_INVALID, "swb_set_tables: fsep_lon / fsep_lat_b / fsep_lat_d required with f_sep");
    if (c->cfg.use_diffusion && (!t->Ay || !t->By || !t->Cy)) return fail(SWB_E_INVALID, "swb_set_tables: Ay/By/Cy required with diffusion");
    ON_DEVICE(c);
    const size_t nloc = (size_t)c->cfg.n_local;
    std::vector<double> host((size_t)T_COUNT * nloc, 0.0);
    const double *src[T_COUNT] = {t->c, t->tg, t->ac, t->f_lat, t->tgMidy, t->cMidy, t->dy, t->dy1, t->dyc, t->dy1c, t->Ay, t->By, t->Cy,
                                  t->fsep_lat_b, t->fsep_lat_d};
    for (int k = 0; k < T_COUNT; ++k)
        if (src[k]) memcpy(&host[(size_t)k * nloc], src[k], sizeof(double) * nloc);
    // (on the context's own non-blocking stream: neither the default stream nor a device-wide
    // synchronisation is involved; the host waits for this stream only)
    cudaStream_t ss = c->setup_stream;
    CK(cudaMemcpyAsync(c->d_tab, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, ss));
    CK(cudaMemcpyAsync(c->d_phi, t->phi, sizeof(double) * (size_t)(c->cfg.M + 3), cudaMemcpyHostToDevice, ss));
    if (c->cfg.f_sep)
        CK(cudaMemcpyAsync(c->d_phi + ((size_t)c->cfg.pitch + 8), t->fsep_lon, sizeof(double) * (size_t)(c->cfg.M + 3), cudaMemcpyHostToDevice, ss));
    // the dt-independent stage of the row records, once
    const int nb = (c->cfg.n_local + 2 * kRecPad + 127) / 128;
    if (c->cfg.mode == SWB_STRICT) prerec_kernel<true><<<nb, 128, 0, ss>>>(c->prm, c->d_pre, nullptr, nullptr);
    else prerec_kernel<false><<<nb, 128, 0, ss>>>(c->prm, c->d_pre, c->d_pred, c->d_prec);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ss));
    c->tables_set = true;
    return 0;
}

int swb_bind_state(swb_ctx *c, void *h0, void *u0, void *v0, void *h1, void *u1, void *v1, const void *f2d, const void *hs)
{
    if (!c || !h0 || !u0 || !v0 || !h1 || !u1 || !v1) return fail(SWB_E_INVALID, "swb_bind_state: null state pointer");
    if (c->cfg.f_is_2d && !f2d) return fail(SWB_E_INVALID, "swb_bind_state: f_is_2d set but no f array");
    if (c->cfg.has_hs && !hs) return fail(SWB_E_INVALID, "swb_bind_state: has_hs set but no hs array");
    ON_DEVICE(c);  // tensor maps, function attributes and the occupancy query below belong to the context's device
    Params &p = c->prm;
    void *all[8] = {h0, u0, v0, h1, u1, v1, const_cast<void *>(f2d), const_cast<void *>(hs)};
    for (void *q : all)
        if (q && ((uintptr_t)q % 16) != 0) return fail(SWB_E_INVALID, "swb_bind_state: arrays must be 16-byte aligned");
    p.buf[0][0] = h0; p.buf[0][1] = u0; p.buf[0][2] = v0;
    p.buf[1][0] = h1; p.buf[1][1] = u1; p.buf[1][2] = v1;
    p.f2d = f2d;
    p.hs = hs;
    c->state_bound = true;
    c->cfl_ready = false;

    // The TMA ring serves flat topography with latitude-only Coriolis (with or without
    // diffusion); SWB_KERNEL=ldg forces the direct-load kernel, SWB_TMA_VARIANT picks a ring
    // shape.
    const char *force = getenv("SWB_KERNEL");
    const bool plain = !c->cfg.f_is_2d && !c->cfg.has_hs;  // diffusion included: its star is read separately
    // one tensor box carries h, u and v: the three arrays of a buffer set must be equally spaced
    const ptrdiff_t fs0 = (char *)u0 - (char *)h0, fs1 = (char *)u1 - (char *)h1;
    const bool spaced = fs0 > 0 && (char *)v0 - (char *)u0 == fs0 && fs1 > 0 && (char *)v1 - (char *)u1 == fs1 && fs0 % 16 == 0 && fs1 % 16 == 0;
    const bool force_tma = force && strcmp(force, "tma") == 0;
    c->tma_ready = plain && spaced && !(force && strcmp(force, "ldg") == 0);
    c->use_tma = c->tma_ready;
    // With diffusion the ring (which then feeds the star as well) pays off from a size on; below it
    // the direct loads' L1 hits win.  Measured us/step, direct vs ring -- one launch per step:
    // 1440 x 720 29.3 / 32.8, 2000 x 1000 69.7 / 61.1, 4000 x 2000 232 / 198; resident kernel:
    // 1000 x 500 17.2 / 20.6, 1440 x 720 27.1 / 25.7, 2000 x 1000 68.2 / 45.9.
    const long long ncell = (long long)(p.M + 1) * (p.jl_last - p.jl_first + 1);
    c->res_tma = c->tma_ready && (!c->cfg.use_diffusion || ncell >= 1000000LL || force_tma);
    if (c->cfg.use_diffusion && ncell < 2000000LL && !force_tma) c->use_tma = false;
    c->cfl_filter = false;
    if (force_tma && !c->use_tma) return fail(SWB_E_INVALID, "SWB_KERNEL=tma but this configuration cannot use the TMA kernel");
    if (c->tma_ready) {
        const bool f32 = c->cfg.dtype == SWB_F32;
        const char *v = getenv(f32 ? "SWB_TMA_VARIANT_F32" : "SWB_TMA_VARIANT");
        c->tma_variant = v ? atoi(v) : (f32 ? kDefaultTmaVariantF32 : kDefaultTmaVariant);
        if (c->tma_variant < 0 || c->tma_variant >= (f32 ? kNumTmaVariantsF32 : kNumTmaVariants)) return fail(SWB_E_INVALID, "bad SWB_TMA_VARIANT");
        TmaVariant tv = f32 ? kTmaVariantsF32[c->tma_variant] : kTmaVariants[c->tma_variant];
        if (c->cfg.use_diffusion) tv = TmaVariant{4, 3, 2, 4};
        // CFL filter: worth its extra (normally empty) launch per step on grids whose step takes
        // far longer than a launch; SWB_CFL_FILTER=0/1 overrides
        const long long cells = (long long)(p.M + 1) * (p.jl_last - p.jl_first + 1);
        const bool filter_ok = !f32 && !c->cfg.use_diffusion && !c->cfg.f_sep && c->cfg.mode != SWB_STRICT && c->tma_variant == kDefaultTmaVariant;
        c->cfl_filter = filter_ok && cells >= 4000000LL;
        if (const char *e = getenv("SWB_CFL_FILTER")) c->cfl_filter = filter_ok && atoi(e) != 0;
        for (int s = 0; s < 2; ++s) {
            // (fp64 with diffusion: the box also carries the star's two columns either side)
            int rc = make_map(&c->maps.in[s], p.buf[s][0], (uint64_t)(p.M + 3), (uint64_t)p.nrows, (uint64_t)p.pitch, (uint64_t)(s ? fs1 : fs0),
                              (uint32_t)((c->cfg.use_diffusion && !f32) ? box_cols<double, true, true>() : kStripCols), (uint32_t)tv.R, f32);
            if (rc) return rc;
        }
    }
    p.chunk = c->base_chunk;
    // Guided tail (grids of several waves of blocks): the last block rows march 16 latitudes instead of 64.  The
    // hardware dispatches blocks in order, so the short marches run last and the launch drains in a quarter of the
    // time a wave of long marches takes.  SWB_TAIL_ROWS (0 = off) / SWB_TAIL_CHUNK override.
    p.n_big = 0;
    p.chunk_small = 16;
    {
        const int rows = p.jl_last - p.jl_first + 1;
        const int gx = (p.nstrips + kWarpsPerBlock - 1) / kWarpsPerBlock;
        const long long slots = (long long)c->sm_count * 2;
        long long tail_rows = 0;
        // (about one wave of long marches: measured at 7200 x 3600, 0 / 192 / 320 / 640 tail rows give 274.1 / 266.4 / 260.3 /
        // 256.3 us per step; at 16384 x 8192, 0 / 128 / 256 rows 1263 / 1247 / 1246 us; 8- or 32-row tails are no better)
        if (p.chunk >= 32 && (long long)gx * ((rows + 63) / 64) >= 3 * slots) tail_rows = 64LL * ((2 * slots + gx) / (2 * gx));
        if (const char *e = getenv("SWB_TAIL_ROWS")) tail_rows = atoll(e);
        if (const char *e = getenv("SWB_TAIL_CHUNK")) p.chunk_small = atoi(e) > 0 ? atoi(e) : 16;
        // (a second level -- 4-row marches in the very last block rows -- was measured and dropped: 268.6 / 269.7 / 266.8 / 271.9 us
        // with 0 / 40 / 80 / 160 such rows at 7200 x 3600, inside the noise)
        if (tail_rows > 0 && tail_rows < rows && p.chunk_small < p.chunk) {
            p.n_big = (int)((rows - tail_rows) / p.chunk);
            if (p.n_big < 1) p.n_big = 0;
        }
    }
    return setup_resident(c);
}

int swb_resident(const swb_ctx *c) { return (c && c->resident) ? (c->streaming ? 2 : 1) : 0; }

int swb_enable_resident_bands(swb_ctx *c)
{
    if (!c) return fail(SWB_E_INVALID, "swb_enable_resident_bands: null context");
    if (c->cfg.world < 2 || !c->linked) return fail(SWB_E_STATE, "swb_enable_resident_bands: link the peers first (world > 1)");
    if (!c->state_bound) return fail(SWB_E_STATE, "swb_enable_resident_bands: no state bound");
    ON_DEVICE(c);
    c->resident_bands = true;
    c->prm.chunk = c->base_chunk;
    return setup_resident(c);
}

int swb_to_device_layout(swb_ctx *c, const void *ref, int ref_pitch, void *dev, void *stream)
{
    if (!c || !ref || !dev) return fail(SWB_E_INVALID, "swb_to_device_layout: null argument");
    if (ref_pitch < c->cfg.n_local) return fail(SWB_E_INVALID, "swb_to_device_layout: ref_pitch < n_local");
    ON_DEVICE(c);
    const int rows = c->cfg.n_local, cols = c->cfg.M + 3;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    if (c->cfg.dtype == SWB_F32)
        transpose_kernel<double, float><<<grid, 256, 0, (cudaStream_t)stream>>>((const double *)ref, ref_pitch, (float *)dev, c->cfg.pitch, rows, cols);
    else
        transpose_kernel<double, double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double *)ref, ref_pitch, (double *)dev, c->cfg.pitch, rows, cols);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int swb_from_device_layout(swb_ctx *c, const void *dev, void *ref, int ref_pitch, void *stream)
{
    if (!c || !ref || !dev) return fail(SWB_E_INVALID, "swb_from_device_layout: null argument");
    if (ref_pitch < c->cfg.n_local) return fail(SWB_E_INVALID, "swb_from_device_layout: ref_pitch < n_local");
    ON_DEVICE(c);
    const int rows = c->cfg.M + 3, cols = c->cfg.n_local;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    if (c->cfg.dtype == SWB_F32)
        transpose_kernel<float, double><<<grid, 256, 0, (cudaStream_t)stream>>>((const float *)dev, c->cfg.pitch, (double *)ref, ref_pitch, rows, cols);
    else
        transpose_kernel<double, double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double *)dev, c->cfg.pitch, (double *)ref, ref_pitch, rows, cols);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int swb_cfl_init(swb_ctx *c, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_cfl_init: null context");
    if (!c->state_bound) return fail(SWB_E_STATE, "swb_cfl_init: no state bound");
    if (c->cfg.world > 1 && !c->linked) return fail(SWB_E_STATE, "swb_cfl_init: call swb_link_peers first (world > 1)");
    ON_DEVICE(c);
    cudaStream_t st = (cudaStream_t)stream;
    Params &p = c->prm;
    CK(cudaMemsetAsync(c->d_ctrl, 0, sizeof(Ctrl) * 3, st));
    CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), st));
    CK(cudaMemsetAsync(c->d_done, 0, sizeof(unsigned int) * 4, st));  // barrier counter, admission token, halt flag
    // every local row this band owns or mirrors takes part at step 1 (ghosts hold real
    // initial values); rows that are another band's interior are scanned by that band.
    const int jl0 = p.pole_s ? 0 : p.jl_first, jl1 = p.pole_n ? p.nrows : p.jl_last + 1;
    const int blocks = c->sm_count * 8;
    if (c->cfg.dtype == SWB_F32)
        cfl_init_kernel<float><<<blocks, 256, 0, st>>>((const float *)p.buf[0][0], (const float *)p.buf[0][1], (const float *)p.buf[0][2], p.M + 3, jl0, jl1,
                                                       p.pitch, p.g, c->d_ctrl);
    else
        cfl_init_kernel<double><<<blocks, 256, 0, st>>>((const double *)p.buf[0][0], (const double *)p.buf[0][1], (const double *)p.buf[0][2], p.M + 3, jl0,
                                                        jl1, p.pitch, p.g, c->d_ctrl);
    c->launches++;
    CK(cudaGetLastError());
    c->launch = 0;
    c->since_init = 0;
    if (c->linked) {  // deposit the maxima of slot 0 with the first step's stamp
        p.launch = 2;
        p.stamp = c->seq - 1;
        publish_kernel<<<1, 32, 0, st>>>(p);
        c->launches++;
        CK(cudaGetLastError());
    }
    c->cfl_ready = true;
    return 0;
}

int swb_enqueue_steps(swb_ctx *c, int n, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_enqueue_steps: null context");
    if (!c->tables_set || !c->state_bound) return fail(SWB_E_STATE, "swb_enqueue_steps: tables / state not set");
    if (!c->cfl_ready) return fail(SWB_E_STATE, "swb_enqueue_steps: call swb_cfl_init first");
    ON_DEVICE(c);
    cudaStream_t st = (cudaStream_t)stream;
    c->prm.fixed = 0;
    if (c->resident) {  // the whole batch in one cooperative launch
        int left = n;
        while (left > 0) {
            const int m = left < (1 << 20) ? left : (1 << 20);
            c->prm.launch = c->launch;
            c->prm.nsteps = m;
            c->prm.stamp = c->seq;
            CK(cudaMemsetAsync(c->d_done, 0, 2 * sizeof(unsigned int), st));  // arrival counter and admission flag
            void *args[2] = {(void *)&c->prm, (void *)&c->maps};
#ifdef SWB_RES_TIMING
            const size_t nb = (size_t)c->res_grid.x * c->res_grid.y;
            if (!c->prm.dbg) CK(cudaMalloc(&c->prm.dbg, nb * 16 * sizeof(long long)));
            CK(cudaMemsetAsync(c->prm.dbg, 0, nb * 16 * sizeof(long long), st));
#endif
            cudaError_t le = cudaLaunchCooperativeKernel(c->res_fn, c->res_grid, dim3(kWarpsPerBlock * 32), args, (size_t)c->res_smem, st);
            if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported) {
                // the device cannot hold the whole grid right now (shared with another context, MIG
                // slice, ...): this context goes back to one launch per step, for good
                (void)cudaGetLastError();
                c->resident = false;
                c->streaming = false;
                c->prm.chunk = c->base_chunk;
                break;
            }
            CK(le);
#ifdef SWB_RES_TIMING
            if (m >= 100) {  // phase durations (SM clock ticks) of the launch's last-but-one step: mean and max over blocks
                std::vector<long long> hst(nb * 16);
                CK(cudaStreamSynchronize(st));
                CK(cudaMemcpy(hst.data(), c->prm.dbg, nb * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
                const char *names[7] = {"prologue", "records", "startup", "march", "epilogue+sync", "fence+arrive", "wait"};
                fprintf(stderr, "[res-timing] grid %u x %u chunk %d:", c->res_grid.x, c->res_grid.y, c->prm.chunk);
                for (int k = 0; k < 7; ++k) {
                    double sum = 0, mx = 0;
                    for (size_t b = 0; b < nb; ++b) {
                        const double d = (double)(hst[b * 16 + k + 1] - hst[b * 16 + k]);
                        sum += d;
                        if (d > mx) mx = d;
                    }
                    fprintf(stderr, " %s %.0f/%.0f", names[k], sum / nb, mx);
                }
                fprintf(stderr, "\n");
                if (const char *path0 = getenv("SWB_RES_TIMING_FILE")) {  // raw stamps: nb x 16 int64 (0..7 clock64 per phase, 8 = SM id) after a 3-int header
                    const std::string path = std::string(path0) + (c->cfg.world > 1 ? "." + std::to_string(c->cfg.rank) : "");
                    if (FILE *f = fopen(path.c_str(), "wb")) {
                        const int hdr[4] = {(int)c->res_grid.x, (int)c->res_grid.y, c->prm.chunk, 0};
                        fwrite(hdr, sizeof(int), 4, f);
                        fwrite(hst.data(), sizeof(long long), nb * 16, f);
                        fclose(f);
                    }
                }
            }
#endif
            c->launches++;
            c->launch = (c->launch + m) % 3000000;
            c->seq += (unsigned long long)m;
            c->since_init += m;
            left -= m;
        }
        if (c->resident) return 0;
        n = left;
    }
    for (int k = 0; k < n; ++k) {
        c->prm.launch = c->launch;
        c->prm.stamp = c->seq;
        c->prm.par_guess = (int)(c->since_init & 1);
        int rc = launch_one(c, st);
        if (rc) return rc;
        c->launch = (c->launch + 1) % 3000000;  // stays a multiple-of-3 cycle
        c->seq++;
        c->since_init++;
    }
    return 0;
}

int swb_step_fixed_dt(swb_ctx *c, double dt, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_step_fixed_dt: null context");
    if (!c->tables_set || !c->state_bound) return fail(SWB_E_STATE, "swb_step_fixed_dt: tables / state not set");
    ON_DEVICE(c);
    c->prm.fixed = 1;
    c->prm.fixed_dt = dt;
    c->prm.launch = 0;
    c->prm.par_guess = 0;
    int rc = launch_one(c, (cudaStream_t)stream);
    c->prm.fixed = 0;
    return rc;
}

int swb_laplacian(swb_ctx *c, const void *q, void *out, void *stream)
{
    if (!c || !q || !out) return fail(SWB_E_INVALID, "swb_laplacian: null argument");
    if (!c->tables_set || !c->cfg.use_diffusion) return fail(SWB_E_STATE, "swb_laplacian: needs tables with diffusion weights");
    if (c->cfg.dtype != SWB_F64) return fail(SWB_E_STATE, "swb_laplacian: the operator-level Laplacian is fp64 (strict)");
    ON_DEVICE(c);
    const Params &p = c->prm;
    const int rows = p.jl_last - p.jl_first + 1;
    dim3 grid((unsigned)((p.M + 1 + 127) / 128), (unsigned)(rows < 32768 ? rows : 32768));
    laplacian_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const double *)q, (double *)out, p);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int swb_laplacian_ext(swb_ctx *c, const void *qext, int pitch_ext, void *out, void *stream)
{
    if (!c || !qext || !out) return fail(SWB_E_INVALID, "swb_laplacian_ext: null argument");
    if (!c->tables_set || !c->cfg.use_diffusion) return fail(SWB_E_STATE, "swb_laplacian_ext: needs tables with diffusion weights");
    if (c->cfg.dtype != SWB_F64) return fail(SWB_E_STATE, "swb_laplacian_ext: the operator-level Laplacian is fp64 (strict)");
    if (c->cfg.world != 1 || c->cfg.n_local != c->cfg.N) return fail(SWB_E_STATE, "swb_laplacian_ext: one band holding the whole grid only");
    if (pitch_ext < c->cfg.M + 5) return fail(SWB_E_INVALID, "swb_laplacian_ext: pitch_ext < M + 5");
    ON_DEVICE(c);
    const Params &p = c->prm;
    const int rows = p.jl_last - p.jl_first + 1;
    dim3 grid((unsigned)((p.M + 1 + 127) / 128), (unsigned)(rows < 32768 ? rows : 32768));
    laplacian_ext_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const double *)qext, pitch_ext, (double *)out, p);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int swb_status_async(swb_ctx *c, swb_status *host_out, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_status_async: null context");
    static_assert(sizeof(Ctrl) == sizeof(swb_status), "Ctrl mirrors swb_status");
    ON_DEVICE(c);
    swb_status *dst = host_out ? host_out : c->h_status;
    CK(cudaMemcpyAsync(dst, c->d_ctrl + (c->launch % 3), sizeof(Ctrl), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaMemcpyAsync(&dst->error, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}

int swb_status_peek(swb_ctx *c, swb_status *out)
{
    if (!c || !out) return fail(SWB_E_INVALID, "swb_status_peek: null argument");
    memcpy(out, c->h_status, sizeof *out);
    return 0;
}

int swb_get_status(swb_ctx *c, swb_status *out, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_get_status: null context");
    ON_DEVICE(c);  // (also for the synchronisation: a NULL stream means the CURRENT device's default stream)
    int rc = swb_status_async(c, nullptr, stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return swb_status_peek(c, out);
}

int swb_max_speed_async(swb_ctx *c, int current, void *out, void *stream)
{
    if (!c || !out) return fail(SWB_E_INVALID, "swb_max_speed_async: null argument");
    if (!c->state_bound) return fail(SWB_E_STATE, "swb_max_speed_async: no state bound");
    if (current != 0 && current != 1) return fail(SWB_E_INVALID, "swb_max_speed_async: current must be 0 or 1");
    ON_DEVICE(c);
    cudaStream_t st = (cudaStream_t)stream;
    const Params &p = c->prm;
    CK(cudaMemsetAsync(c->d_scratch, 0, sizeof(unsigned long long), st));
    const int jl0 = p.pole_s ? 0 : p.jl_first, jl1 = p.pole_n ? p.nrows : p.jl_last + 1;
    if (c->cfg.dtype == SWB_F32)
        max_speed_kernel<float><<<c->sm_count * 8, 256, 0, st>>>((const float *)p.buf[current][1], (const float *)p.buf[current][2], p.M + 3, jl0, jl1, p.pitch, c->d_scratch);
    else
        max_speed_kernel<double><<<c->sm_count * 8, 256, 0, st>>>((const double *)p.buf[current][1], (const double *)p.buf[current][2], p.M + 3, jl0, jl1, p.pitch, c->d_scratch);
    c->launches++;
    CK(cudaGetLastError());
    // (the bit pattern of a non-negative double IS that double)
    CK(cudaMemcpyAsync(out, c->d_scratch, sizeof(unsigned long long), cudaMemcpyDefault, st));
    return 0;
}

int swb_max_speed(swb_ctx *c, double *out_host, void *stream)
{
    if (!c || !out_host) return fail(SWB_E_INVALID, "swb_max_speed: null argument");
    if (!c->state_bound) return fail(SWB_E_STATE, "swb_max_speed: no state bound");
    ON_DEVICE(c);
    swb_status s;
    int rc = swb_get_status(c, &s, stream);
    if (rc) return rc;
    unsigned long long bits = 0;
    rc = swb_max_speed_async(c, s.current, &bits, stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(out_host, &bits, 8);
    return 0;
}

int swb_set_spin_limit(swb_ctx *c, double seconds)
{
    if (!c || !(seconds > 0.0)) return fail(SWB_E_INVALID, "swb_set_spin_limit: need a context and a positive time");
    c->prm.spin_limit = (long long)(seconds * 2.0e9);  // SM clock ticks at ~2 GHz
    return 0;
}

int swb_read_log(swb_ctx *c, double *dt_out, double *time_out, int count, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_read_log: null context");
    if (count < 0 || count > c->cfg.log_capacity) return fail(SWB_E_INVALID, "swb_read_log: count exceeds log_capacity");
    if (count == 0) return 0;
    ON_DEVICE(c);
    cudaStream_t st = (cudaStream_t)stream;
    if (dt_out) CK(cudaMemcpyAsync(dt_out, c->d_log, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, st));
    if (time_out) CK(cudaMemcpyAsync(time_out, c->d_log + c->cfg.log_capacity, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int64_t swb_launch_count(const swb_ctx *c) { return c ? c->launches : 0; }

// ---- multi-GPU plumbing ------------------------------------------------------------

int swb_ipc_export(const void *dptr, unsigned char handle[SWB_IPC_HANDLE_BYTES], uint64_t *offset)
{
    if (!dptr || !handle || !offset) return fail(SWB_E_INVALID, "swb_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SWB_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
    memcpy(handle, &h, sizeof h);
    // offset of dptr inside its allocation (the handle names the whole allocation):
    // cuMemGetAddressRange through the runtime's driver entry point
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    if (!fn) return fail(SWB_E_INVALID, "swb_ipc_export: cuMemGetAddressRange unavailable");
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    unsigned long long b = 0;
    size_t size = 0;
    if (((range_fn)fn)(&b, &size, (unsigned long long)(uintptr_t)dptr) != 0)
        return fail(SWB_E_INVALID, "swb_ipc_export: cuMemGetAddressRange failed");
    const void *base = (const void *)(uintptr_t)b;
    *offset = (uint64_t)((const char *)dptr - (const char *)base);
    return 0;
}

int swb_ipc_open(const unsigned char handle[SWB_IPC_HANDLE_BYTES], int device, void **base)
{
    if (!handle || !base) return fail(SWB_E_INVALID, "swb_ipc_open: null argument");
    DeviceGuard guard(device);
    CK(guard.err);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CK(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int swb_ipc_close(void *base)
{
    if (!base) return 0;
    CK(cudaIpcCloseMemHandle(base));
    return 0;
}

int swb_exchange_block(swb_ctx *c, void **dptr, uint64_t *bytes)
{
    if (!c || !dptr || !bytes) return fail(SWB_E_INVALID, "swb_exchange_block: null argument");
    *dptr = c->d_xblock;
    *bytes = sizeof(XBlock);
    return 0;
}

int swb_link_peers(swb_ctx *c, void *const *exchange_blocks, void *const south_bufs[6], int south_j_begin, int south_pitch,
                   void *const north_bufs[6], int north_j_begin, int north_pitch)
{
    if (!c || !exchange_blocks) return fail(SWB_E_INVALID, "swb_link_peers: null argument");
    const swb_config &cfg = c->cfg;
    if (cfg.world < 2) return fail(SWB_E_STATE, "swb_link_peers: the context was created with world == 1");
    const bool has_s = cfg.rank > 0, has_n = cfg.rank < cfg.world - 1;
    if ((has_s && !south_bufs) || (has_n && !north_bufs)) return fail(SWB_E_INVALID, "swb_link_peers: a neighbour's buffers are missing");
    Params &p = c->prm;
    for (int r = 0; r < cfg.world; ++r) {
        if (!exchange_blocks[r]) return fail(SWB_E_INVALID, "swb_link_peers: null exchange block");
        p.xpeer[r] = (XBlock *)exchange_blocks[r];
    }
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < 3; ++k) {
            p.south[s][k] = has_s ? south_bufs[s * 3 + k] : nullptr;
            p.north[s][k] = has_n ? north_bufs[s * 3 + k] : nullptr;
            if ((has_s && !p.south[s][k]) || (has_n && !p.north[s][k])) return fail(SWB_E_INVALID, "swb_link_peers: null neighbour buffer");
        }
    // our first updated row (global j_first) is the south neighbour's first halo row above
    // its interior; our last updated row is the north neighbour's last halo row below it
    p.south_row = cfg.j_first - south_j_begin;
    p.south_pitch = south_pitch;
    p.north_row = cfg.j_last - north_j_begin;
    p.north_pitch = north_pitch;
    p.world = cfg.world;
    c->linked = true;
    c->cfl_ready = false;
    return 0;
}

}  // extern "C"
