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

}  // namespace

struct swb_ctx {
    swb_config cfg;
    Params prm;
    bool tables_set = false, state_bound = false, cfl_ready = false;
    double *d_phi = nullptr, *d_tab = nullptr, *d_log = nullptr;
    Ctrl *d_ctrl = nullptr;
    unsigned long long *d_scratch = nullptr;
    swb_status *h_status = nullptr;  // pinned mirror
    int launch = 0;                  // index of the next step launch
    int64_t launches = 0;
    int sm_count = 148;
    bool use_tma = false;
    int tma_variant = 0;  // index into kTmaVariants
    TmaMaps maps;
};

namespace {

typedef void (*step_fn)(const Params);

template <bool S, bool D, bool F, bool H>
void launch_step(const Params &p, dim3 grid, cudaStream_t st)
{
    step_kernel<S, D, F, H><<<grid, kWarpsPerBlock * 32, 0, st>>>(p);
}

template <bool S, bool D, bool F>
void pick_h(const Params &p, dim3 g, cudaStream_t st, bool hs)
{
    if (hs) launch_step<S, D, F, true>(p, g, st);
    else launch_step<S, D, F, false>(p, g, st);
}
template <bool S, bool D>
void pick_f(const Params &p, dim3 g, cudaStream_t st, bool f2d, bool hs)
{
    if (f2d) pick_h<S, D, true>(p, g, st, hs);
    else pick_h<S, D, false>(p, g, st, hs);
}
template <bool S>
void pick_d(const Params &p, dim3 g, cudaStream_t st, bool diff, bool f2d, bool hs)
{
    if (diff) pick_f<S, true>(p, g, st, f2d, hs);
    else pick_f<S, false>(p, g, st, f2d, hs);
}

// TMA pipeline shapes built into the library: {rows per box R, stages S}
struct TmaVariant { int R, S; };
constexpr TmaVariant kTmaVariants[] = {{2, 4}, {4, 3}, {4, 4}, {2, 6}, {8, 3}};
constexpr int kNumTmaVariants = (int)(sizeof(kTmaVariants) / sizeof(kTmaVariants[0]));

template <bool STRICT, int R, int S>
cudaError_t launch_tma_rs(const Params &p, const TmaMaps &m, dim3 grid, cudaStream_t st)
{
    using LY = TmaLayout<R, S>;
    static bool configured = false;  // per instantiation; contexts of one process share devices' function attributes
    auto fn = step_kernel_tma<STRICT, R, S>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::kBlockBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    fn<<<grid, kWarpsPerBlock * 32, LY::kBlockBytes, st>>>(p, m);
    return cudaSuccess;
}

template <bool STRICT>
cudaError_t launch_tma(int variant, const Params &p, const TmaMaps &m, dim3 grid, cudaStream_t st)
{
    switch (variant) {
        case 0: return launch_tma_rs<STRICT, 2, 4>(p, m, grid, st);
        case 1: return launch_tma_rs<STRICT, 4, 3>(p, m, grid, st);
        case 2: return launch_tma_rs<STRICT, 4, 4>(p, m, grid, st);
        case 3: return launch_tma_rs<STRICT, 2, 6>(p, m, grid, st);
        default: return launch_tma_rs<STRICT, 8, 3>(p, m, grid, st);
    }
}

int launch_one(swb_ctx *c, cudaStream_t st)
{
    const swb_config &cfg = c->cfg;
    const Params &p = c->prm;
    dim3 grid((unsigned)((p.nstrips + kWarpsPerBlock - 1) / kWarpsPerBlock), (unsigned)((p.M + 1 + p.chunk - 1) / p.chunk));
    if (c->use_tma) {
        CK(cfg.mode == SWB_STRICT ? launch_tma<true>(c->tma_variant, p, c->maps, grid, st)
                                  : launch_tma<false>(c->tma_variant, p, c->maps, grid, st));
    } else if (cfg.mode == SWB_STRICT) {
        pick_d<true>(p, grid, st, cfg.use_diffusion, cfg.f_is_2d, cfg.has_hs);
    } else {
        pick_d<false>(p, grid, st, cfg.use_diffusion, cfg.f_is_2d, cfg.has_hs);
    }
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D fp64 tensor map over a (rows x cols) window of a state array with `pitch` elements per row
int make_map(CUtensorMap *out, void *base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_cols, uint32_t box_rows)
{
    static encode_tiled_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn) return fail(SWB_E_NODEVICE, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (encode_tiled_fn)fn;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch * sizeof(double)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d (base %p, pitch %llu)", (int)r, base, (unsigned long long)pitch);
        return fail(SWB_E_INVALID, buf);
    }
    return 0;
}

// marching length: long enough to amortise the one-row start-up, short enough to fill
// the machine with several waves of warps.
int pick_chunk(int M, int nstrips, int sm_count)
{
    const long long want_warps = (long long)sm_count * 64;  // >= 4 waves of 16 warps/SM
    int chunk = 128;
    while (chunk > 8 && (long long)nstrips * ((M + 1 + chunk - 1) / chunk) < want_warps) chunk >>= 1;
    return chunk;
}

}  // namespace

extern "C" {

int swb_abi_version(void) { return SWB_ABI_VERSION; }
const char *swb_last_error(void) { return g_err.c_str(); }

int swb_create(swb_ctx **out, const swb_config *cfg)
{
    if (!out || !cfg) return fail(SWB_E_INVALID, "swb_create: null argument");
    *out = nullptr;
    if (cfg->struct_bytes != (int32_t)sizeof(swb_config)) return fail(SWB_E_INVALID, "swb_create: swb_config size mismatch (ABI)");
    if (cfg->M < 2 || cfg->N < 2 || (cfg->N & 1)) return fail(SWB_E_INVALID, "swb_create: need M >= 2 and an even N >= 2");
    if (cfg->dtype != SWB_F64) return fail(SWB_E_INVALID, "swb_create: only SWB_F64 is built in this version");
    if (cfg->mode != SWB_STRICT && cfg->mode != SWB_FAST) return fail(SWB_E_INVALID, "swb_create: bad mode");
    if (cfg->lead_pad < 0 || cfg->lead_pad > 8) return fail(SWB_E_INVALID, "swb_create: bad lead_pad");
    if (cfg->n_local < 3 || cfg->pitch < cfg->lead_pad + cfg->n_local) return fail(SWB_E_INVALID, "swb_create: bad n_local / pitch");
    if (cfg->j_begin < 0 || cfg->j_begin + cfg->n_local > cfg->N) return fail(SWB_E_INVALID, "swb_create: band outside the grid");
    if (cfg->j_first < 1 || cfg->j_last > cfg->N - 2 || cfg->j_first > cfg->j_last) return fail(SWB_E_INVALID, "swb_create: bad j_first / j_last");
    if (cfg->j_first - 1 < cfg->j_begin || cfg->j_last + 1 >= cfg->j_begin + cfg->n_local)
        return fail(SWB_E_INVALID, "swb_create: the band must hold one row beyond each updated edge");
    if (cfg->world < 1 || cfg->world > kMaxRanks || cfg->rank < 0 || cfg->rank >= cfg->world) return fail(SWB_E_INVALID, "swb_create: bad rank / world");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(SWB_E_NODEVICE, "swb_create: no CUDA device (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SWB_E_INVALID, "swb_create: bad device ordinal");
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(SWB_E_NODEVICE, "swb_create: kernels are built for sm_100a (Blackwell B200) only");

    swb_ctx *c = new (std::nothrow) swb_ctx();
    if (!c) return fail(SWB_E_INVALID, "swb_create: out of host memory");
    c->cfg = *cfg;
    c->sm_count = prop.multiProcessorCount;
    const size_t nloc = (size_t)(cfg->lead_pad + cfg->n_local);
    CK(cudaMalloc(&c->d_phi, sizeof(double) * (size_t)(cfg->M + 3 + 16)));  // padded: the last tile of a march may peek past M+2
    CK(cudaMemset(c->d_phi, 0, sizeof(double) * (size_t)(cfg->M + 3 + 16)));
    CK(cudaMalloc(&c->d_tab, sizeof(double) * (size_t)T_COUNT * nloc));
    CK(cudaMemset(c->d_tab, 0, sizeof(double) * (size_t)T_COUNT * nloc));
    CK(cudaMalloc(&c->d_ctrl, sizeof(Ctrl) * 3));
    CK(cudaMemset(c->d_ctrl, 0, sizeof(Ctrl) * 3));
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
    p.tab = c->d_tab;
    p.ctrl = c->d_ctrl;
    p.log = c->d_log;
    p.log_cap = cfg->log_capacity;
    p.M = cfg->M;
    p.ncols = cfg->lead_pad + cfg->n_local;
    p.jl_lo = cfg->lead_pad;
    p.jl_hi = p.ncols - 1;
    p.pitch = cfg->pitch;
    p.jl_first = cfg->j_first - cfg->j_begin + cfg->lead_pad;
    p.jl_last = cfg->j_last - cfg->j_begin + cfg->lead_pad;
    p.pole_s = (cfg->j_begin == 0);
    p.pole_n = (cfg->j_begin + cfg->n_local == cfg->N);
    p.nstrips = (p.jl_last - p.jl_first + 1 + kCellsPerWarp - 1) / kCellsPerWarp;
    p.chunk = pick_chunk(p.M, p.nstrips, c->sm_count);
    if (const char *ch = getenv("SWB_CHUNK")) {  // tuning knob: a power of two >= 8
        const int v = atoi(ch);
        if (v >= 8 && (v & (v - 1)) == 0) p.chunk = v;
    }
    p.courant = cfg->courant;
    p.final_time = cfg->final_time;
    p.dxmin = cfg->dxmin;
    p.dymin = cfg->dymin;
    p.g = cfg->g;
    p.a = cfg->a;
    p.nu = cfg->nu;
    p.dphi = cfg->dphi;
    *out = c;
    return 0;
}

int swb_destroy(swb_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    cudaFree(c->d_phi);
    cudaFree(c->d_tab);
    cudaFree(c->d_ctrl);
    cudaFree(c->d_scratch);
    cudaFree(c->d_log);
    cudaFreeHost(c->h_status);
    delete c;
    return 0;
}

int swb_set_tables(swb_ctx *c, const swb_tables *t)
{
    if (!c || !t) return fail(SWB_E_INVALID, "swb_set_tables: null argument");
    if (!t->phi || !t->c || !t->tg || !t->ac || !t->tgMidy || !t->cMidy || !t->dy || !t->dy1 || !t->dyc || !t->dy1c)
        return fail(SWB_E_INVALID, "swb_set_tables: a required table is null");
    if (!c->cfg.f_is_2d && !t->f_lat) return fail(SWB_E_INVALID, "swb_set_tables: f_lat required when f_is_2d == 0");
    if (c->cfg.use_diffusion && (!t->Ay || !t->By || !t->Cy)) return fail(SWB_E_INVALID, "swb_set_tables: Ay/By/Cy required with diffusion");
    CK(cudaSetDevice(c->cfg.device));
    const size_t nloc = (size_t)c->cfg.n_local, pad = (size_t)c->cfg.lead_pad, ncols = nloc + pad;
    std::vector<double> host((size_t)T_COUNT * ncols, 0.0);
    const double *src[T_COUNT] = {t->c, t->tg, t->ac, t->f_lat, t->tgMidy, t->cMidy, t->dy, t->dy1, t->dyc, t->dy1c, t->Ay, t->By, t->Cy};
    for (int k = 0; k < T_COUNT; ++k)
        if (src[k]) {
            for (size_t j = 0; j < pad; ++j) host[(size_t)k * ncols + j] = src[k][0];  // padding columns replay row 0
            memcpy(&host[(size_t)k * ncols + pad], src[k], sizeof(double) * nloc);
        }
    CK(cudaMemcpy(c->d_tab, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_phi, t->phi, sizeof(double) * (size_t)(c->cfg.M + 3), cudaMemcpyHostToDevice));
    c->tables_set = true;
    return 0;
}

int swb_bind_state(swb_ctx *c, void *h0, void *u0, void *v0, void *h1, void *u1, void *v1, const void *f2d, const void *hs)
{
    if (!c || !h0 || !u0 || !v0 || !h1 || !u1 || !v1) return fail(SWB_E_INVALID, "swb_bind_state: null state pointer");
    if (c->cfg.f_is_2d && !f2d) return fail(SWB_E_INVALID, "swb_bind_state: f_is_2d set but no f array");
    if (c->cfg.has_hs && !hs) return fail(SWB_E_INVALID, "swb_bind_state: has_hs set but no hs array");
    Params &p = c->prm;
    p.buf[0][0] = (double *)h0; p.buf[0][1] = (double *)u0; p.buf[0][2] = (double *)v0;
    p.buf[1][0] = (double *)h1; p.buf[1][1] = (double *)u1; p.buf[1][2] = (double *)v1;
    p.f2d = (const double *)f2d;
    p.hs = (const double *)hs;
    c->state_bound = true;
    c->cfl_ready = false;

    // The TMA pipeline serves the plain variant (flat topography, latitude-only Coriolis,
    // no diffusion); SWB_KERNEL=ldg forces the direct-load kernel, SWB_TMA_VARIANT picks a
    // pipeline shape.
    const char *force = getenv("SWB_KERNEL");
    const bool plain = !c->cfg.use_diffusion && !c->cfg.f_is_2d && !c->cfg.has_hs;
    bool aligned = (p.pitch % 2 == 0) && (p.jl_first % 2 == 0) && (p.jl_first >= 2);  // TMA boxes start on 16-byte columns
    for (int s = 0; s < 2 && aligned; ++s)
        for (int k = 0; k < 3; ++k) aligned = aligned && (((uintptr_t)p.buf[s][k]) % 16 == 0);
    c->use_tma = plain && aligned && !(force && strcmp(force, "ldg") == 0);
    if (force && strcmp(force, "tma") == 0 && !c->use_tma) return fail(SWB_E_INVALID, "SWB_KERNEL=tma but this configuration cannot use the TMA kernel");
    if (c->use_tma) {
        const char *v = getenv("SWB_TMA_VARIANT");
        c->tma_variant = v ? atoi(v) : 1;
        if (c->tma_variant < 0 || c->tma_variant >= kNumTmaVariants) return fail(SWB_E_INVALID, "bad SWB_TMA_VARIANT");
        const TmaVariant tv = kTmaVariants[c->tma_variant];
        CK(cudaSetDevice(c->cfg.device));
        for (int s = 0; s < 2; ++s)
            for (int k = 0; k < 3; ++k) {
                int rc = make_map(&c->maps.in[s][k], p.buf[s][k], (uint64_t)p.ncols, (uint64_t)(p.M + 3), (uint64_t)p.pitch, 34, (uint32_t)tv.R);
                if (rc) return rc;
                rc = make_map(&c->maps.out[s][k], p.buf[s][k], (uint64_t)(p.jl_last + 1), (uint64_t)(p.M + 2), (uint64_t)p.pitch, 30, (uint32_t)tv.R);
                if (rc) return rc;
            }
    }
    return 0;
}

int swb_cfl_init(swb_ctx *c, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_cfl_init: null context");
    if (!c->state_bound) return fail(SWB_E_STATE, "swb_cfl_init: no state bound");
    CK(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const Params &p = c->prm;
    CK(cudaMemsetAsync(c->d_ctrl, 0, sizeof(Ctrl) * 3, st));
    // every local row this band owns or mirrors takes part at step 1 (ghosts hold real
    // initial values); rows that are another band's interior are scanned by that band.
    const int jl0 = p.pole_s ? p.jl_lo : p.jl_first, jl1 = p.pole_n ? p.jl_hi + 1 : p.jl_last + 1;
    const int blocks = c->sm_count * 8;
    cfl_init_kernel<<<blocks, 256, 0, st>>>(p.buf[0][0], p.buf[0][1], p.buf[0][2], p.M + 3, jl0, jl1, p.pitch, p.g, c->d_ctrl);
    c->launches++;
    CK(cudaGetLastError());
    c->launch = 0;
    c->cfl_ready = true;
    return 0;
}

int swb_enqueue_steps(swb_ctx *c, int n, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_enqueue_steps: null context");
    if (!c->tables_set || !c->state_bound) return fail(SWB_E_STATE, "swb_enqueue_steps: tables / state not set");
    if (!c->cfl_ready) return fail(SWB_E_STATE, "swb_enqueue_steps: call swb_cfl_init first");
    CK(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    c->prm.fixed = 0;
    for (int k = 0; k < n; ++k) {
        c->prm.launch = c->launch;
        int rc = launch_one(c, st);
        if (rc) return rc;
        c->launch = (c->launch + 1) % 3000000;  // stays a multiple-of-3 cycle
    }
    return 0;
}

int swb_step_fixed_dt(swb_ctx *c, double dt, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_step_fixed_dt: null context");
    if (!c->tables_set || !c->state_bound) return fail(SWB_E_STATE, "swb_step_fixed_dt: tables / state not set");
    CK(cudaSetDevice(c->cfg.device));
    c->prm.fixed = 1;
    c->prm.fixed_dt = dt;
    c->prm.launch = 0;
    int rc = launch_one(c, (cudaStream_t)stream);
    c->prm.fixed = 0;
    return rc;
}

int swb_laplacian(swb_ctx *c, const void *q, void *out, void *stream)
{
    if (!c || !q || !out) return fail(SWB_E_INVALID, "swb_laplacian: null argument");
    if (!c->tables_set || !c->cfg.use_diffusion) return fail(SWB_E_STATE, "swb_laplacian: needs tables with diffusion weights");
    CK(cudaSetDevice(c->cfg.device));
    const Params &p = c->prm;
    dim3 grid((unsigned)((p.jl_last - p.jl_first + 1 + 127) / 128), (unsigned)(p.M + 1));
    laplacian_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const double *)q, (double *)out, p);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

int swb_status_async(swb_ctx *c, swb_status *host_out, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_status_async: null context");
    static_assert(sizeof(Ctrl) == sizeof(swb_status), "Ctrl mirrors swb_status");
    CK(cudaSetDevice(c->cfg.device));
    void *dst = host_out ? (void *)host_out : (void *)c->h_status;
    CK(cudaMemcpyAsync(dst, c->d_ctrl + (c->launch % 3), sizeof(Ctrl), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
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
    int rc = swb_status_async(c, nullptr, stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return swb_status_peek(c, out);
}

int swb_max_speed(swb_ctx *c, double *out_host, void *stream)
{
    if (!c || !out_host) return fail(SWB_E_INVALID, "swb_max_speed: null argument");
    if (!c->state_bound) return fail(SWB_E_STATE, "swb_max_speed: no state bound");
    CK(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    swb_status s;
    int rc = swb_get_status(c, &s, stream);
    if (rc) return rc;
    const Params &p = c->prm;
    const int cur = s.current;
    CK(cudaMemsetAsync(c->d_scratch, 0, sizeof(unsigned long long), st));
    const int jl0 = p.pole_s ? p.jl_lo : p.jl_first, jl1 = p.pole_n ? p.jl_hi + 1 : p.jl_last + 1;
    max_speed_kernel<<<c->sm_count * 8, 256, 0, st>>>(p.buf[cur][1], p.buf[cur][2], p.M + 3, jl0, jl1, p.pitch, c->d_scratch);
    c->launches++;
    CK(cudaGetLastError());
    unsigned long long bits = 0;
    CK(cudaMemcpyAsync(&bits, c->d_scratch, sizeof bits, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out_host, &bits, 8);
    return 0;
}

int swb_read_log(swb_ctx *c, double *dt_out, double *time_out, int count, void *stream)
{
    if (!c) return fail(SWB_E_INVALID, "swb_read_log: null context");
    if (count < 0 || count > c->cfg.log_capacity) return fail(SWB_E_INVALID, "swb_read_log: count exceeds log_capacity");
    if (count == 0) return 0;
    CK(cudaSetDevice(c->cfg.device));
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
    CK(cudaSetDevice(device));
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

int swb_exchange_block(swb_ctx *, void **, uint64_t *) { return fail(SWB_E_STATE, "swb_exchange_block: multi-GPU linking not built yet"); }

int swb_link_peers(swb_ctx *, void *const *, void *const[6], int, int, void *const[6], int, int)
{
    return fail(SWB_E_STATE, "swb_link_peers: multi-GPU linking not built yet");
}

}  // extern "C"
