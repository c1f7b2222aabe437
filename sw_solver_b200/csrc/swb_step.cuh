// swb_step.cuh -- one time step of a block's tile (`march_step`): step prologue (dt from the
// device-resident CFL maxima), row records, TMA ring plumbing, CFL filter, the march itself.
// Called by march_kernel and resident_kernel (swb_kernels.cuh).
#pragma once
#include "swb_scheme.cuh"

namespace swb {

// ------------------------------------------------------------------------------------
// step prologue: dt from the device-resident CFL maxima (numpy.py:524-532)
// ------------------------------------------------------------------------------------
// Executed by thread 0 of every block (every block needs dt); block (0,0) also advances
// the loop state.  Returns through shared memory: dt and flags (bit 0 active, bit 1 parity).
// With several latitude bands (world > 1) `gather_kernel` has replaced the slot's maxima by
// the maxima over all bands before this kernel starts.
// RES (resident_kernel): the launch stops at the first inactive step, so that step leaves
// the carried-forward state in BOTH other slots -- whatever launch index comes next reads it.
template <bool RES>
__device__ __forceinline__ void step_prologue(const Params &p, int launch, bool lead, double *s_dt, int *s_flags)
{
    Ctrl *const cur_slot = p.ctrl + (launch % 3);
    Ctrl *const new_slot = p.ctrl + ((launch + 1) % 3);
    if (p.fixed) {
        *s_dt = p.fixed_dt;
        *s_flags = 1;
        return;
    }
    // (RES: written by other SMs earlier in this launch -- read at the L2)
    const double time = ld1<(RES ? LD_CG : LD_PLAIN)>(&cur_slot->time);
    const long long nsteps = ld1<(RES ? LD_CG : LD_PLAIN)>(&cur_slot->nsteps);
    const unsigned long long exb = ld1<(RES ? LD_CG : LD_PLAIN)>(&cur_slot->ex_bits), eyb = ld1<(RES ? LD_CG : LD_PLAIN)>(&cur_slot->ey_bits);
    const double tx = __ddiv_rn(p.dxmin, __longlong_as_double((long long)exb));
    const double ty = __ddiv_rn(p.dymin, __longlong_as_double((long long)eyb));
    const double dtmax = (tx != tx) ? tx : ((ty != ty) ? ty : (tx < ty ? tx : ty));  // np.minimum propagates NaN
    double dt = __dmul_rn(p.courant, dtmax);
    // (a kernel that timed out waiting for a peer band or for the grid barrier raises done_ctr[2]: the loop ends there
    // instead of running on with unsynchronised halo rows and CFL maxima)
    const bool halted = ld1<(RES ? LD_CG : LD_PLAIN)>(p.done_ctr + 2) != 0u;
    const bool active = time < p.final_time && !halted;  // numpy.py:496 (false for NaN)
    double tnew = time;
    if (active) {
        if (__dadd_rn(time, dt) > p.final_time) {
            dt = __dsub_rn(p.final_time, time);
            tnew = p.final_time;
        } else {
            tnew = __dadd_rn(time, dt);
        }
    }
    *s_dt = dt;
    *s_flags = (active ? 1 : 0) | (int)((nsteps & 1) << 1);
    if (lead) {
        Ctrl *const clr_slot = p.ctrl + ((launch + 2) % 3);
        clr_slot->ex_bits = 0ULL;
        clr_slot->ey_bits = 0ULL;
        clr_slot->reserved = 0;  // found-mask of the CFL filter (march_kernel<..., CFLF>)
        new_slot->time = tnew;
        new_slot->nsteps = nsteps + (active ? 1 : 0);
        new_slot->current = (int)((nsteps + (active ? 1 : 0)) & 1);
        new_slot->last_dt = active ? dt : ld1<(RES ? LD_CG : LD_PLAIN)>(&cur_slot->last_dt);
        new_slot->done = !(tnew < p.final_time);
        new_slot->error = 0;
        if (!active) {  // nothing will be accumulated: carry the maxima forward
            new_slot->ex_bits = exb;
            new_slot->ey_bits = eyb;
            new_slot->reserved = 3;
            if (RES) *clr_slot = *new_slot;
        } else if (p.log != nullptr && nsteps < (long long)p.log_cap) {
            p.log[nsteps] = dt;
            p.log[(size_t)p.log_cap + nsteps] = tnew;
        }
    }
}

// Row records are built in two stages so that a resident kernel can keep the first one:
//   stage 1 (`build_row_pre`): everything that does not depend on dt -- table look-ups and,
//            in FAST mode, the reciprocals of the metric spacings;
//   stage 2 (`scale_row_record`): the record of this time step.  FAST: one multiplication by dt
//            per slot; STRICT: the reference's own quotients hdt/dy1, hdt/dy, dt/dy1c, dt/dyc.
// march_kernel runs both stages every step, resident_kernel stage 1 once per launch: the
// arithmetic, and therefore every bit of the result, is the same.
template <bool STRICT>
__device__ __forceinline__ void build_row_pre(const Params &p, int jl, double *__restrict__ pre)
{
    using O = Op<STRICT>;
    const int nr = p.nrows;
    const int j = min(max(jl, 0), nr - 1), jp = min(j + 1, nr - 1), jm = max(j - 1, 0);
    const double *tab = p.tab;
    const double c = tab[(size_t)T_C * nr + j], tg = tab[(size_t)T_TG * nr + j], ac = tab[(size_t)T_AC * nr + j];
    const double f_j = tab[(size_t)T_F * nr + j], f_jp = tab[(size_t)T_F * nr + jp];
    const double tgm = tab[(size_t)T_TGMIDY * nr + j], cm = tab[(size_t)T_CMIDY * nr + j];
    const double dy = tab[(size_t)T_DY * nr + j], dy1 = tab[(size_t)T_DY1 * nr + j];
    const double dyc = tab[(size_t)T_DYC * nr + j], dy1c = tab[(size_t)T_DY1C * nr + j];
    pre[RC_C] = c;
    pre[RC_AC] = ac;
    pre[RC_CM] = cm;
    pre[RC_DY1S] = O::add(tab[(size_t)T_DY1 * nr + jm], dy1);
    if (STRICT) {
        pre[RC_KX] = tg;   // tgMidx == tan(theta_j): 0.5*(theta+theta) is exact (grid.py:107)
        pre[RC_FKX] = f_j; // 0.5*(f+f) is exact
        pre[RC_CDX] = 0.0;
        pre[RC_CXU] = 0.0;
        pre[RC_KY] = tgm;
        pre[RC_FKY] = O::mul(0.5, O::add(f_jp, f_j));
        pre[RC_CY1F] = dy1;
        pre[RC_CYF] = dy;
        pre[RC_CY1U] = dy1c;
        pre[RC_CYU] = dyc;
        pre[RC_KC] = tg;
        pre[RC_FD] = f_j;
    } else {
        // (times dt in stage 2; hdt = dt/2 and the doubled face values are folded in here)
        const double ra = 1.0 / p.a;
        const double rdx = 1.0 / (ac * p.dphi);  // analytic dx_j (the reference's differs by roundoff only)
        pre[RC_KX] = 0.25 * tg * ra;             // hdt * 0.5 * tan / a
        pre[RC_FKX] = 0.5 * f_j;                 // hdt * f
        pre[RC_CDX] = rdx;                       // dt / dx
        pre[RC_CXU] = 0.5 * rdx;                 // 0.5 dt / dx
        pre[RC_KY] = 0.25 * tgm * ra;
        pre[RC_FKY] = 0.25 * (f_jp + f_j);       // hdt * 0.5 * (f_j + f_j+1)
        pre[RC_CY1F] = 1.0 / dy1;
        pre[RC_CYF] = 1.0 / dy;
        pre[RC_CY1U] = 0.5 / dy1c;
        pre[RC_CYU] = 0.5 / dyc;
        pre[RC_KC] = 0.03125 * tg * ra;          // (dt*0.25) * 0.5[doubled sums] * 0.25 * tan/a
        pre[RC_FD] = 0.125 * f_j;                // (dt*0.25) * 0.5 * f
    }
}

// slot k of the record of time step dt from slot k of the dt-independent record
template <bool STRICT>
__device__ __forceinline__ double scale_row_slot(int k, double pre, double dt)
{
    using O = Op<STRICT>;
    if (STRICT) {
        if (k == RC_CY1F || k == RC_CYF) return O::div(O::mul(0.5, dt), pre);  // hdt / dy1, hdt / dy
        if (k == RC_CY1U || k == RC_CYU) return O::div(dt, pre);               // dt / dy1c, dt / dyc
        return pre;
    }
    constexpr unsigned kPlain = (1u << RC_C) | (1u << RC_AC) | (1u << RC_CM) | (1u << RC_DY1S);
    return ((kPlain >> k) & 1u) ? pre : dt * pre;
}

// FAST diffusion record of local latitude jl: [0..4] combined latitude weights,
// [5] = 1/(2 dx_j)^2 (both independent of dt), [6] = dt*nu (stage 2)
constexpr int kDiffDtSlot = 6;  // the slot of the diffusion record that carries the factor dt
__device__ __forceinline__ void build_diff_pre(const Params &p, int jl, double *__restrict__ rd)
{
    const int nr = p.nrows;
    const int j = min(max(jl, 0), nr - 1);
    const W3 w = y_weights(p.tab, nr, j);
    rd[0] = w.C[1] * w.C[0];
    rd[1] = w.A[1] * w.C[1] + w.C[1] * w.A[0];
    rd[2] = w.A[1] * w.A[1] + w.B[1] * w.C[2] + w.C[1] * w.B[0];
    rd[3] = w.A[1] * w.B[1] + w.B[1] * w.A[2];
    rd[4] = w.B[1] * w.B[2];
    const double dxa = p.tab[(size_t)T_AC * nr + j] * p.dphi;
    rd[5] = 0.25 / (dxa * dxa);
    rd[6] = p.nu;
    rd[7] = 0.0;
}

// ------------------------------------------------------------------------------------
// TMA plumbing
// ------------------------------------------------------------------------------------
struct TmaMaps {
    CUtensorMap in[2];  // per buffer set: box {64 longitudes, R latitudes, 3 fields} over (M+3, nrows, {h, u, v})
};

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// contiguous bytes global -> shared by the TMA unit (size a multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// 16 bytes global -> shared without passing through registers (LDGSTS, L2 only)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
constexpr int kFirstRowsBytes = 6 * 32 * 16;  // per warp: h, u, v pairs of the march's first two latitudes

// Columns of a TMA box: the strip, plus two columns either side when the diffusion star is read
// from the ring as well (fp64 with diffusion: SSTAR in march_step)
template <typename T, bool DIFF, bool TMA>
__host__ __device__ constexpr int box_cols() { return (DIFF && TMA && sizeof(T) == 8) ? kStripCols + 4 : kStripCols; }

// doubles per row of the streamed record table: the row record, then the diffusion record
__host__ __device__ constexpr int rec_stride(bool diff) { return kRowRec + (diff ? kRowDiff : 0); }
// bytes of row records a stage of the streaming kernel carries behind its h, u, v tile: latitudes j0-1 .. j0+R
template <int R>
__host__ __device__ constexpr int stage_rec_bytes(bool stream, bool diff) { return stream ? (R + 2) * rec_stride(diff) * 8 : 0; }

template <int R, int S, typename T = double, int COLS = kStripCols, int EXTRA = 0>
struct TmaLayout {
    static constexpr int kField = R * COLS * (int)sizeof(T);  // one field, one stage
    static constexpr int kStage = 3 * kField + EXTRA;         // (EXTRA: the streamed row records)
    static constexpr int kRing = S * kStage;
    static constexpr int kBars = 128;
    static constexpr int kWarpBytes = kRing + kBars;
};

// bytes of row records in front of the per-warp rings (multiple of 1024)
__host__ __device__ constexpr int rec_bytes(int chunk, bool diff)
{
    return (((chunk + 2) * (kRowRec + (diff ? kRowDiff : 0)) * 8) + 1023) / 1024 * 1024;
}


// fp32 kernels built with -DSWB_F32_PACKED=1 run their hot rows through Blackwell's two-wide FP32 instructions (PK; P2 in
// swb_scheme.cuh).  OFF by default: measured on the headline grid the packed kernel executes 25 % fewer instructions (issue slots
// busy 80 % -> 60 %) but is no faster -- 0.685 against 0.690 ms per step -- because FFMA2 / FADD2 / FMUL2 issue at half the rate of
// the scalar instructions (scripts/probes/ffma2_rate.cu: 0.48 against 0.95 warp-instructions per cycle and sub-partition: the
// FP32 pipe delivers the same flops either way) and the pairs cost ptxas ~30 register moves per latitude and 40 more registers
// (three blocks per SM instead of four); profiles/r2_fp32_ncu.md.  Kept for A/B:
//   SWB_NVCC_FLAGS=-DSWB_F32_PACKED=1 python -m sw_solver_b200.build --out=sw_solver_b200/libswb200_pk.so
//   SWB_LIB=sw_solver_b200/libswb200_pk.so SWB_TEST_F32_PACKED=1 pytest tests/test_gpu_parity.py -k packed
// A packed kernel keeps the coefficients that enter the scheme negated (-dt/dx, -dt/dy ...: PTX cannot negate an operand of a
// two-wide instruction) with their sign flipped in the row record; slot k of such a record, as the scalar code wants it:
#ifndef SWB_F32_PACKED
#define SWB_F32_PACKED 0
#endif
constexpr unsigned kRecNegated = (1u << RC_CDX) | (1u << RC_CXU) | (1u << RC_CY1F) | (1u << RC_CYF) | (1u << RC_CY1U) | (1u << RC_CYU);
template <bool PK, typename T>
__device__ __forceinline__ T rec_slot(const T *rec, int k)
{
    const T v = rec[k];
    return (PK && ((kRecNegated >> k) & 1u)) ? -v : v;
}

// Store a lane's pair of values without branching: both (one 16-byte store), or only the
// first / only the second (the strip's edge lanes and the right end of the grid).
__device__ __forceinline__ void st_pair(double *addr, double a, double b, int both, int only_a, int only_b)
{
    asm volatile(
        "{\n\t"
        ".reg .pred pb, pa, pc;\n\t"
        "setp.ne.s32 pb, %3, 0;\n\t"
        "setp.ne.s32 pa, %4, 0;\n\t"
        "setp.ne.s32 pc, %5, 0;\n\t"
        "@pb st.global.v2.f64 [%0], {%1, %2};\n\t"
        "@pa st.global.f64 [%0], %1;\n\t"
        "@pc st.global.f64 [%0+8], %2;\n\t"
        "}" ::"l"(addr), "d"(a), "d"(b), "r"(both), "r"(only_a), "r"(only_b) : "memory");
}
__device__ __forceinline__ void st_pair(float *addr, float a, float b, int both, int only_a, int only_b)
{
    asm volatile(
        "{\n\t"
        ".reg .pred pb, pa, pc;\n\t"
        "setp.ne.s32 pb, %3, 0;\n\t"
        "setp.ne.s32 pa, %4, 0;\n\t"
        "setp.ne.s32 pc, %5, 0;\n\t"
        "@pb st.global.v2.f32 [%0], {%1, %2};\n\t"
        "@pa st.global.f32 [%0], %1;\n\t"
        "@pc st.global.f32 [%0+4], %2;\n\t"
        "}" ::"l"(addr), "f"(a), "f"(b), "r"(both), "r"(only_a), "r"(only_b) : "memory");
}


// one predicated scalar store (no branch: the unrolled tile stays one basic block)
__device__ __forceinline__ void st_if(double *addr, double a, int pred)
{
    asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.s32 pp, %2, 0;\n\t@pp st.global.f64 [%0], %1;\n\t}" ::"l"(addr), "d"(a), "r"(pred) : "memory");
}
__device__ __forceinline__ void st_if(float *addr, float a, int pred)
{
    asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.s32 pp, %2, 0;\n\t@pp st.global.f32 [%0], %1;\n\t}" ::"l"(addr), "f"(a), "r"(pred) : "memory");
}

// |u| + sqrt(g|h|) and |v| + sqrt(g|h|) of one cell as ordering keys (FAST arithmetic)
template <typename T>
__device__ __forceinline__ void cfl_keys_fast(T h, T u, T v, T g, bits_t<T> &kx, bits_t<T> &ky)
{
    const T cs = fast_sqrt(g * fabs(h));
    kx = nonneg_bits(fabs(u) + cs);
    ky = nonneg_bits(fabs(v) + cs);
}

// CFL filter (fp64 FAST, large grids).  The exact keys above cost 10 FP64-pipe instructions
// per cell; all the step needs is their MAXIMUM, which moves by ~1e-5 per step.  So every
// cell first forms a cheap estimate -- the MUFU.RSQ64H seed instead of the refined root, 4
// instructions -- and only cells whose estimate reaches the trigger threshold (the band's
// previous maximum less 2^-9) are evaluated exactly, from the values just stored.  A result
// counts only if it reaches the validity threshold (previous maximum less 2^-10): every
// cell at or above it is guaranteed to trigger (the seed errs by < 2^-20), so the maximum
// found is the exact one.  If nothing valid was found -- the maximum fell by more than
// 2^-10 within one step -- `cfl_fixup_kernel` recomputes it over the whole band.
struct CflFilter {
    int trig_x, trig_y;                  // hi words of the trigger thresholds
    unsigned long long valid_x, valid_y; // validity thresholds (bit patterns)
};
__device__ __forceinline__ CflFilter make_cfl_filter(unsigned long long exb, unsigned long long eyb)
{
    CflFilter f;
    const double ex = __longlong_as_double((long long)exb), ey = __longlong_as_double((long long)eyb);
    auto trig = [](double m) {
        const double t = m * (1.0 - 1.0 / 512.0);
        return (t > 0.0 && t < 1e300) ? __double2hiint(t) : 0;  // unknown / non-finite: everything triggers
    };
    auto valid = [](double m) {
        const double t = m * (1.0 - 1.0 / 1024.0);
        return (t > 0.0 && t < 1e300) ? (unsigned long long)__double_as_longlong(t) : 0ULL;
    };
    f.trig_x = trig(ex); f.trig_y = trig(ey);
    f.valid_x = valid(ex); f.valid_y = valid(ey);
    return f;
}
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// ------------------------------------------------------------------------------------
// the fused step kernel
// ------------------------------------------------------------------------------------
// Template axes: arithmetic (STRICT / FAST), diffusion, 2-D Coriolis array, topography,
// TMA ring (R latitudes per box, S stages) or direct loads, and whether FAST evaluates the
// reference's hMid > 0 selects (GUARD) or assumes them true and raises SWB_E_DEPTH when a
// half-step depth turns out non-positive.
// SSTAR: where a lane finds its pair of latitudes jl-2, jl-1, jl and jl+2 in the ring (field 0;
// latitude jl+1 is the row the march streams anyway)
template <typename V2>
struct StarRows {
    const V2 *m2, *m1, *c0, *p2;
};

#ifdef SWB_RES_TIMING
__device__ __forceinline__ unsigned swb_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define SWB_STAMP(k)                                                                                       \
    do {                                                                                                   \
        if (RES && p.dbg != nullptr && threadIdx.x == 0 && launch == p.launch + p.nsteps - 2)              \
            p.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (k)] = ((k) == 8) ? (long long)swb_smid() : clock64();                     \
    } while (0)
#else
#define SWB_STAMP(k) do { } while (0)
#endif

// One time step of the block's tile; the body of both step kernels below.  RES = called from
// resident_kernel's loop: coherent (L2) loads of the state, dt-independent row records and
// mbarriers kept across steps (`first` = the launch's first step sets them up; `tbase` counts
// the TMA tiles this warp consumed so far, which fixes ring stage and mbarrier parity), ghost
// columns and pole rows handled inside the unrolled tiles (in a single wave of blocks the
// slowest warp sets the step time), the block's CFL maxima gathered in `s_max`.
// `par_io` (RES): buffer set holding the current state if the caller knows it (>= 0) -- the
// first loads then start before dt is known -- and on return the set holding the new state.
// Returns false when the loop has ended (time >= final_time) -- uniform over the grid.
// STREAM (resident_kernel only): the call covers `tl`, any range of latitudes of one block column; nothing
// is shared between the warps of the block but dt -- the dt-independent row records of latitudes
// j0-1 .. j0+R arrive with every tile (one more bulk copy on the tile's mbarrier, from Params::prec)
// and are scaled by dt in place once the tile has landed.
// FM: how the Coriolis parameter reaches a cell -- 0 latitude-only (folded into the row records), 1 a 2-D array, 2 separable
// (rebuilt per cell from a longitude and two latitude tables in the reference's operation order, see Params::fa).
// RESMIX (resident kernel with diffusion, short marches): a tile that holds a pole row / a neighbour band's halo row runs only THAT
// row through the boundary body and its other rows through the fast one (see the tile loop).
template <typename T, bool STRICT, bool DIFF, int FM, bool HS, bool TMA, int R, int S, bool GUARD, int W, bool CFLF, bool RES, bool RESBANDS = false,
          bool STREAM = false, bool RESMIX = false>
__device__ __forceinline__ bool march_step(const Params &p, const TmaMaps &tm, const int launch, const bool first, int &tbase,
                                           int &par_io, unsigned char *smem_raw, double &s_dt, int &s_flags, CflFilter &s_filter,
                                           unsigned long long *s_max, const Tile tl)
{
    static_assert(!STREAM || (RES && TMA && !STRICT && sizeof(T) == 8), "the streaming kernel: resident, TMA ring, FAST fp64");
    constexpr bool F2D = FM != 0;  // the Coriolis parameter varies along a latitude: it travels with the cells
    static_assert(FM == 0 || sizeof(T) == 8, "2-D Coriolis parameters serve the fp64 kernels");
    static_assert(!STRICT || sizeof(T) == 8, "STRICT reproduces numpy's float64 arithmetic");
    static_assert(!CFLF || (!STRICT && TMA && sizeof(T) == 8), "the CFL filter serves the fp64 FAST TMA kernel");
    static_assert(!RES || (sizeof(T) == 8 && FM != 1 && !HS && (!CFLF || STREAM)), "resident kernels: fp64, Coriolis from tables, flat topography");
    using O = Op<STRICT, T>;
    // SSTAR: the diffusion star comes out of the ring too -- boxes two columns wider on either
    // side, one tile more at either end of the march, tiles t-1, t, t+1 resident while tile t
    // is marched (S >= 3)
    constexpr bool SSTAR = DIFF && TMA && sizeof(T) == 8;
    constexpr int kBoxCols = box_cols<T, DIFF, TMA>();
    constexpr int kLead = SSTAR ? 1 : 0;  // ring tile u holds latitudes ja+1+(u-kLead)R ..; the march consumes u = kLead ..
    static_assert(!SSTAR || (S >= 3 && R == 4), "the star needs three four-row tiles in the ring");
    constexpr int kRecStride = rec_stride(DIFF);                    // STREAM: doubles per row of the streamed records
    constexpr int kRecBytes = stage_rec_bytes<R>(STREAM, DIFF);
    using LY = TmaLayout<R, S, T, kBoxCols, kRecBytes>;
    using V2 = vec2_t<T>;
    constexpr int kValid = sizeof(T) == 8 ? kStripValid : kStripValidF32;
    constexpr bool kGuard = STRICT || GUARD;
    // PK: the fp32 kernels carry a lane's two cells as one packed pair through two-wide instructions (hot rows)
    constexpr bool PK = SWB_F32_PACKED && sizeof(T) == 4 && !STRICT && FM == 0 && !HS;
    static_assert(!PK || !RES, "packed fp32 arithmetic: launch-per-step kernels");
    constexpr int kLdStar = RES ? LD_CA : LD_PLAIN;                   // diffusion star
    constexpr int kLdRow = RES ? (DIFF ? LD_CA : LD_CG) : LD_PLAIN;   // the march's own rows

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int strip = tl.strip0 + warp;
    const bool warp_on = strip < p.nstrips;
    const int ja = tl.ja;  // first latitude updated
    const int jb = tl.jb;  // one past the last
    const int c0 = kValid * strip;
    const int colA = c0 + 2 * lane;  // this lane's cells: (., colA) and (., colA + 1)
    const size_t P = (size_t)p.pitch;

    // ---- TMA ring and the march's first two latitudes ----------------------------------
    const int recb = STREAM ? 0 : rec_bytes(p.chunk, DIFF) * ((RES && p.pre_smem) ? 2 : 1);  // RES, short marches: the dt-independent records follow
    unsigned char *wbase = smem_raw + recb + (size_t)warp * LY::kWarpBytes;
    const uint32_t ring_a = smem_u32(wbase), bar_a = ring_a + LY::kRing;
    const int ntiles = (jb - ja + R - 1) / R;  // tiles marched
    // tiles loaded: with SSTAR one before (latitudes ja-3 .. ja) and, when the last marched
    // tile ends on jb, one after (the star of latitude jb-1 reaches jb+1)
    const int nload = ntiles + (SSTAR ? 1 + ((ntiles * R < jb - ja + 1) ? 1 : 0) : 0);
    const int tb = RES ? tbase : 0;  // ring tiles consumed before this step
    int par = 0;
    const T *__restrict__ H, *__restrict__ U, *__restrict__ V;
    T *__restrict__ Hn, *__restrict__ Un, *__restrict__ Vn;
    const CUtensorMap *maps;
    auto bind = [&](int pr) {
        par = pr;
        H = (const T *)p.buf[pr][0]; U = (const T *)p.buf[pr][1]; V = (const T *)p.buf[pr][2];
        Hn = (T *)p.buf[pr ^ 1][0]; Un = (T *)p.buf[pr ^ 1][1]; Vn = (T *)p.buf[pr ^ 1][2];
        maps = &tm.in[pr];
    };
    auto stage_of = [&](int u) { return (tb + u) % S; };
    auto wait_tile = [&](int u) { mbar_wait(bar_a + 8 * stage_of(u), (uint32_t)(((tb + u) / S) & 1)); };
    auto issue_load = [&](int u) {  // lane 0 only: ring tile u -- h, u, v of R latitudes in one box (out-of-range parts read as zero)
        const int s = stage_of(u);
        const uint32_t bar = bar_a + 8 * s;
        mbar_expect_tx(bar, LY::kStage);
        tma_load_3d(ring_a + s * LY::kStage, maps, c0 - (SSTAR ? 2 : 0), ja + 1 + (u - kLead) * R, 0, bar);
        if (STREAM)  // records of latitudes j0-1 .. j0+R of the tile's marched rows j0 = ja + (u - kLead) R ..  (the table is padded)
            tma_load_1d(ring_a + s * LY::kStage + 3 * LY::kField, p.prec + (ptrdiff_t)(ja - 1 + (u - kLead) * R) * kRecStride, kRecBytes, bar);
    };
    constexpr int kFirstIssue = S - 1 + kLead;  // tiles requested before the march starts
    if (STREAM) {  // every warp, also one whose strip lies outside this block column: a later piece may need its ring
        if (first && lane == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s) mbar_init(bar_a + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    auto start_ring = [&](int u_lo, int u_hi) {
        if (TMA && warp_on) {
            if (lane == 0) {
                if (!STREAM && (!RES || first) && u_lo == 0) {
#pragma unroll
                    for (int s = 0; s < S; ++s) mbar_init(bar_a + 8 * s, 1);
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                }
#pragma unroll
                for (int u = 0; u < kFirstIssue; ++u)
                    if (u >= u_lo && u < u_hi && u < nload) issue_load(u);
            }
            __syncwarp();
        }
    };
    V2 h0, u0, v0, h1, u1, v1;  // latitudes ja-1 and ja
    auto load_first_rows = [&]() {
        const size_t k0 = (size_t)(ja - 1) * P + colA, k1 = k0 + P;
        h0 = ld2<kLdRow>(H + k0); u0 = ld2<kLdRow>(U + k0); v0 = ld2<kLdRow>(V + k0);
        h1 = ld2<kLdRow>(H + k1); u1 = ld2<kLdRow>(U + k1); v1 = ld2<kLdRow>(V + k1);
    };
    // RES, early: the same twelve values parked in shared memory by asynchronous copies (each lane
    // copies and later reads its own 6 x 16 bytes) -- in the ring stage that is not in flight yet
    // (TMA), or in a staging area behind the records
    unsigned char *stg = (TMA ? wbase + (size_t)((tb + S - 1) % S) * LY::kStage : smem_raw + recb + (size_t)warp * kFirstRowsBytes) + lane * 16;
    constexpr int kEarly = 1 + kLead;  // tiles requested before dt is known (RES); SSTAR: its first tile holds the first rows
    auto first_rows_async = [&]() {
        const size_t k0 = (size_t)(ja - 1) * P + colA, k1 = k0 + P;
        const uint32_t d = smem_u32(stg);
        cp_async16(d, H + k0); cp_async16(d + 512, U + k0); cp_async16(d + 1024, V + k0);
        cp_async16(d + 1536, H + k1); cp_async16(d + 2048, U + k1); cp_async16(d + 2560, V + k1);
    };
    auto first_rows_landed = [&]() {
        cp_async_wait_all();
        const V2 *q = reinterpret_cast<const V2 *>(stg);
        h0 = q[0]; u0 = q[32]; v0 = q[64]; h1 = q[96]; u1 = q[128]; v1 = q[160];
    };
    // RES, second step onwards: the parity is known, so the first tile is requested while thread 0
    // works out dt (only the first: the SM's whole share of the state arrives in microseconds anyway,
    // and a deeper prefetch by every warp at once only queues up in front of the first tiles)
    // march_kernel (fp64 TMA kernels; round 2): the host passes the buffer parity it expects -- all launches since swb_cfl_init
    // active; if they were not, this step is inactive too and nothing of it is used -- so the first tile and the first rows
    // are requested before thread 0 has worked out dt (7200 x 3600: 257.9 -> 254.1 us per step; the 16-byte cp.async of the
    // first rows needs the fp64 pairs' alignment, hence not fp32)
    const bool early = (RES || (TMA && sizeof(T) == 8)) && par_io >= 0;
    if (early) {
        bind(par_io);
        if (warp != 0) {  // (warp 0: after the prologue -- its lane 0 is the prologue's thread)
            start_ring(0, kEarly);
            if (warp_on && !SSTAR) first_rows_async();
        }
    }
    SWB_STAMP(9);

    if (threadIdx.x == 0 && (!STREAM || tl.prologue)) {
        step_prologue<RES>(p, launch, tl.lead, &s_dt, &s_flags);
        if (RES) s_max[0] = s_max[1] = 0ULL;
        if (RES && CFLF) s_max[2] = 0ULL;
        if (CFLF) {
            // thresholds from THIS band's maxima of the current state (with several bands the
            // slot holds the global maxima; the band's own live in its exchange block)
            // (RES: written by other SMs earlier in this launch -- read at the L2)
            constexpr int kLd = RES ? LD_CG : LD_PLAIN;
            const Ctrl *cur_slot = p.ctrl + (launch % 3);
            unsigned long long exb = ld1<kLd>(&cur_slot->ex_bits), eyb = ld1<kLd>(&cur_slot->ey_bits);
            if (p.world > 1) {
                exb = ld1<kLd>(&p.xlocal->slot[launch % 3][p.rank].ex_bits);
                eyb = ld1<kLd>(&p.xlocal->slot[launch % 3][p.rank].ey_bits);
            }
            s_filter = p.fixed ? make_cfl_filter(0ULL, 0ULL) : make_cfl_filter(exb, eyb);
        }
    }
    SWB_STAMP(10);
    if (early && warp == 0) {
        start_ring(0, kEarly);
        if (!SSTAR) first_rows_async();
    }
    SWB_STAMP(11);
    if (!STREAM || tl.prologue) __syncthreads();  // (STREAM: later pieces of a step find dt and the flags in place)
    SWB_STAMP(1);
    const int flags = s_flags;
    if (!(flags & 1)) {
        if (early && TMA && warp_on && lane == 0)  // the loads already under way must land before the block may exit
            for (int u = 0; u < kEarly && u < nload; ++u) wait_tile(u);
        return false;
    }
    if (!RES && early && ((flags >> 1) & 1) != par) {  // cannot happen (see above); if it did, the prefetched tile would be the wrong buffer's
        if (TMA && warp_on && lane == 0)
            for (int u = 0; u < kEarly && u < nload; ++u) wait_tile(u);
        if (threadIdx.x == 0) atomicMin(p.err, -2);  // SWB_E_STATE
        return false;
    }
    if (!early) bind((flags >> 1) & 1);
    start_ring(early ? kEarly : 0, kFirstIssue);
    if (RES) par_io = par ^ 1;
    const T dt = (T)s_dt;
    const T g_ = (T)p.g;

    // ---- row records of latitudes ja-1 .. jb ----------------------------------------
    // (built in fp64 whatever T is -- SURVEY 7, hard part 6 -- and rounded once)
    T *rowc = reinterpret_cast<T *>(smem_raw);
    T *rowd = rowc + (size_t)(p.chunk + 2) * kRowRec;
    if (STREAM) {
        // nothing here: the records travel with the tiles
    } else if (RES) {  // (march_kernel reading the same table instead of rebuilding stage 1: measured, no gain -- 16384 x 8192 1337 -> 1349 us)
        // stage 1 (everything that does not depend on dt: look-ups, reciprocals of the metric
        // spacings) was computed once per context by prerec_kernel and is kept in shared memory for
        // the launch; stage 2 is one multiplication per slot (STRICT: the reference's own
        // quotients), spread over the block's threads.
        const int nrec = jb - ja + 2;
        const double *src = p.pre + (size_t)(ja - 1) * kRowRec, *srcd = (DIFF && !STRICT) ? p.pred + (size_t)(ja - 1) * kRowDiff : nullptr;
        if (p.pre_smem) {  // short marches: stage 1 stays in shared memory for the launch
            double *prec = reinterpret_cast<double *>(smem_raw + rec_bytes(p.chunk, DIFF));
            double *pred = prec + (size_t)(p.chunk + 2) * kRowRec;
            if (first) {
                for (int e = threadIdx.x; e < nrec * kRowRec; e += blockDim.x) prec[e] = src[e];
                if (DIFF && !STRICT)
                    for (int e = threadIdx.x; e < nrec * kRowDiff; e += blockDim.x) pred[e] = srcd[e];
                __syncthreads();
            }
            src = prec;
            srcd = pred;
        }
        const double dt64 = s_dt;
        for (int e = threadIdx.x; e < nrec * kRowRec; e += blockDim.x) rowc[e] = (T)scale_row_slot<STRICT>(e % kRowRec, src[e], dt64);
        if (DIFF && !STRICT)
            for (int e = threadIdx.x; e < nrec * kRowDiff; e += blockDim.x) rowd[e] = (T)((e % kRowDiff) == kDiffDtSlot ? dt64 * srcd[e] : srcd[e]);
    } else {
        // one launch per step: both stages here, one latitude per thread (the same arithmetic; a
        // table look-up instead was tried and cost the hot loop registers)
        for (int r = threadIdx.x; r < jb - ja + 2; r += blockDim.x) {
            double pre[kRowRec];
            build_row_pre<STRICT>(p, ja - 1 + r, pre);
#pragma unroll
            for (int k = 0; k < kRowRec; ++k) {
                const T val = (T)scale_row_slot<STRICT>(k, pre[k], s_dt);
                rowc[(size_t)r * kRowRec + k] = (PK && ((kRecNegated >> k) & 1u)) ? -val : val;
            }
            if (DIFF && !STRICT) {
                double rd[kRowDiff];
                build_diff_pre(p, ja - 1 + r, rd);
#pragma unroll
                for (int k = 0; k < kRowDiff; ++k) rowd[(size_t)r * kRowDiff + k] = (T)(k == kDiffDtSlot ? s_dt * rd[k] : rd[k]);
            }
        }
    }
    if (!STREAM) __syncthreads();
    SWB_STAMP(2);
    if (!warp_on) return true;
    if (RES) tbase = tb + nload;
    constexpr int kFieldV = LY::kField / (int)sizeof(V2), kRowV = kBoxCols / 2;  // ring strides in lane pairs
    auto tile_of = [&](int u) {  // this lane's pair in row 0, field 0 of ring tile u
        return reinterpret_cast<const V2 *>(wbase + (size_t)stage_of(u) * LY::kStage) + lane + (SSTAR ? 1 : 0);
    };
    if (SSTAR) {  // latitudes ja-1, ja: the last two rows of the leading tile
        wait_tile(0);
        const V2 *q = tile_of(0);
        h0 = q[(R - 2) * kRowV]; u0 = q[kFieldV + (R - 2) * kRowV]; v0 = q[2 * kFieldV + (R - 2) * kRowV];
        h1 = q[(R - 1) * kRowV]; u1 = q[kFieldV + (R - 1) * kRowV]; v1 = q[2 * kFieldV + (R - 1) * kRowV];
    } else if (early) {
        first_rows_landed();
    } else {
        load_first_rows();
    }

    // ---- per-lane set-up ------------------------------------------------------------
    const bool vA = (lane >= 1) && (colA <= c0 + kValid) && (colA <= p.M + 1);
    const bool vB = (colA + 1 <= c0 + kValid) && (colA + 1 <= p.M + 1);
    const T hdt = O::mul(T(0.5), dt);
    const T hgCell = O::mul(T(0.5), g_);               // numpy.py:203, 255: 0.5*g
    const T hgK = STRICT ? hgCell : T(0.25) * g_;      // FAST faces are doubled
    const T raK = STRICT ? (T)p.a : (T)(1.0 / p.a);
    T ph_m = 0, ph_0 = 0, ph_1 = 0, ph_2 = 0;  // STRICT: phi at colA-1 .. colA+2
    if (STRICT) {
        ph_m = (T)p.phi[max(colA - 1, 0)];
        ph_0 = (T)p.phi[colA];
        ph_1 = (T)p.phi[colA + 1];
        ph_2 = (T)p.phi[colA + 2];
    }
    // extra copies a cell owes: periodic ghost columns (numpy.py:402)
    const bool wrapA = vA && (colA == p.M || colA == 2), wrapB = vB && (colA + 1 == p.M || colA + 1 == 2);
    // with diffusion, column M+1 reaches across the periodic seam too (its i+2 is column 3)
    const bool seam = DIFF && ((vA && colA == p.M + 1) || (vB && colA + 1 == p.M + 1));
    const bool warp_wraps = __any_sync(0xffffffffu, wrapA || wrapB || seam);
    const int st_both = (vA && vB) ? 1 : 0, st_onlyA = (vA && !vB) ? 1 : 0, st_onlyB = (vB && !vA) ? 1 : 0;
    // resident kernel: the ghost-column copy of a wrapping cell as one predicated store per field
    // (column M -> column 0, column 2 -> column M+2; M >= 4 so that no column is both)
    const bool res_wrap = RES && p.M >= 4;
    const int gpA = (res_wrap && wrapA) ? 1 : 0, gpB = (res_wrap && wrapB) ? 1 : 0;
    const long long goA = (long long)((colA == p.M) ? 0 : p.M + 2) - colA;      // element offset from (jl, colA) to A's ghost column
    const long long goB = (long long)((colA + 1 == p.M) ? 0 : p.M + 2) - colA;  // ... and to B's
    bits_t<T> mxA = 0, myA = 0, mxB = 0, myB = 0;  // running CFL maxima (bit patterns)
    const int trig_x = CFLF ? s_filter.trig_x : 0, trig_y = CFLF ? s_filter.trig_y : 0;
    bool trig = false;  // CFLF: some cell of the current tile reached a trigger threshold
    int negA = 0, negB = 0;  // sign watch of the half-step depths (FAST without GUARD)

    // STREAM: the records of the tile being marched (latitudes rec_j0 ..) sit behind its h, u, v in the ring
    const T *rec_tile = rowc;
    int rec_j0 = ja - 1;
    constexpr int kRecStep = STREAM ? kRecStride : kRowRec;  // elements between the records of consecutive latitudes
    auto rec_of = [&](int jl) { return STREAM ? rec_tile + (jl - rec_j0) * kRecStep : rowc + (size_t)(jl - (ja - 1)) * kRecStep; };
    auto set_rec_tile = [&](int u) {
        rec_tile = reinterpret_cast<const T *>(wbase + (size_t)stage_of(u) * LY::kStage + 3 * LY::kField);
        rec_j0 = ja - 1 + (u - kLead) * R;
    };
    // STREAM: a landed tile's records times dt, in place (slot k of a row: scale_row_slot; the diffusion
    // part: slot 6).  The same products as the block-shared records of the other kernels.
    auto scale_tile = [&](int u) {
        double2 *rec = reinterpret_cast<double2 *>(wbase + (size_t)stage_of(u) * LY::kStage + 3 * LY::kField);
        constexpr int kPairs = (R + 2) * kRecStride / 2;
        const double dt64 = s_dt;
#pragma unroll
        for (int e0 = 0; e0 < kPairs; e0 += 32) {
            const int e = e0 + lane;
            if (e < kPairs) {
                const int k = (2 * e) % kRecStride;  // slot of the pair's first element (kRecStride is even)
                double2 v = rec[e];
                if (k < kRowRec) {
                    v.x = scale_row_slot<false>(k, v.x, dt64);
                    v.y = scale_row_slot<false>(k + 1, v.y, dt64);
                } else if (k - kRowRec == kDiffDtSlot) {
                    v.x = dt64 * v.x;
                }
                rec[e] = v;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes before the stage's next tensor load
        __syncwarp();
    };
    auto load_f = [&](int jl) -> V2 {
        if (FM == 1) return ldg2((const T *)p.f2d + (size_t)min(jl, p.nrows - 1) * P + colA);
        if (FM == 2) {  // ic.py:126-133 of the reference, IEEE operation by operation (the tables are constant: read-only path)
            const int j = min(jl, p.nrows - 1);
            const double b = __ldg(p.tab + (size_t)T_FB * p.nrows + j), d = __ldg(p.tab + (size_t)T_FD * p.nrows + j);
            const double2 a = __ldg(reinterpret_cast<const double2 *>(p.fa + colA));
            const double fA = __dmul_rn(p.fs_scale, __dadd_rn(__dmul_rn(__dmul_rn(a.x, b), p.fs_sa), d));
            const double fB = __dmul_rn(p.fs_scale, __dadd_rn(__dmul_rn(__dmul_rn(a.y, b), p.fs_sa), d));
            return mk2((T)fA, (T)fB);
        }
        return mk2(T(0), T(0));
    };

    if (STREAM) {  // the first marched tile's records serve the first y-face
        wait_tile(kLead);
        scale_tile(kLead);
        set_rec_tile(kLead);
    }
    Cell<T> curA, curB;
    Face<T> y0A, y0B;
    {
        const V2 f0 = load_f(ja - 1), f1 = load_f(ja);
        const T *r0 = rec_of(ja - 1), *r1 = rec_of(ja);
        auto rs0 = [&](int k) { return rec_slot<PK>(r0, k); };
        const Cell<T> pA = make_cell<STRICT>(h0.x, u0.x, v0.x, rs0(RC_C), hgCell, f0.x);
        const Cell<T> pB = make_cell<STRICT>(h0.y, u0.y, v0.y, rs0(RC_C), hgCell, f0.y);
        curA = make_cell<STRICT>(h1.x, u1.x, v1.x, rec_slot<PK>(r1, RC_C), hgCell, f1.x);
        curB = make_cell<STRICT>(h1.y, u1.y, v1.y, rec_slot<PK>(r1, RC_C), hgCell, f1.y);
        auto fy = [&](const Cell<T> &a, const Cell<T> &b) -> T {
            if (!F2D) return rs0(RC_FKY);
            return STRICT ? O::mul(T(0.5), O::add(b.f, a.f)) : hdt * T(0.5) * (b.f + a.f);
        };
        y0A = y_face<STRICT, kGuard>(pA, curA, rs0(RC_CY1F), rs0(RC_CYF), rs0(RC_KY), fy(pA, curA), rs0(RC_CM), hdt, hgK, raK);
        y0B = y_face<STRICT, kGuard>(pB, curB, rs0(RC_CY1F), rs0(RC_CYF), rs0(RC_KY), fy(pB, curB), rs0(RC_CM), hdt, hgK, raK);
    }
    // PK: between rows the lane's cells and their lower y-faces are carried packed
    Cell2 cur2;
    Face2 y02;
    if constexpr (PK) {
        cur2 = pack_cells(curA, curB);
        y02 = pack_faces(y0A, y0B);
    }

    // rows at which a cell owes more than its own store: zero-gradient pole rows
    // (numpy.py:403-404) and the rows a neighbour band reads as its halo
    int plain_lo = ja, plain_hi = jb - 1;  // rows in [plain_lo, plain_hi] need nothing special
    if (p.pole_s) plain_lo = max(plain_lo, 2);
    if (p.pole_n) plain_hi = min(plain_hi, p.nrows - 3);
    if (p.world > 1) {
        if (p.south[0][0] != nullptr) plain_lo = max(plain_lo, p.jl_first + p.halo);
        if (p.north[0][0] != nullptr) plain_hi = min(plain_hi, p.jl_last - p.halo);
    }
    if (warp_wraps || p.force_rare) plain_hi = plain_lo - 1;  // strips that feed the periodic ghost columns: every row
    // resident kernel: only the pole rows (and degenerate grids) leave the fast body
    int res_lo = p.pole_s ? 2 : 0, res_hi = p.pole_n ? p.nrows - 3 : p.nrows;
    if (RESBANDS) {  // rows a neighbour band reads as its halo are stored there too: boundary body
        if (p.south[0][0] != nullptr) res_lo = max(res_lo, p.jl_first + p.halo);
        if (p.north[0][0] != nullptr) res_hi = min(res_hi, p.jl_last - p.halo);
    }
    if (p.force_rare || (warp_wraps && !res_wrap)) res_hi = res_lo - 1;

    // Resident kernel, ring-fed star, strips at the periodic seam: the two neighbours that lie
    // across the seam (numpy.py:376-378: column 1 reaches back to column M-1, column M+1 forward
    // to column 3) are written INTO the landed tile, in place of the out-of-range / padding
    // columns the box brought -- once per tile, latencies overlapped -- so that these strips run
    // the same branch-free body as all others.  The lane that writes a value is the lane that
    // reads it (its pair to the left / right), hence no synchronisation.
    const bool seam_b1 = vB && (colA + 1 == 1), seam_ae = vA && (colA == p.M + 1), seam_be = vB && (colA + 1 == p.M + 1);
    auto patch_tile = [&](int u) {
        if (STREAM || !(RES && SSTAR) || !warp_wraps || !res_wrap) return;
        if (seam_b1 || seam_ae || seam_be) {
            T *base = reinterpret_cast<T *>(wbase + (size_t)stage_of(u) * LY::kStage) + (2 * lane + 2);  // box element of (row 0, colA)
            const int src_col = seam_b1 ? p.M - 1 : 3;
            const int dst_off = seam_b1 ? -1 : (seam_ae ? 2 : 3);  // column -1, or M+3, relative to colA
            const int row0 = ja + 1 + (u - kLead) * R;
            T val[3][R];
#pragma unroll
            for (int f = 0; f < 3; ++f)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const size_t k = (size_t)min(max(row0 + r, 0), p.nrows - 1) * P + src_col;
                    val[f][r] = ld1<kLdStar>((f == 0 ? H : (f == 1 ? U : V)) + k);
                }
#pragma unroll
            for (int f = 0; f < 3; ++f)
#pragma unroll
                for (int r = 0; r < R; ++r) base[(f * R + r) * kBoxCols + dst_off] = val[f][r];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes before the stage's next tensor load
        }
    };

    // One latitude: cells (jl, colA), (jl, colA+1) from `cur`, the row above in (h2,u2,v2).
    // `hot` (a compile-time tag) = a live row in [plain_lo, plain_hi]: own stores only.
    size_t krow = (size_t)ja * P + colA;  // element index of (jl, colA), advanced row by row
    auto do_row = [&](auto hot_tag, int jl, V2 h2, V2 u2, V2 v2, bool live, const StarRows<V2> &sr) {
        constexpr bool HOT = decltype(hot_tag)::value;
        // resident kernel: ghost columns and pole rows are served inside the fast (unrolled) body
        constexpr bool EXTRA = RES && HOT && !STREAM;  // (the streaming kernel deals many tiles per block: its few seam strips take the boundary body)
        T hnA, unA, vnA, hnB, unB, vnB;
        Cell<T> nxtA, nxtB;
        Face<T> y1A, y1B, xAB, xBN;  // (of the x-faces the tail below reads the half-step depths only)
        V2 r_c_ac;
        if constexpr (PK && HOT) {
            // ---- fp32, hot rows: both cells of the lane through two-wide instructions (swb_scheme.cuh, P2) -------------
            // the record: four slots per 16-byte read; one register feeds both halves of a packed instruction
            const T *rec = rec_of(jl);
            const float4 *r4 = reinterpret_cast<const float4 *>(rec);
            const float4 q0 = r4[0], q1 = r4[1], q2 = r4[2], q3 = r4[3];
            static_assert(RC_KX == 2 && RC_CDX == 4 && RC_KY == 6 && RC_CM == 8 && RC_CYF == 10 && RC_CYU == 12 && RC_FD == 14, "record layout");
            const P2 kx = pdup(q0.z), fkx = pdup(q0.w), ncdx = pdup(q1.x), ncxu = pdup(q1.y), ky = pdup(q1.z), fky = pdup(q1.w);
            const P2 cm = pdup(q2.x), ncy1f = pdup(q2.y), ncyf = pdup(q2.z), ncy1u = pdup(q2.w), ncyu = pdup(q3.x), kc = pdup(q3.y), fd = pdup(q3.z);
            const P2 cn = pdup(rec[kRecStep + RC_C]);
            const P2 hgC2 = pdup(hgCell), hgK2 = pdup(hgK);
            const Cell2 nxt2 = make_cell2(pk(h2), pk(u2), pk(v2), cn, hgC2);
            const Face2 y12 = y_face2<kGuard>(cur2, nxt2, ncy1f, ncyf, ky, fky, cm, hgK2);
            // x-faces of this latitude: (A | B) in the low half, (B | next lane's A) in the high half
            Cell2 b2;
            b2.h = pk(phi(cur2.h), shfl_down1(plo(cur2.h)));    b2.u = pk(phi(cur2.u), shfl_down1(plo(cur2.u)));
            b2.hu = pk(phi(cur2.hu), shfl_down1(plo(cur2.hu))); b2.hv = pk(phi(cur2.hv), shfl_down1(plo(cur2.hv)));
            b2.Ux = pk(phi(cur2.Ux), shfl_down1(plo(cur2.Ux))); b2.Vx = pk(phi(cur2.Vx), shfl_down1(plo(cur2.Vx)));
            const Face2 xf = x_face2<kGuard>(cur2, b2, ncdx, kx, fkx, hgK2);
            Face2 x0;  // (previous lane's B | A) and (A | B)
            x0.huM = pk(shfl_up1(phi(xf.huM)), plo(xf.huM)); x0.hvM = pk(shfl_up1(phi(xf.hvM)), plo(xf.hvM));
            x0.q = pk(shfl_up1(phi(xf.q)), plo(xf.q));
            x0.F1 = pk(shfl_up1(phi(xf.F1)), plo(xf.F1));    x0.F2 = pk(shfl_up1(phi(xf.F2)), plo(xf.F2));
            P2 hn2, un2, vn2;
            update_cell2(cur2, x0, xf, y02, y12, ncxu, ncy1u, ncyu, kc, fd, hn2, un2, vn2);
            hnA = plo(hn2); hnB = phi(hn2); unA = plo(un2); unB = phi(un2); vnA = plo(vn2); vnB = phi(vn2);
            xAB.hM = plo(xf.hM); xBN.hM = phi(xf.hM); y1A.hM = plo(y12.hM); y1B.hM = phi(y12.hM);
            if (DIFF) {  // the star's centre values
                curA.h = plo(cur2.h); curB.h = phi(cur2.h); curA.u = plo(cur2.u); curB.u = phi(cur2.u); curA.v = plo(cur2.v); curB.v = phi(cur2.v);
            }
            cur2 = nxt2;
            y02 = y12;
        } else {
        if constexpr (PK) {  // a boundary row of a packed kernel: the scalar body on the unpacked state
            unpack_cells(cur2, curA, curB);
            unpack_faces(y02, y0A, y0B);
        }
        // the row record, read as 16-byte pairs (broadcast LDS.128)
        auto rec_pair = [&](int k) -> V2 {
            if constexpr (PK) return mk2(rec_slot<true>(rec_of(jl), k), rec_slot<true>(rec_of(jl), k + 1));
            else return reinterpret_cast<const V2 *>(rec_of(jl))[k / 2];
        };
        r_c_ac = rec_pair(RC_C);
        const V2 r_kx_fkx = rec_pair(RC_KX), r_cdx_cxu = rec_pair(RC_CDX), r_ky_fky = rec_pair(RC_KY);
        const V2 r_cm_cy1f = rec_pair(RC_CM), r_cyf_cy1u = rec_pair(RC_CYF), r_cyu_kc = rec_pair(RC_CYU), r_fd_dy1s = rec_pair(RC_FD);
        const T cn = rec_slot<PK>(rec_of(jl) + kRecStep, RC_C);
        const V2 f2 = load_f(jl + 1);
        nxtA = make_cell<STRICT>(h2.x, u2.x, v2.x, cn, hgCell, f2.x);
        nxtB = make_cell<STRICT>(h2.y, u2.y, v2.y, cn, hgCell, f2.y);

        // y-faces (jl | jl+1)
        T fyA = r_ky_fky.y, fyB = fyA;
        if (F2D) {
            fyA = STRICT ? O::mul(T(0.5), O::add(nxtA.f, curA.f)) : hdt * T(0.5) * (nxtA.f + curA.f);
            fyB = STRICT ? O::mul(T(0.5), O::add(nxtB.f, curB.f)) : hdt * T(0.5) * (nxtB.f + curB.f);
        }
        y1A = y_face<STRICT, kGuard>(curA, nxtA, r_cm_cy1f.y, r_cyf_cy1u.x, r_ky_fky.x, fyA, r_cm_cy1f.x, hdt, hgK, raK);
        y1B = y_face<STRICT, kGuard>(curB, nxtB, r_cm_cy1f.y, r_cyf_cy1u.x, r_ky_fky.x, fyB, r_cm_cy1f.x, hdt, hgK, raK);

        // x-faces of this latitude: (A | B) and (B | next lane's A)
        Cell<T> nA;
        nA.h = shfl_down1(curA.h);   nA.u = shfl_down1(curA.u);
        nA.hu = shfl_down1(curA.hu); nA.hv = shfl_down1(curA.hv);
        nA.Ux = shfl_down1(curA.Ux); nA.Vx = shfl_down1(curA.Vx);
        nA.f = F2D ? shfl_down1(curA.f) : T(0);
        T cdxAB, cdxBN, cxA, cxB, dxsA = 0, dxsB = 0;
        if (STRICT) {
            const T ac = r_c_ac.y;
            const T xm = O::mul(ac, ph_m), x0 = O::mul(ac, ph_0), x1 = O::mul(ac, ph_1), x2 = O::mul(ac, ph_2);
            const T dxm = O::sub(x0, xm), dxA = O::sub(x1, x0), dxB = O::sub(x2, x1);
            cdxAB = O::div(hdt, dxA);
            cdxBN = O::div(hdt, dxB);
            dxsA = O::add(dxm, dxA);
            dxsB = O::add(dxA, dxB);
            cxA = O::div(dt, O::mul(T(0.5), dxsA));
            cxB = O::div(dt, O::mul(T(0.5), dxsB));
        } else {
            cdxAB = cdxBN = r_cdx_cxu.x;
            cxA = cxB = r_cdx_cxu.y;
            if (HS) dxsA = dxsB = T(2.0) * r_c_ac.y * (T)p.dphi;
        }
        T fxAB = r_kx_fkx.y, fxBN = fxAB;
        if (F2D) {
            fxAB = STRICT ? O::mul(T(0.5), O::add(curB.f, curA.f)) : hdt * T(0.5) * (curB.f + curA.f);
            fxBN = STRICT ? O::mul(T(0.5), O::add(nA.f, curB.f)) : hdt * T(0.5) * (nA.f + curB.f);
        }
        xAB = x_face<STRICT, kGuard>(curA, curB, cdxAB, r_kx_fkx.x, fxAB, hdt, hgK, raK);
        xBN = x_face<STRICT, kGuard>(curB, nA, cdxBN, r_kx_fkx.x, fxBN, hdt, hgK, raK);
        Face<T> xL;  // (previous lane's B | A)
        xL.huM = shfl_up1(xBN.huM); xL.hvM = shfl_up1(xBN.hvM); xL.q = shfl_up1(xBN.q);
        xL.F1 = shfl_up1(xBN.F1);   xL.F2 = shfl_up1(xBN.F2);
        xL.hM = HS ? shfl_up1(xBN.hM) : T(0);

        // topography differences (numpy.py:310-317, 342-349)
        T hsxA = 0, hsxB = 0, hsyA = 0, hsyB = 0;
        if (HS) {
            const T *hr = (const T *)p.hs + (size_t)jl * P + colA;
            const T *hu_ = (const T *)p.hs + (size_t)min(jl + 1, p.nrows - 1) * P + colA;
            const T *hd_ = (const T *)p.hs + (size_t)max(jl - 1, 0) * P + colA;
            const T e0 = hr[colA > 0 ? -1 : 0], e1 = hr[0], e2 = hr[1], e3 = hr[2];
            hsxA = O::sub(e2, e0);
            hsxB = O::sub(e3, e1);
            hsyA = O::sub(hu_[0], hd_[0]);
            hsyB = O::sub(hu_[1], hd_[1]);
        }

        const T fcA = F2D ? (STRICT ? curA.f : dt * T(0.125) * curA.f) : r_fd_dy1s.x;
        const T fcB = F2D ? (STRICT ? curB.f : dt * T(0.125) * curB.f) : r_fd_dy1s.x;
        update_cell<STRICT, HS>(curA, xL, xAB, y0A, y1A, cxA, r_cyf_cy1u.y, r_cyu_kc.x, r_cyu_kc.y, fcA, dt, g_, raK, hsxA, hsyA,
                                dxsA, r_fd_dy1s.y, hnA, unA, vnA);
        update_cell<STRICT, HS>(curB, xAB, xBN, y0B, y1B, cxB, r_cyf_cy1u.y, r_cyu_kc.x, r_cyu_kc.y, fcB, dt, g_, raK, hsxB, hsyB,
                                dxsB, r_fd_dy1s.y, hnB, unB, vnB);
#ifdef SWB_COPY_ONLY  // dev probe (scripts/r3_copy_probe.sh): the kernel's memory traffic without its arithmetic -- the row above is stored as the result
        hnA = h2.x; hnB = h2.y; unA = u2.x; unB = u2.y; vnA = v2.x; vnB = v2.y;
#endif
        }  // (scalar body)

        if (DIFF) {  // numpy.py:541-544: Laplacian of the OLD fields; u, v (not hu, hv)
            const int iA = min(max(colA, 1), p.M + 1), iB = min(max(colA + 1, 1), p.M + 1);  // keep invalid lanes in bounds
            Star<T> shA, shB, suA, suB, svA, svB;
            // HOT: no pole row, no periodic column within reach: plain neighbours (EXTRA: checked per row)
            if (HOT && SSTAR) {
                // out of the ring: per field the pairs left and right of the lane's own and the
                // lane's pairs two and one latitudes below and two above (LDS.128, conflict-free)
                auto star = [&](int fld, T q0A, T q0B, V2 p1, Star<T> &sA, Star<T> &sB) {
                    const V2 L = sr.c0[fld * kFieldV - 1], Rr = sr.c0[fld * kFieldV + 1];
                    const V2 m2 = sr.m2[fld * kFieldV], m1 = sr.m1[fld * kFieldV], p2 = sr.p2[fld * kFieldV];
                    sA.q0 = q0A; sA.xmm = L.x; sA.xm = L.y; sA.xp = q0B;  sA.xpp = Rr.x;
                    sB.q0 = q0B; sB.xmm = L.y; sB.xm = q0A; sB.xp = Rr.x; sB.xpp = Rr.y;
                    sA.ymm = m2.x; sA.ym = m1.x; sA.yp = p1.x; sA.ypp = p2.x;
                    sB.ymm = m2.y; sB.ym = m1.y; sB.yp = p1.y; sB.ypp = p2.y;
                };
                star(0, curA.h, curB.h, h2, shA, shB);
                star(1, curA.u, curB.u, u2, suA, suB);
                star(2, curA.v, curB.v, v2, svA, svB);
            } else if (HOT && (!EXTRA || !warp_wraps)) {  // (direct loads: the seam strips keep the general star -- patching costs registers here)
                load_star_pair<T, kLdStar>(H, p.pitch, jl, colA, curA.h, curB.h, shA, shB);
                load_star_pair<T, kLdStar>(U, p.pitch, jl, colA, curA.u, curB.u, suA, suB);
                load_star_pair<T, kLdStar>(V, p.pitch, jl, colA, curA.v, curB.v, svA, svB);
            } else {
                shA = load_star<T, kLdStar>(H, p.pitch, p.nrows, p.M, jl, iA); shB = load_star<T, kLdStar>(H, p.pitch, p.nrows, p.M, jl, iB);
                suA = load_star<T, kLdStar>(U, p.pitch, p.nrows, p.M, jl, iA); suB = load_star<T, kLdStar>(U, p.pitch, p.nrows, p.M, jl, iB);
                svA = load_star<T, kLdStar>(V, p.pitch, p.nrows, p.M, jl, iA); svB = load_star<T, kLdStar>(V, p.pitch, p.nrows, p.M, jl, iB);
            }
            if (STRICT) {
                const T dnu = O::mul(dt, (T)p.nu);
                const W3 wy = y_weights(p.tab, p.nrows, jl);
                const W3 wxA = x_weights_strict(p.phi, r_c_ac.y, iA, p.M), wxB = x_weights_strict(p.phi, r_c_ac.y, iB, p.M);
                auto lap = [&](const Star<T> &s, const W3 &wx) -> T {
                    return O::add(second_diff_strict(wx, s.xmm, s.xm, s.q0, s.xp, s.xpp),
                                  second_diff_strict(wy, s.ymm, s.ym, s.q0, s.yp, s.ypp));
                };
                hnA = O::add(hnA, O::mul(dnu, lap(shA, wxA))); hnB = O::add(hnB, O::mul(dnu, lap(shB, wxB)));
                unA = O::add(unA, O::mul(dnu, lap(suA, wxA))); unB = O::add(unB, O::mul(dnu, lap(suB, wxB)));
                vnA = O::add(vnA, O::mul(dnu, lap(svA, wxA))); vnB = O::add(vnB, O::mul(dnu, lap(svB, wxB)));
            } else {
                const T *rd = STREAM ? rec_of(jl) + kRowRec : rowd + (size_t)(jl - (ja - 1)) * kRowDiff;
                const T dnu = rd[kDiffDtSlot];
                hnA = fma(dnu, laplacian_fast(shA, rd), hnA); hnB = fma(dnu, laplacian_fast(shB, rd), hnB);
                unA = fma(dnu, laplacian_fast(suA, rd), unA); unB = fma(dnu, laplacian_fast(suB, rd), unB);
                vnA = fma(dnu, laplacian_fast(svA, rd), vnA); vnB = fma(dnu, laplacian_fast(svB, rd), vnB);
            }
        }

        // CFL maxima of the new state: max(|q-c|, |q|, |q+c|) == |q| + c for c >= 0
        if (CFLF && HOT) {
            // estimate only; tiles in which an estimate reaches the trigger threshold are
            // re-evaluated exactly after their last row (see CflFilter)
            // (Round 2: the estimate costs 4.9 % of the headline step -- a build without it, -DSWB_NO_CFL_EST, scripts/r3_cflest_probe.sh.
            // Moving it to fp32 -- three cvt.rz.f32.f64, two FMUL, MUFU.RSQ, two FADD per cell instead of two DMUL, MUFU.RSQ64H, two
            // DADD; exact all the same, 240 instead of 254 registers -- was measured 0.5 % SLOWER on the same box (1.169 against
            // 1.163 ms): the conversions run on the XU pipe at one per eight cycles, next to the MUFU seeds of the reciprocals.)
#ifndef SWB_NO_CFL_EST
            const T xA = g_ * fabs(hnA), xB = g_ * fabs(hnB);
            const T cA = xA * (T)rsqrt_seed((double)xA), cB = xB * (T)rsqrt_seed((double)xB);
            // (bitwise | on purpose: no short-circuit branches inside the hot tile)
            trig = trig | ((hi_word(fabs(unA) + cA) & 0x7fffffff) >= trig_x) | ((hi_word(fabs(vnA) + cA) & 0x7fffffff) >= trig_y) |
                   ((hi_word(fabs(unB) + cB) & 0x7fffffff) >= trig_x) | ((hi_word(fabs(vnB) + cB) & 0x7fffffff) >= trig_y);
#endif
        } else if (HOT || live) {
            bits_t<T> kxA, kyA, kxB, kyB;
            if (STRICT) {
                const T csA = (T)__dsqrt_rn(O::mul(g_, fabs(hnA))), csB = (T)__dsqrt_rn(O::mul(g_, fabs(hnB)));
                kxA = nonneg_bits(O::add(fabs(unA), csA)); kyA = nonneg_bits(O::add(fabs(vnA), csA));
                kxB = nonneg_bits(O::add(fabs(unB), csB)); kyB = nonneg_bits(O::add(fabs(vnB), csB));
            } else {
                cfl_keys_fast(hnA, unA, vnA, g_, kxA, kyA);
                cfl_keys_fast(hnB, unB, vnB, g_, kxB, kyB);
            }
            mxA = umax64(mxA, kxA); myA = umax64(myA, kyA);
            mxB = umax64(mxB, kxB); myB = umax64(myB, kyB);
        }
        if (!kGuard && (HOT || live)) {
            negA |= hi_word(xAB.hM) | hi_word(y1A.hM) | hi_word(hnA);
            negB |= hi_word(xBN.hM) | hi_word(y1B.hM) | hi_word(hnB);
        }

        // stores: the pair is 16-byte aligned; the strip's first and last column are not ours
        if (HOT || live) {
            st_pair(Hn + krow, hnA, hnB, st_both, st_onlyA, st_onlyB);
            st_pair(Un + krow, unA, unB, st_both, st_onlyA, st_onlyB);
            st_pair(Vn + krow, vnA, vnB, st_both, st_onlyA, st_onlyB);
        }
        if (EXTRA) {  // periodic ghost columns (numpy.py:402), branch-free; pole rows never come here
            st_if(Hn + krow + goA, hnA, gpA); st_if(Un + krow + goA, unA, gpA); st_if(Vn + krow + goA, vnA, gpA);
            st_if(Hn + krow + goB, hnB, gpB); st_if(Un + krow + goB, unB, gpB); st_if(Vn + krow + goB, vnB, gpB);
        }
        if (!HOT && live) {
            // boundary conditions (numpy.py:402-404): periodic ghost columns, zero-gradient
            // pole rows (which also copy the ghost columns), halo rows of the neighbour bands
            const bool south = p.pole_s && (jl == 1), north = p.pole_n && (jl == p.nrows - 2);
            if (warp_wraps || south || north) {
                auto put = [&](size_t kk, T hh, T uu, T vv) { Hn[kk] = hh; Un[kk] = uu; Vn[kk] = vv; };
                auto dup = [&](bool valid, bool wraps, int col, T hh, T uu, T vv) {
                    if (!valid) return;
                    // a wrapping column is mirrored at column 0 (col == M) and / or M+2 (col == 2)
                    const bool w0 = wraps && (col == p.M), w2 = wraps && (col == 2);
                    auto row_put = [&](int row, bool centre) {
                        if (centre) put((size_t)row * P + col, hh, uu, vv);
                        if (w0) put((size_t)row * P, hh, uu, vv);
                        if (w2) put((size_t)row * P + p.M + 2, hh, uu, vv);
                    };
                    row_put(jl, false);
                    if (south) row_put(jl - 1, true);
                    if (north) row_put(jl + 1, true);
                };
                dup(vA, wrapA, colA, hnA, unA, vnA);
                dup(vB, wrapB, colA + 1, hnB, unB, vnB);
            }
            if (p.world > 1) {
                // rows a neighbour band reads as its halo: store them into its arrays as well
                const int ds = jl - p.jl_first, dn = p.jl_last - jl;
                auto peer_put = [&](T *const *nb, int prow, int ppitch) {
                    const size_t kb = (size_t)prow * ppitch + colA;
                    auto one = [&](int col_off, bool valid, bool wraps, T hh, T uu, T vv) {
                        if (!valid) return;
                        nb[0][kb + col_off] = hh; nb[1][kb + col_off] = uu; nb[2][kb + col_off] = vv;
                        const int col = colA + col_off;
                        if (wraps && col == p.M) {
                            const size_t kg = (size_t)prow * ppitch;
                            nb[0][kg] = hh; nb[1][kg] = uu; nb[2][kg] = vv;
                        }
                        if (wraps && col == 2) {
                            const size_t kg = (size_t)prow * ppitch + p.M + 2;
                            nb[0][kg] = hh; nb[1][kg] = uu; nb[2][kg] = vv;
                        }
                    };
                    one(0, vA, wrapA, hnA, unA, vnA);
                    one(1, vB, wrapB, hnB, unB, vnB);
                };
                if (ds < p.halo && p.south[par ^ 1][0] != nullptr) peer_put(reinterpret_cast<T *const *>(p.south[par ^ 1]), p.south_row + ds, p.south_pitch);
                if (dn < p.halo && p.north[par ^ 1][0] != nullptr) peer_put(reinterpret_cast<T *const *>(p.north[par ^ 1]), p.north_row - dn, p.north_pitch);
            }
        }
        krow += P;
        if constexpr (PK) {
            if constexpr (!HOT) {
                cur2 = pack_cells(nxtA, nxtB);
                y02 = pack_faces(y1A, y1B);
            }
        } else {
            curA = nxtA; curB = nxtB;
            y0A = y1A;   y0B = y1B;
        }
    };
    using HotTag = std::integral_constant<bool, true>;
    using RareTag = std::integral_constant<bool, false>;

    // ---- the march ------------------------------------------------------------------
    SWB_STAMP(3);
    const StarRows<V2> no_star{nullptr, nullptr, nullptr, nullptr};
    if (TMA) {
        if (SSTAR) {
            wait_tile(1);
            patch_tile(0);
            patch_tile(1);
        }
        for (int t = 0; t < ntiles; ++t) {
            const int u = t + kLead;
            if (!SSTAR) wait_tile(u);
            if (STREAM) {
                if (!SSTAR && t > 0) scale_tile(u);
                set_rec_tile(u);
            }
            const V2 *tile = tile_of(u);
            const V2 *tprev = SSTAR ? tile_of(u - 1) : tile, *tnext = SSTAR ? tile_of(u + 1) : tile;
            // row i of the marched tile, reaching into its neighbours in the ring (i is a constant after unrolling)
            auto ring_row = [&](int i) { return i < 0 ? tprev + (i + R) * kRowV : (i >= R ? tnext + (i - R) * kRowV : tile + i * kRowV); };
            // SSTAR ring protocol: tile u-1 is dead once row R-2 of tile u is done (the last row reaches
            // into tiles u and u+1 only), so its stage is refilled THERE -- and tile u+1, requested at the
            // same point of the previous tile, is awaited there: a whole tile of slack with three stages.
            auto turn_ring = [&]() {
                __syncwarp();
                if (lane == 0 && u + S - 1 < nload) issue_load(u + S - 1);
                if (u + 1 < nload) {
                    wait_tile(u + 1);
                    patch_tile(u + 1);
                    if (STREAM && u + 1 < kLead + ntiles) scale_tile(u + 1);
                }
            };
            const int j0 = ja + t * R;
            // warp-uniform; the resident kernel's fast body serves every complete tile
            const bool hot_tile = (RES && !STREAM) ? (j0 + R <= jb && j0 >= res_lo && j0 + R - 1 <= res_hi) : (j0 >= plain_lo && j0 + R - 1 <= plain_hi);
            if (hot_tile) {
                // (With the ring-fed star the compiler leaves this loop rolled -- ncu, round 1: the `rr == R-1` test and the
                // ring_row selects execute per row, 20 % of that kernel's instructions are IMAD address arithmetic.  Spelling the
                // four rows out (compile-time row index) was tried in round 2: the spills grow from 16 / 52 to 144 / 212 bytes
                // and every diffusion grid gets SLOWER -- 7200 x 3600 454 -> 470 us, 1440 x 720 resident 25.9 -> 27.9 us.)
                // (Unrolling by two instead of four: 234 instead of 248 registers for the headline kernel, and slower --
                // 16384 x 8192 1333 -> 1356 us, 7200 x 3600 274 -> 281 us, fp32 740 -> 762 us per step.)
                // (A compiler fence in front of the star's shared-memory loads changes nothing: 255 registers either way.)
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    if (SSTAR && rr == R - 1) turn_ring();
                    // the streamed row rr is latitude jl+1: jl-2, jl-1, jl, jl+2 are rows rr-3, rr-2, rr-1, rr+1
                    const StarRows<V2> sr{ring_row(rr - 3), ring_row(rr - 2), ring_row(rr - 1), ring_row(rr + 1)};
                    do_row(HotTag(), j0 + rr, tile[rr * kRowV], tile[kFieldV + rr * kRowV], tile[2 * kFieldV + rr * kRowV], true, SSTAR ? sr : no_star);
                }
                if (CFLF) {
                    if (__any_sync(0xffffffffu, trig)) {  // rare: exact keys of the tile's cells, from the values just stored
#pragma unroll 1
                        for (int rr = 0; rr < R; ++rr) {
                            const size_t k = (size_t)(j0 + rr) * P + colA;
                            const V2 hh = ldg2(Hn + k), uu = ldg2(Un + k), vv = ldg2(Vn + k);
                            bits_t<T> kxA, kyA, kxB, kyB;
                            cfl_keys_fast(hh.x, uu.x, vv.x, g_, kxA, kyA);
                            cfl_keys_fast(hh.y, uu.y, vv.y, g_, kxB, kyB);
                            mxA = umax64(mxA, kxA); myA = umax64(myA, kyA);
                            mxB = umax64(mxB, kxB); myB = umax64(myB, kyB);
                        }
                    }
                    trig = false;
                }
            } else if (RES && !STREAM && RESMIX) {
                // A tile that is not fast as a whole -- the march's last, partial tile, or one that holds a pole row or a halo row of
                // a neighbour band: the rows that exist, each through the body it needs (a rolled loop).  The blocks that own the
                // pole rows are the ones every step of a small grid waits for: with whole tiles through the boundary body their
                // march is 20 % longer than the mean (profiles/r1_small_grids.md).  1440 x 720 with diffusion (15-row marches)
                // 26.3 -> 24.7 us per step; NOT for longer marches and not without diffusion, where the extra loop costs the
                // kernel's other paths more than it saves (2000 x 1000 + diffusion, 32 rows: 47.2 -> 48.4 us; 1440 x 720: 16.3 -> 18.4).
#pragma unroll 1
                for (int rr = 0; rr < min(R, jb - j0); ++rr) {
                    const int jl = j0 + rr;
                    if (jl >= res_lo && jl <= res_hi) {
                        const StarRows<V2> sr{ring_row(rr - 3), ring_row(rr - 2), ring_row(rr - 1), ring_row(rr + 1)};
                        do_row(HotTag(), jl, tile[rr * kRowV], tile[kFieldV + rr * kRowV], tile[2 * kFieldV + rr * kRowV], true, SSTAR ? sr : no_star);
                    } else {
                        do_row(RareTag(), jl, tile[rr * kRowV], tile[kFieldV + rr * kRowV], tile[2 * kFieldV + rr * kRowV], true, no_star);
                    }
                }
            } else if (RES && !STREAM && j0 + R > jb && j0 >= res_lo && jb - 1 <= res_hi) {
                // resident kernel, the march's last tile when its length is not a multiple of R: the fast body for the rows
                // that exist (a rolled loop: at most R-1 rows).  Lets the host pick ANY marching length, i.e. the one that
                // fills the GPU's block slots evenly (1440 x 720: 15-row marches in 288 blocks instead of 16 rows in 270).
#pragma unroll 1
                for (int rr = 0; rr < jb - j0; ++rr) {
                    const StarRows<V2> sr{ring_row(rr - 3), ring_row(rr - 2), ring_row(rr - 1), ring_row(rr + 1)};
                    do_row(HotTag(), j0 + rr, tile[rr * kRowV], tile[kFieldV + rr * kRowV], tile[2 * kFieldV + rr * kRowV], true, SSTAR ? sr : no_star);
                }
            } else {
#pragma unroll 1
                for (int rr = 0; rr < R; ++rr)
                    do_row(RareTag(), j0 + rr, tile[rr * kRowV], tile[kFieldV + rr * kRowV], tile[2 * kFieldV + rr * kRowV],
                           j0 + rr < jb, no_star);
            }
            if (SSTAR) {
                if (!hot_tile) turn_ring();
            } else {
                __syncwarp();  // every lane is done with the stage that is refilled next
                if (lane == 0 && u + S - 1 < nload) issue_load(u + S - 1);
            }
        }
    } else {
#pragma unroll 1
        for (int jl = ja; jl < jb; ++jl) {
            const size_t k = (size_t)min(jl + 1, p.nrows - 1) * P + colA;
            if (RES ? (jl >= res_lo && jl <= res_hi) : (jl >= plain_lo && jl <= plain_hi)) do_row(HotTag(), jl, ld2<kLdRow>(H + k), ld2<kLdRow>(U + k), ld2<kLdRow>(V + k), true, no_star);
            else do_row(RareTag(), jl, ld2<kLdRow>(H + k), ld2<kLdRow>(U + k), ld2<kLdRow>(V + k), true, no_star);
        }
    }

    // ---- epilogue: publish the CFL maxima and the depth watch -------------------------
    SWB_STAMP(4);
    if (!p.fixed) {
        bits_t<T> mx = umax64(vA ? mxA : bits_t<T>(0), vB ? mxB : bits_t<T>(0));
        bits_t<T> my = umax64(vA ? myA : bits_t<T>(0), vB ? myB : bits_t<T>(0));
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            mx = umax64(mx, __shfl_xor_sync(0xffffffffu, mx, s));
            my = umax64(my, __shfl_xor_sync(0xffffffffu, my, s));
        }
        if (lane == 0) {
            Ctrl *const new_slot = p.ctrl + ((launch + 1) % 3);
            if (RES && CFLF) {  // stream_kernel: per block first, with the found-mask of the filter
                const unsigned long long wx = widen_bits(mx), wy = widen_bits(my);
                if (wx != 0ULL) atomicMax(&s_max[0], wx);
                if (wy != 0ULL) atomicMax(&s_max[1], wy);
                const unsigned long long found = (wx >= s_filter.valid_x && wx != 0ULL ? 1ULL : 0ULL) | (wy >= s_filter.valid_y && wy != 0ULL ? 2ULL : 0ULL);
                if (found) atomicOr(&s_max[2], found);
            } else if (RES) {  // per block first; the resident kernels forward the block's maxima behind their barrier's __syncthreads
                atomicMax(&s_max[0], widen_bits(mx));
                atomicMax(&s_max[1], widen_bits(my));
            } else if (CFLF) {
                // most warps hold nothing: their cells stayed below the trigger threshold
                const unsigned long long wx = widen_bits(mx), wy = widen_bits(my);
                if (wx != 0ULL) atomicMax(&new_slot->ex_bits, wx);
                if (wy != 0ULL) atomicMax(&new_slot->ey_bits, wy);
                const int found = (wx >= s_filter.valid_x && wx != 0ULL ? 1 : 0) | (wy >= s_filter.valid_y && wy != 0ULL ? 2 : 0);
                if (found) atomicOr(&new_slot->reserved, found);
            } else {
                atomicMax(&new_slot->ex_bits, widen_bits(mx));
                atomicMax(&new_slot->ey_bits, widen_bits(my));
            }
        }
    }
    if (!kGuard) {
        const bool bad = (vA && negA < 0) || (vB && negB < 0);
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicMin(p.err, -6);  // SWB_E_DEPTH
    }
    return true;
}

}  // namespace swb
