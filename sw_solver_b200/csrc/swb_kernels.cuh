// swb_kernels.cuh -- device code of libswb200.so (sm_100a).
//
// The hot path is one device function, `march_step`, that replaces the body of the loop at
// src/sw_solver/numpy.py:496-549 of the reference.  Three kernels call it: `march_kernel` (one
// launch per time step: any grid, any band decomposition; the launches of a step are chained by
// programmatic dependent launch and the last block rows march fewer latitudes -- the guided tail),
// `resident_kernel` (grids whose blocks all fit the GPU at once: the whole batch of steps in one
// cooperative launch, a grid barrier per step) and `stream_kernel` (the resident loop for grids
// beyond one wave of blocks: several tiles per block per step; opt-in).  Per step:
//   * prologue: dt from the device-resident CFL maxima + final-time clamp (numpy.py:524-532)
//   * Lax-Wendroff half-step face values and full-step update      (numpy.py:185-354)
//   * optional Laplacian diffusion increment on the OLD fields       (numpy.py:541-544, 66-133)
//   * boundary conditions written in place of the ghost cells        (numpy.py:402-404)
//   * epilogue: CFL maxima of the NEW state for the next step        (numpy.py:503-521)
//
// Data layout (HBM): LATITUDE-MAJOR, state[jl][i] with i = longitude 0..M+2 contiguous and
// `pitch` elements between consecutive latitudes (the transpose of the reference's array;
// swb_import_state / swb_export_state convert).  A latitude band's halo rows are therefore
// contiguous, and every metric term (a function of latitude only) is uniform across a warp.
//
// Work decomposition: a warp owns a strip of 64 consecutive longitudes -- two cells per
// lane -- and MARCHES along latitude.  The y-face (j | j+1) computed at one iteration is
// the lower face of the next one and stays in registers; of the two x-faces a lane
// computes, one lies between its own two cells and the other is shared with the next
// lane through warp shuffles.  Columns 1..62 of a strip are updated (strips overlap by
// two columns), so every face is computed once per warp.  Per-row metric/Coriolis/dt
// factors are built once per block in shared memory and broadcast; nothing 2-D but
// h, u, v is read from HBM.
//
// Files: swb_types.cuh (constants, parameter blocks) <- swb_scheme.cuh (the arithmetic of the
// scheme) <- swb_step.cuh (`march_step`: one time step of a block's tile) <- this file (kernels).
#pragma once
#include "swb_step.cuh"

namespace swb {

#ifdef SWB_BAND_TIMING
__device__ __forceinline__ unsigned long long swb_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define SWB_TLOG(p, k) do { if ((p).tlog != nullptr) (p).tlog[((p).stamp % kTlogCap) * kTlogRow + (k)] = swb_globaltimer(); } while (0)
#define SWB_TLOG_MAX(p, k) do { if ((p).tlog != nullptr) atomicMax(&(p).tlog[((p).stamp % kTlogCap) * kTlogRow + (k)], swb_globaltimer()); } while (0)
#else
#define SWB_TLOG(p, k) do { } while (0)
#define SWB_TLOG_MAX(p, k) do { } while (0)
#endif

// Programmatic dependent launch: the kernels of the per-step chain (gather -> march -> cfl_fixup -> publish) are
// launched with programmatic stream serialisation, so the next kernel's blocks are scheduled while the previous
// kernel drains; `pdl_wait` holds them until the previous grid has completed and flushed its memory operations.
// Every kernel of the chain triggers its dependents at once and waits before it touches global memory.  (Both
// are no-ops in a kernel launched without the attribute.)
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// One launch per time step (any grid size, any band decomposition).
template <typename T, bool STRICT, bool DIFF, int FM, bool HS, bool TMA, int R, int S, bool GUARD, int MINB, int W = kWarpsPerBlock,
          bool CFLF = false>
__global__ void __launch_bounds__(W * 32, MINB)
march_kernel(const __grid_constant__ Params p, const __grid_constant__ TmaMaps tm)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double s_dt;
    __shared__ int s_flags;
    __shared__ CflFilter s_filter;
    int tbase = 0, par = -1;
    par = p.par_guess;  // (used by the fp64 TMA kernels only, see march_step)
    pdl_enter();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) SWB_TLOG(p, 2);
    const int by = (int)blockIdx.y;
    int ja = p.jl_first + by * p.chunk, len = p.chunk;
    if (p.n_big > 0 && by >= p.n_big) {  // guided tail: shorter marches in the block rows dispatched last
        ja = p.jl_first + p.n_big * p.chunk + (by - p.n_big) * p.chunk_small;
        len = p.chunk_small;
    }
    const Tile tl{(int)blockIdx.x * W, ja, min(ja + len, p.jl_last + 1), blockIdx.x == 0 && blockIdx.y == 0, true};
    march_step<T, STRICT, DIFF, FM, HS, TMA, R, S, GUARD, W, CFLF, false>(p, tm, p.launch, true, tbase, par, smem_raw, s_dt, s_flags, s_filter,
                                                                          nullptr, tl);
#ifdef SWB_BAND_TIMING
    __syncthreads();
    if (threadIdx.x == 0) SWB_TLOG_MAX(p, 3);
#endif
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Latitude bands inside the resident kernel: the exchange of publish_kernel / gather_kernel (below),
// run by one warp of block (0,0) between two steps.  Same slots, same stamps: a band that runs
// resident and a neighbour that launches per step understand each other.
//   band_publish: lane r deposits this band's CFL maxima of the state after step `launch` in rank
//   r's exchange block and releases the stamp (system scope: the halo rows this band stored into
//   its neighbours' arrays are visible before it).
__device__ __forceinline__ void band_publish(const Params &p, int launch, unsigned long long stamp, unsigned long long ex,
                                             unsigned long long ey, int lane)
{
    if (lane < p.world) {
        // (the release is cumulative: it covers the maxima stored just before it and, through the
        // arrival counter this warp has acquired, the halo rows every block of the band stored)
        XSlot *dst = &p.xpeer[lane]->slot[(launch + 1) % 3][p.rank];
        dst->ex_bits = ex;
        dst->ey_bits = ey;
        st_release_sys(&dst->stamp, stamp + 1);
    }
}
//   band_gather: lane r waits (on this GPU's own memory) for rank r's stamp of step `launch`; lane 0
//   writes the maxima over all bands into the slot the step reads.  False after a time-out.
__device__ __forceinline__ bool band_gather(const Params &p, int launch, unsigned long long stamp, int lane)
{
    unsigned long long ex = 0ULL, ey = 0ULL;
    bool ok = true;
    if (lane < p.world) {
        const XSlot *xs = &p.xlocal->slot[launch % 3][lane];
        const long long t0 = clock64();
        while (ld_acquire_sys(&xs->stamp) != stamp) {
            if (clock64() - t0 > p.spin_limit) {
                atomicMin(p.err, -4);            // SWB_E_TIMEOUT
                atomicExch(p.done_ctr + 2, 1u);  // ... and the loop ends here (step_prologue)
                ok = false;
                break;
            }
        }
        ex = *reinterpret_cast<const volatile unsigned long long *>(&xs->ex_bits);
        ey = *reinterpret_cast<const volatile unsigned long long *>(&xs->ey_bits);
    }
    // every lane's acquire happens before anything lane 0 releases afterwards (the shuffles below
    // exchange values but order no memory accesses)
    __syncwarp();
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        ex = umax64(ex, __shfl_xor_sync(0xffffffffu, ex, s));
        ey = umax64(ey, __shfl_xor_sync(0xffffffffu, ey, s));
    }
    if (lane == 0) {
        Ctrl *cur_slot = p.ctrl + (launch % 3);
        cur_slot->ex_bits = ex;
        cur_slot->ey_bits = ey;
    }
    return __all_sync(0xffffffffu, ok);
}

// Grids that fit the GPU as ONE wave of blocks (360 x 180, 1440 x 720 -- the
// state lives in the L2, a step takes microseconds and a launch per step would cost as much
// again): the whole time loop of numpy.py:496-549 in one cooperative launch.  Every block
// keeps its tile; a step ends in a grid barrier (one arrival atomic per block, release /
// acquire at gpu scope), behind which the next step reads the CFL maxima -- accumulated by
// the same atomics as in march_kernel -- and its neighbours' new rows and columns straight
// from the L2.  Same step body, same arithmetic: bit-identical to a launch per step.
template <typename T, bool STRICT, bool DIFF, bool TMA, int R, int S, bool GUARD, int MINB, bool BANDS = false, int W = kWarpsPerBlock, int FM = 0,
          bool MIX = false>
__global__ void __launch_bounds__(W * 32, MINB)
resident_kernel(const __grid_constant__ Params p, const __grid_constant__ TmaMaps tm)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double s_dt;
    __shared__ int s_flags;
    __shared__ CflFilter s_filter;
    __shared__ unsigned long long s_max[2];
    const unsigned int nblocks = gridDim.x * gridDim.y;
    // latitude bands (world > 1): warp 0 of block (0,0) runs the exchange with the other GPUs; the
    // blocks are admitted to a step through `go` (done_ctr[1]) once the maxima of ALL bands are in
    constexpr bool bands = BANDS;  // (a template axis: the one-band kernels keep their register allocation)
    const bool xwarp = bands && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 32;
    unsigned int *const go = p.done_ctr + 1;
    const int lane = threadIdx.x & 31;
    int tbase = 0, par = -1;
    int s = 0;
    for (; s < p.nsteps; ++s) {
#ifdef SWB_RES_TIMING
        constexpr bool RES = true;
        const int launch = p.launch + s;
        SWB_STAMP(0);
        SWB_STAMP(8);
#endif
        if (bands) {
            if (xwarp) {
                SWB_STAMP(13);
                const bool ok = band_gather(p, p.launch + s, p.stamp + (unsigned long long)s, lane);
                SWB_STAMP(14);
                if (lane == 0) {
                    __threadfence();
                    st_release_gpu(go, ok ? (unsigned int)(s + 1) : 0x7fffffffu);  // (after a time-out: nobody waits any more)
                }
                SWB_STAMP(15);
            }
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                while ((int)(ld_acquire_gpu(go) - (unsigned int)(s + 1)) < 0) {
                    if (clock64() - t0 > 2 * p.spin_limit) {
                        atomicMin(p.err, -4);
                        atomicExch(p.done_ctr + 2, 1u);
                        break;
                    }
                }
            }
            __syncthreads();
        }
        const int ja = p.jl_first + blockIdx.y * p.chunk;
        const Tile tl{(int)blockIdx.x * W, ja, min(ja + p.chunk, p.jl_last + 1), blockIdx.x == 0 && blockIdx.y == 0, true};
        const bool active = march_step<T, STRICT, DIFF, FM, false, TMA, R, S, GUARD, W, false, true, BANDS, false, MIX>(p, tm, p.launch + s, s == 0, tbase, par, smem_raw,
                                                                                                         s_dt, s_flags, s_filter, s_max, tl);
        if (!active) break;
        // Grid barrier: all stores and atomics of step s happen before any load of step s+1.
        // (TMA: every thread first orders its generic-proxy stores before the async-proxy reads
        // the next step's tensor loads are.)
        if (TMA) asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncthreads();
        SWB_STAMP(5);
        Ctrl *const new_slot = p.ctrl + ((p.launch + s + 1) % 3);
        const unsigned int target = (unsigned int)(s + 1) * nblocks;
        if (threadIdx.x == 0) {
            atomicMax(&new_slot->ex_bits, s_max[0]);
            atomicMax(&new_slot->ey_bits, s_max[1]);
            if (bands || s + 1 < p.nsteps) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.done_ctr) : "memory");
            SWB_STAMP(6);
            if ((!bands && s + 1 < p.nsteps) || xwarp) {  // one band: every block waits here; bands: the exchange warp only
                const long long t0 = clock64();
                while ((int)(ld_acquire_gpu(p.done_ctr) - target) < 0) {
                    if (clock64() - t0 > p.spin_limit) {  // never in a healthy run: turn a lost block into an error, not a hang
                        atomicMin(p.err, -4);             // SWB_E_TIMEOUT
                        atomicExch(p.done_ctr + 2, 1u);
                        s_flags = -1;
                        break;
                    }
                }
                SWB_STAMP(7);
            }
        }
        if (xwarp) {  // every block of this band is done with step s: tell the other bands
            unsigned long long ex = 0ULL, ey = 0ULL;
            if (lane == 0) {
                ex = ld1<LD_CG>(&new_slot->ex_bits);
                ey = ld1<LD_CG>(&new_slot->ey_bits);
            }
            ex = __shfl_sync(0xffffffffu, ex, 0);
            ey = __shfl_sync(0xffffffffu, ey, 0);
            band_publish(p, p.launch + s, p.stamp + (unsigned long long)s, ex, ey, lane);
            SWB_STAMP(12);
        }
        __syncthreads();
        if (s_flags == -1) break;
    }
    // Bands: a launch that ended early (time >= final_time at step s) still owes the other bands the
    // stamps of its remaining steps -- they are no-ops everywhere, with the maxima carried forward
    if (xwarp && s < p.nsteps && s_flags != -1) {
        for (int k = s; k < p.nsteps; ++k) {
            const Ctrl *carry = p.ctrl + ((p.launch + k + 1) % 3);
            unsigned long long ex = 0ULL, ey = 0ULL;
            if (lane == 0) {
                ex = ld1<LD_CG>(&carry->ex_bits);
                ey = ld1<LD_CG>(&carry->ey_bits);
            }
            ex = __shfl_sync(0xffffffffu, ex, 0);
            ey = __shfl_sync(0xffffffffu, ey, 0);
            band_publish(p, p.launch + k, p.stamp + (unsigned long long)k, ex, ey, lane);
            if (k + 1 < p.nsteps && !band_gather(p, p.launch + k + 1, p.stamp + (unsigned long long)(k + 1), lane)) break;
        }
    }
}

// ------------------------------------------------------------------------------------
// stream_kernel: the resident loop for grids BEYOND one wave of blocks (7200 x 3600, 16384 x 8192,
// latitude bands of any height).  One cooperative launch runs a whole batch of time steps; per step
// every block marches several tiles -- map 1: tiles of `stream_unit` rows dealt round-robin in the
// row-major (tile row, block column) order, so that the blocks sweep the grid together like the
// blocks of a launch per step (DRAM-page / TLB locality); map 0: one contiguous range of the
// column-major space per block (equal work whatever the grid, but 296 scattered streams) -- with its
// row records arriving through the TMA ring next to the tiles (march_step<..., STREAM>): no launch per
// step, no block-shared records, no __syncthreads inside a step but the one behind the prologue.
// CFLF: the CFL filter of march_kernel (estimate per cell, exact evaluation near the previous
// maximum) with its fix-up INSIDE the launch: behind the grid barrier every block reads the
// found-mask; in the rare step in which a maximum fell by more than 2^-10 all blocks re-evaluate
// their tiles exactly (one more barrier).  BANDS: the exchange with the other GPUs as in
// resident_kernel, by warp 0 of block 0; admission tokens on `go`: 2s+2 = step s of this launch may
// start, 2s+1 = first recompute the maxima of the state step s reads (fix-up round).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool stream_next_piece(const Params &p, long long &it, const long long it_end, const long long lin1, int &col, int &r0,
                                                  int &len)
{
    if (it >= it_end) return false;
    if (p.stream_map != 0) {  // tile `it` of the row-major (tile row, block column) order
        col = (int)(it % p.stream_gx);
        r0 = (int)(it / p.stream_gx) * p.stream_unit;
        len = min(p.stream_unit, p.stream_rows - r0);
        it += gridDim.x;
    } else {                  // the part of this block's range that lies in block column `col`
        col = (int)(it / p.stream_rows);
        r0 = (int)(it % p.stream_rows);
        len = (int)min((long long)(p.stream_rows - r0), lin1 - it);
        it += len;
    }
    return true;
}

template <bool DIFF, int S, bool GUARD, bool BANDS, bool CFLF>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2)
stream_kernel(const __grid_constant__ Params p, const __grid_constant__ TmaMaps tm)
{
    constexpr int W = kWarpsPerBlock, R = 4;
    static_assert(!CFLF || !DIFF, "the CFL filter serves the kernels without diffusion");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double s_dt;
    __shared__ int s_flags;
    __shared__ CflFilter s_filter;
    __shared__ unsigned long long s_max[3];  // block maxima (x, y) and, CFLF, which of them the filter validated
    __shared__ int s_fix;
    const unsigned int nblocks = gridDim.x;
    const bool lead = blockIdx.x == 0;
    const bool xwarp = BANDS && lead && threadIdx.x < 32;
    unsigned int *const go = p.done_ctr + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long it0 = p.stream_map != 0 ? (long long)blockIdx.x : (long long)blockIdx.x * p.stream_unit;
    const long long lin1 = min((long long)(blockIdx.x + 1) * p.stream_unit, p.stream_total);
    const long long it_end = p.stream_map != 0 ? p.stream_total : lin1;
    unsigned int epoch = 0;  // arrival rounds of this launch so far (the same in every block)
    int tbase = 0, par = -1;

    auto arrive = [&]() {  // thread 0
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.done_ctr) : "memory");
    };
    auto wait_all = [&]() {  // thread 0: every block has arrived `epoch` times
        const unsigned int target = epoch * nblocks;
        const long long t0 = clock64();
        while ((int)(ld_acquire_gpu(p.done_ctr) - target) < 0) {
            if (clock64() - t0 > p.spin_limit) {  // never in a healthy run: turn a lost block into an error, not a hang
                atomicMin(p.err, -4);             // SWB_E_TIMEOUT
                atomicExch(p.done_ctr + 2, 1u);   // ... and into the end of the loop (step_prologue)
                s_flags = -1;
                break;
            }
        }
    };
    // thread 0: the block's maxima (CFLF: only what the filter let through, and its found-mask) -> the slot the next step reads
    auto flush = [&](Ctrl *slot, bool with_mask) {
        if (!CFLF || s_max[0] != 0ULL) atomicMax(&slot->ex_bits, s_max[0]);
        if (!CFLF || s_max[1] != 0ULL) atomicMax(&slot->ey_bits, s_max[1]);
        if (CFLF && with_mask && s_max[2] != 0ULL) atomicOr(&slot->reserved, (int)s_max[2]);
    };
    // CFLF fix-up round: the exact CFL keys of every cell this block updated, from the new state (set `cur`)
    auto fixup = [&](int cur) {
        if (threadIdx.x == 0) s_max[0] = s_max[1] = 0ULL;
        __syncthreads();
        const double *__restrict__ Hq = (const double *)p.buf[cur][0], *__restrict__ Uq = (const double *)p.buf[cur][1],
                     *__restrict__ Vq = (const double *)p.buf[cur][2];
        unsigned long long mx = 0ULL, my = 0ULL;
        long long it = it0;
        int col, r0, len;
        while (stream_next_piece(p, it, it_end, lin1, col, r0, len)) {
            const int strip = col * W + warp;
            if (strip >= p.nstrips) continue;
            const int c0 = kStripValid * strip, colA = c0 + 2 * lane;
            const bool vA = (lane >= 1) && (colA <= c0 + kStripValid) && (colA <= p.M + 1);
            const bool vB = (colA + 1 <= c0 + kStripValid) && (colA + 1 <= p.M + 1);
            if (!vA && !vB) continue;
            for (int jl = p.jl_first + r0; jl < p.jl_first + r0 + len; ++jl) {
                const size_t k = (size_t)jl * p.pitch + colA;
                const double2 hh = __ldcg(reinterpret_cast<const double2 *>(Hq + k)), uu = __ldcg(reinterpret_cast<const double2 *>(Uq + k)),
                              vv = __ldcg(reinterpret_cast<const double2 *>(Vq + k));
                unsigned long long kxA, kyA, kxB, kyB;
                cfl_keys_fast(hh.x, uu.x, vv.x, p.g, kxA, kyA);
                cfl_keys_fast(hh.y, uu.y, vv.y, p.g, kxB, kyB);
                if (vA) { mx = umax64(mx, kxA); my = umax64(my, kyA); }
                if (vB) { mx = umax64(mx, kxB); my = umax64(my, kyB); }
            }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            mx = umax64(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
            my = umax64(my, __shfl_xor_sync(0xffffffffu, my, sft));
        }
        if (lane == 0) {
            atomicMax(&s_max[0], mx);
            atomicMax(&s_max[1], my);
        }
        __syncthreads();
    };
    // BANDS, blocks other than block 0: wait for the token that admits them to step k of this launch (2k+2);
    // CFLF: a fix-up round (token 2k+1) for the state that step reads may come first
    auto wait_go = [&](int k, Ctrl *slot) {
        bool fixed = false;
        for (;;) {
            if (threadIdx.x == 0) {
                const unsigned int need = (CFLF && !fixed) ? (unsigned int)(2 * k + 1) : (unsigned int)(2 * k + 2);
                unsigned int v = 0;
                const long long t0 = clock64();
                while ((int)((v = ld_acquire_gpu(go)) - need) < 0) {
                    if (clock64() - t0 > 2 * p.spin_limit) {
                        atomicMin(p.err, -4);
                        atomicExch(p.done_ctr + 2, 1u);
                        v = 0x7fffffffu;
                        break;
                    }
                }
                s_fix = (CFLF && v == (unsigned int)(2 * k + 1)) ? 1 : 0;
            }
            __syncthreads();
            if (!CFLF || !s_fix) break;
            fixup(par);
            if (threadIdx.x == 0) {
                flush(slot, false);
                arrive();
            }
            ++epoch;
            fixed = true;
            __syncthreads();
        }
    };

    int s = 0;
    for (; s < p.nsteps; ++s) {
        Ctrl *const cur_slot = p.ctrl + ((p.launch + s) % 3);
        if (BANDS) {
            if (xwarp) {
                const bool ok = band_gather(p, p.launch + s, p.stamp + (unsigned long long)s, lane);
                if (lane == 0) {
                    __threadfence();
                    st_release_gpu(go, ok ? (unsigned int)(2 * s + 2) : 0x7fffffffu);  // (after a time-out: nobody waits any more)
                }
            }
            if (lead) __syncthreads();
            else wait_go(s, cur_slot);
        }
        // ---- the step: this block's tiles ----
        bool active = true;
        {
            int par_in = par;  // buffer set holding the current state (-1: not known yet), the same for every piece of the step
            bool head = true;
            long long it = it0;
            int col, r0, len;
            while (stream_next_piece(p, it, it_end, lin1, col, r0, len)) {
                const Tile tl{col * W, p.jl_first + r0, p.jl_first + r0 + len, lead, head};
                int pio = par_in;
                active = march_step<double, false, DIFF, 0, false, true, R, S, GUARD, W, CFLF, true, BANDS, true>(p, tm, p.launch + s, s == 0 && head, tbase, pio,
                                                                                                                    smem_raw, s_dt, s_flags, s_filter, s_max, tl);
                if (!active) break;
                par = pio;  // the set holding the NEW state: the next step's current one
                par_in = pio ^ 1;
                head = false;
            }
        }
        if (!active) break;
        // ---- grid barrier: all stores and atomics of step s happen before any load of step s+1.  (Every thread
        // first orders its generic-proxy stores before the async-proxy reads the next step's tensor loads are.)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncthreads();
        Ctrl *const new_slot = p.ctrl + ((p.launch + s + 1) % 3);
        const bool last = s + 1 == p.nsteps;
        if (!BANDS) {
            if (CFLF || !last) {
                if (threadIdx.x == 0) {
                    flush(new_slot, true);
                    arrive();
                }
                ++epoch;
                if (threadIdx.x == 0) {
                    wait_all();
                    if (CFLF) s_fix = ((ld1<LD_CG>(&new_slot->reserved) & 3) != 3) ? 1 : 0;
                }
                if (CFLF) {
                    __syncthreads();
                    if (s_fix) {  // uniform over the grid: every block reads the same word behind the barrier
                        fixup(par);
                        if (threadIdx.x == 0) {
                            flush(new_slot, false);
                            arrive();
                        }
                        ++epoch;
                        if (threadIdx.x == 0) wait_all();
                    }
                }
            } else if (threadIdx.x == 0) {
                flush(new_slot, true);
            }
            __syncthreads();
        } else {
            if (threadIdx.x == 0) {
                flush(new_slot, true);
                arrive();
            }
            ++epoch;
            if (lead) {
                if (threadIdx.x == 0) {
                    wait_all();  // every block of this band is done with step s
                    if (CFLF) {
                        s_fix = ((ld1<LD_CG>(&new_slot->reserved) & 3) != 3) ? 1 : 0;
                        if (s_fix) st_release_gpu(go, (unsigned int)(2 * (s + 1) + 1));  // the other blocks: fix-up round
                    }
                }
                if (CFLF) {
                    __syncthreads();
                    if (s_fix) {
                        fixup(par);
                        if (threadIdx.x == 0) {
                            flush(new_slot, false);
                            arrive();
                        }
                        ++epoch;
                        if (threadIdx.x == 0) wait_all();
                        __syncthreads();
                    }
                }
                if (xwarp) {  // tell the other bands
                    unsigned long long ex = 0ULL, ey = 0ULL;
                    if (lane == 0) {
                        ex = ld1<LD_CG>(&new_slot->ex_bits);
                        ey = ld1<LD_CG>(&new_slot->ey_bits);
                    }
                    ex = __shfl_sync(0xffffffffu, ex, 0);
                    ey = __shfl_sync(0xffffffffu, ey, 0);
                    band_publish(p, p.launch + s, p.stamp + (unsigned long long)s, ex, ey, lane);
                    if (CFLF && last && lane == 0) st_release_gpu(go, (unsigned int)(2 * (s + 1) + 2));  // the launch's final token
                }
                __syncthreads();
            } else if (CFLF && last) {
                wait_go(s + 1, new_slot);  // the state of the last step may still need its fix-up round
            }
        }
        if (s_flags == -1) break;
    }
    // Bands: a launch that ended early (time >= final_time at step s) still owes the other bands the stamps of its
    // remaining steps -- they are no-ops everywhere, with the maxima carried forward
    if (xwarp && s < p.nsteps && s_flags != -1) {
        for (int k = s; k < p.nsteps; ++k) {
            const Ctrl *carry = p.ctrl + ((p.launch + k + 1) % 3);
            unsigned long long ex = 0ULL, ey = 0ULL;
            if (lane == 0) {
                ex = ld1<LD_CG>(&carry->ex_bits);
                ey = ld1<LD_CG>(&carry->ey_bits);
            }
            ex = __shfl_sync(0xffffffffu, ex, 0);
            ey = __shfl_sync(0xffffffffu, ey, 0);
            band_publish(p, p.launch + k, p.stamp + (unsigned long long)k, ex, ey, lane);
            if (k + 1 < p.nsteps && !band_gather(p, p.launch + k + 1, p.stamp + (unsigned long long)(k + 1), lane)) break;
        }
    }
}

// Multi-GPU step epilogue (one warp, after the step kernel in stream order): lane r deposits
// this band's CFL maxima of the new state in rank r's exchange block and then releases the
// stamp at system scope: the release is cumulative, it orders every store of the finished step
// kernel -- the halo rows it wrote into the neighbours' arrays included -- before the stamp.
// (A separate __threadfence_system() in front of it cost ~4 us per step and bought nothing.)
__global__ void publish_kernel(const __grid_constant__ Params p)
{
    const Ctrl *new_slot = p.ctrl + ((p.launch + 1) % 3);
    pdl_enter();
    if (threadIdx.x == 0) SWB_TLOG(p, 5);
    band_publish(p, p.launch, p.stamp, new_slot->ex_bits, new_slot->ey_bits, (int)threadIdx.x);
#ifdef SWB_BAND_TIMING
    __syncwarp();
    if (threadIdx.x == 0) SWB_TLOG(p, 6);
#endif
}

// Multi-GPU step prologue (one warp, before the step kernel in stream order): lane r waits
// (on this GPU's own memory) for rank r's stamp of the previous step, then the warp reduces
// the bands' maxima and writes the GLOBAL maxima into the slot the step kernel reads --
// every band derives the identical dt.  The stamps' arrival is also the halo hand-shake
// and the write-after-read barrier of the ping-pong buffers.
__global__ void gather_kernel(const __grid_constant__ Params p)
{
    const int r = threadIdx.x;
    unsigned long long ex = 0ULL, ey = 0ULL;
    pdl_enter();
    if (r == 0) SWB_TLOG(p, 0);
    if (r < p.world) {
        const XSlot *xs = &p.xlocal->slot[p.launch % 3][r];
        const long long t0 = clock64();
        while (ld_acquire_sys(&xs->stamp) != p.stamp) {
            if (clock64() - t0 > p.spin_limit) {
                atomicMin(p.err, -4);            // SWB_E_TIMEOUT
                atomicExch(p.done_ctr + 2, 1u);  // ... and the loop ends here (step_prologue)
                break;
            }
            __nanosleep(100);
        }
        SWB_TLOG(p, 8 + r);
        ex = *reinterpret_cast<const volatile unsigned long long *>(&xs->ex_bits);
        ey = *reinterpret_cast<const volatile unsigned long long *>(&xs->ey_bits);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        ex = umax64(ex, __shfl_xor_sync(0xffffffffu, ex, s));
        ey = umax64(ey, __shfl_xor_sync(0xffffffffu, ey, s));
    }
    if (r == 0) {
        Ctrl *cur_slot = p.ctrl + (p.launch % 3);
        cur_slot->ex_bits = ex;
        cur_slot->ey_bits = ey;
        SWB_TLOG(p, 1);
    }
}

// CFL filter fix-up (after a step kernel with CFLF, in stream order): nothing to do when both
// maxima were found valid (the normal case: every block reads the mask and returns);
// otherwise the exact keys of every updated cell of the band, with the arithmetic of the
// step kernel.  Ghost columns and pole rows are copies of updated cells.
__global__ void __launch_bounds__(256) cfl_fixup_kernel(const __grid_constant__ Params p)
{
    Ctrl *new_slot = p.ctrl + ((p.launch + 1) % 3);
    pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) SWB_TLOG(p, 4);
    if ((new_slot->reserved & 3) == 3) return;
    const int cur = new_slot->current;
    const double *__restrict__ H = (const double *)p.buf[cur][0];
    const double *__restrict__ U = (const double *)p.buf[cur][1];
    const double *__restrict__ V = (const double *)p.buf[cur][2];
    unsigned long long mx = 0ULL, my = 0ULL;
    const int ncols = p.M + 1, nrows = p.jl_last - p.jl_first + 1;
    const long long total = (long long)nrows * ncols;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int jl = p.jl_first + (int)(k / ncols), i = 1 + (int)(k % ncols);
        const size_t a = (size_t)jl * p.pitch + i;
        unsigned long long kx, ky;
        cfl_keys_fast(H[a], U[a], V[a], p.g, kx, ky);
        mx = umax64(mx, kx);
        my = umax64(my, ky);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        mx = umax64(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        my = umax64(my, __shfl_xor_sync(0xffffffffu, my, s));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&new_slot->ex_bits, mx);
        atomicMax(&new_slot->ey_bits, my);
    }
}

// Stage 1 of the row records for every local latitude, once per context (swb_set_tables).
// `prec` (FAST): the same records once more as ONE table of (nrows + 2 kRecPad) rows -- row record, then
// diffusion record -- for the streaming kernel's bulk copies; the padding rows repeat the end rows.
template <bool STRICT>
__global__ void __launch_bounds__(128) prerec_kernel(const __grid_constant__ Params p, double *__restrict__ pre, double *__restrict__ pred,
                                                     double *__restrict__ prec)
{
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (prec != nullptr && jl < p.nrows + 2 * kRecPad) {  // padded row jl holds local latitude jl - kRecPad (clamped by the builders)
        const int stride = rec_stride(pred != nullptr);
        double rec[kRowRec];
        build_row_pre<STRICT>(p, jl - kRecPad, rec);
#pragma unroll
        for (int k = 0; k < kRowRec; ++k) prec[(size_t)jl * stride + k] = rec[k];
        if (pred != nullptr) {
            double rd[kRowDiff];
            build_diff_pre(p, jl - kRecPad, rd);
#pragma unroll
            for (int k = 0; k < kRowDiff; ++k) prec[(size_t)jl * stride + kRowRec + k] = rd[k];
        }
    }
    if (jl >= p.nrows) return;
    double rec[kRowRec];
    build_row_pre<STRICT>(p, jl, rec);
#pragma unroll
    for (int k = 0; k < kRowRec; ++k) pre[(size_t)jl * kRowRec + k] = rec[k];
    if (pred != nullptr) {
        double rd[kRowDiff];
        build_diff_pre(p, jl, rd);
#pragma unroll
        for (int k = 0; k < kRowDiff; ++k) pred[(size_t)jl * kRowDiff + k] = rd[k];
    }
}

// ------------------------------------------------------------------------------------
// K0: CFL maxima over the FULL ghosted initial state (numpy.py:503-521), into slot 0.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cfl_init_kernel(const T *__restrict__ H, const T *__restrict__ U,
                                                       const T *__restrict__ V, int ncols, int jl0, int jl1,
                                                       int pitch, double g, Ctrl *slot)
{
    unsigned long long mx = 0ULL, my = 0ULL;
    const long long total = (long long)(jl1 - jl0) * ncols;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int jl = jl0 + (int)(k / ncols), i = (int)(k % ncols);
        const size_t a = (size_t)jl * pitch + i;
        const double cs = __dsqrt_rn(__dmul_rn(g, fabs((double)H[a])));
        // the reference's three-way maximum, evaluated literally for step 1
        const double u = U[a], v = V[a];
        const unsigned long long bx0 = nonneg_bits(__dsub_rn(u, cs)), bx1 = nonneg_bits(u), bx2 = nonneg_bits(__dadd_rn(u, cs));
        const unsigned long long by0 = nonneg_bits(__dsub_rn(v, cs)), by1 = nonneg_bits(v), by2 = nonneg_bits(__dadd_rn(v, cs));
        mx = umax64(mx, umax64(bx0, umax64(bx1, bx2)));
        my = umax64(my, umax64(by0, umax64(by1, by2)));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        mx = umax64(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        my = umax64(my, __shfl_xor_sync(0xffffffffu, my, s));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&slot->ex_bits, mx);
        atomicMax(&slot->ey_bits, my);
    }
}

// K5: max sqrt(u*u + v*v) over the full state (numpy.py:553-554).
template <typename T>
__global__ void __launch_bounds__(256) max_speed_kernel(const T *__restrict__ U, const T *__restrict__ V,
                                                        int ncols, int jl0, int jl1, int pitch, unsigned long long *out)
{
    unsigned long long m = 0ULL;
    const long long total = (long long)(jl1 - jl0) * ncols;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int jl = jl0 + (int)(k / ncols), i = (int)(k % ncols);
        const size_t a = (size_t)jl * pitch + i;
        const double uu = U[a], vv = V[a];
        const double s = __dsqrt_rn(__dadd_rn(__dmul_rn(uu, uu), __dmul_rn(vv, vv)));
        m = umax64(m, nonneg_bits(s));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = umax64(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// Operator-level Laplacian (numpy.py:359-382 `compute_diffusion`), strict arithmetic.  Rows are
// folded over blockIdx.y (any number of latitudes).
__global__ void __launch_bounds__(128) laplacian_kernel(const double *__restrict__ q, double *__restrict__ out, Params p)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > p.M + 1) return;
    using O = Op<true>;
    for (int jl = p.jl_first + blockIdx.y; jl <= p.jl_last; jl += gridDim.y) {
        const Star<double> s = load_star(q, p.pitch, p.nrows, p.M, jl, i);
        const W3 wy = y_weights(p.tab, p.nrows, jl);
        const W3 wx = x_weights_strict(p.phi, p.tab[(size_t)T_AC * p.nrows + jl], i, p.M);
        out[(size_t)jl * p.pitch + i] = O::add(second_diff_strict(wx, s.xmm, s.xm, s.q0, s.xp, s.xpp),
                                               second_diff_strict(wy, s.ymm, s.ym, s.q0, s.yp, s.ypp));
    }
}

// The reference's `compute_laplacian` itself (numpy.py:66-133): the same weights applied to a field the
// CALLER has extended -- (N+2, pitch_ext) latitude-major, element (j+1, i+1) = q[i, j] -- with no wrap
// or clamp logic: q[2:-2, 2:-2] is the centre, q[1:-3], q[3:-1], q[:-4], q[4:] its neighbours.
__global__ void __launch_bounds__(128) laplacian_ext_kernel(const double *__restrict__ qe, int pe, double *__restrict__ out, Params p)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > p.M + 1) return;
    using O = Op<true>;
    for (int jl = p.jl_first + blockIdx.y; jl <= p.jl_last; jl += gridDim.y) {
        const double *c = qe + (size_t)(jl + 1) * pe + (i + 1);
        Star<double> s;
        s.q0 = c[0];
        s.xm = c[-1]; s.xp = c[1]; s.xmm = c[-2]; s.xpp = c[2];
        s.ym = c[-(ptrdiff_t)pe]; s.yp = c[pe]; s.ymm = c[-2 * (ptrdiff_t)pe]; s.ypp = c[2 * (ptrdiff_t)pe];
        const W3 wy = y_weights(p.tab, p.nrows, jl);
        const W3 wx = x_weights_strict(p.phi, p.tab[(size_t)T_AC * p.nrows + jl], i, p.M);
        out[(size_t)jl * p.pitch + i] = O::add(second_diff_strict(wx, s.xmm, s.xm, s.q0, s.xp, s.xpp),
                                               second_diff_strict(wy, s.ymm, s.ym, s.q0, s.yp, s.ypp));
    }
}

// K4: layout conversion between the reference's (M+3, n) array (latitude contiguous) and
// the device layout (n, pitch) (longitude contiguous): a tiled transpose through shared
// memory, both sides coalesced.  dst[r][c] = src[c][r] for r < rows, c < cols; converts between
// the reference's float64 and the state type on the way.
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) transpose_kernel(const TS *__restrict__ src, int src_pitch,
                                                        TD *__restrict__ dst, int dst_pitch, int rows, int cols)
{
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int c = c0 + ty + k, r = r0 + tx;  // src row = c, src col = r
        if (c < cols && r < rows) tile[ty + k][tx] = src[(size_t)c * src_pitch + r];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int r = r0 + ty + k, c = c0 + tx;
        if (r < rows && c < cols) dst[(size_t)r * dst_pitch + c] = (TD)tile[tx][ty + k];
    }
}

}  // namespace swb
