// swb_kernels.cuh -- device code of libswb200.so (sm_100a).
//
// The hot path is one device function, `march_step`, that replaces the body of the loop at
// src/sw_solver/numpy.py:496-549 of the reference.  Two kernels call it: `march_kernel` (one
// launch per time step: any grid, any band decomposition) and `resident_kernel` (grids whose
// blocks all fit the GPU at once: the whole batch of steps in one cooperative launch, a grid
// barrier per step).  Per step:
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

// One launch per time step (any grid size, any band decomposition).
template <typename T, bool STRICT, bool DIFF, bool F2D, bool HS, bool TMA, int R, int S, bool GUARD, int MINB, int W = kWarpsPerBlock,
          bool CFLF = false>
__global__ void __launch_bounds__(W * 32, MINB)
march_kernel(const __grid_constant__ Params p, const __grid_constant__ TmaMaps tm)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double s_dt;
    __shared__ int s_flags;
    __shared__ CflFilter s_filter;
    int tbase = 0, par = -1;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) SWB_TLOG(p, 2);
    const int ja = p.jl_first + blockIdx.y * p.chunk;
    const Tile tl{(int)blockIdx.x * W, ja, min(ja + p.chunk, p.jl_last + 1), blockIdx.x == 0 && blockIdx.y == 0, true};
    march_step<T, STRICT, DIFF, F2D, HS, TMA, R, S, GUARD, W, CFLF, false>(p, tm, p.launch, true, tbase, par, smem_raw, s_dt, s_flags, s_filter,
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
                atomicMin(p.err, -4);  // SWB_E_TIMEOUT
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
//
// STREAM: grids beyond one wave of blocks (7200 x 3600, 16384 x 8192, latitude bands of any size).  The
// (block column, latitude) space is cut into one contiguous range of `stream_unit` rows per block --
// equal work for every block, whatever the grid -- which the block's warps march in one go (two, when
// the range crosses into the next block column), each warp on its own: the row records arrive with the
// tiles (see march_step).  No launch per step, no wave tail, no march start-up per 64 rows.
template <typename T, bool STRICT, bool DIFF, bool TMA, int R, int S, bool GUARD, int MINB, bool BANDS = false, int W = kWarpsPerBlock,
          bool STREAM = false>
__global__ void __launch_bounds__(W * 32, MINB)
resident_kernel(const __grid_constant__ Params p, const __grid_constant__ TmaMaps tm)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double s_dt;
    __shared__ int s_flags;
    __shared__ CflFilter s_filter;
    __shared__ unsigned long long s_max[2];
    const unsigned int nblocks = gridDim.x * gridDim.y;
    // STREAM: this block's range of the linearised (block column, row) space
    const long long lin0 = STREAM ? (long long)blockIdx.x * p.stream_unit : 0;
    const long long lin1 = STREAM ? min(lin0 + p.stream_unit, p.stream_total) : 0;
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
                        break;
                    }
                }
            }
            __syncthreads();
        }
        bool active = true;
        if constexpr (STREAM) {
            int par_in = par;  // buffer set holding the current state (-1: not known yet), the same for every piece of the step
            bool head = true;
            for (long long lin = lin0; lin < lin1;) {
                const int col = (int)(lin / p.stream_rows), r0 = (int)(lin % p.stream_rows);
                const int len = (int)min((long long)(p.stream_rows - r0), lin1 - lin);
                const Tile tl{col * W, p.jl_first + r0, p.jl_first + r0 + len, blockIdx.x == 0, head};
                int pio = par_in;
                active = march_step<T, STRICT, DIFF, false, false, TMA, R, S, GUARD, W, false, true, BANDS, true>(p, tm, p.launch + s, s == 0 && head, tbase, pio,
                                                                                                                 smem_raw, s_dt, s_flags, s_filter, s_max, tl);
                if (!active) break;
                par = pio;          // the set holding the NEW state: the next step's current one
                par_in = pio ^ 1;
                head = false;
                lin += len;
            }
        } else {
            const int ja = p.jl_first + blockIdx.y * p.chunk;
            const Tile tl{(int)blockIdx.x * W, ja, min(ja + p.chunk, p.jl_last + 1), blockIdx.x == 0 && blockIdx.y == 0, true};
            active = march_step<T, STRICT, DIFF, false, false, TMA, R, S, GUARD, W, false, true, BANDS>(p, tm, p.launch + s, s == 0, tbase, par, smem_raw,
                                                                                                      s_dt, s_flags, s_filter, s_max, tl);
        }
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

// Multi-GPU step epilogue (one warp, after the step kernel in stream order): lane r deposits
// this band's CFL maxima of the new state in rank r's exchange block and then releases the
// stamp at system scope: the release is cumulative, it orders every store of the finished step
// kernel -- the halo rows it wrote into the neighbours' arrays included -- before the stamp.
// (A separate __threadfence_system() in front of it cost ~4 us per step and bought nothing.)
__global__ void publish_kernel(const __grid_constant__ Params p)
{
    const Ctrl *new_slot = p.ctrl + ((p.launch + 1) % 3);
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
    if (r == 0) SWB_TLOG(p, 0);
    if (r < p.world) {
        const XSlot *xs = &p.xlocal->slot[p.launch % 3][r];
        const long long t0 = clock64();
        while (ld_acquire_sys(&xs->stamp) != p.stamp) {
            if (clock64() - t0 > p.spin_limit) {
                atomicMin(p.err, -4);  // SWB_E_TIMEOUT
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

// Operator-level Laplacian (numpy.py:359-382 `compute_diffusion`), strict arithmetic.
__global__ void __launch_bounds__(128) laplacian_kernel(const double *__restrict__ q, double *__restrict__ out, Params p)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int jl = p.jl_first + blockIdx.y;
    if (jl > p.jl_last || i > p.M + 1) return;
    using O = Op<true>;
    const Star<double> s = load_star(q, p.pitch, p.nrows, p.M, jl, i);
    const W3 wy = y_weights(p.tab, p.nrows, jl);
    const W3 wx = x_weights_strict(p.phi, p.tab[(size_t)T_AC * p.nrows + jl], i, p.M);
    out[(size_t)jl * p.pitch + i] = O::add(second_diff_strict(wx, s.xmm, s.xm, s.q0, s.xp, s.xpp),
                                           second_diff_strict(wy, s.ymm, s.ym, s.q0, s.yp, s.ypp));
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
