// swb_types.cuh -- constants, table / record layouts and the parameter blocks shared by the
// host code (swb200.cu) and the device code of libswb200.so.  Overview: swb_kernels.cuh.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace swb {

constexpr int kWarpsPerBlock = 4;
constexpr int kStripCols = 64;   // longitudes held by a warp (2 per lane)
constexpr int kStripValid = 62;  // longitudes updated by a warp (columns 1..62 of the strip), fp64
constexpr int kStripValidF32 = 60;  // fp32: strips must start on 16-byte columns (TMA), i.e. multiples of 4
constexpr int kMaxChunk = 128;   // latitudes marched by one warp, at most
constexpr int kRowRec = 16;      // doubles per row record
constexpr int kRowDiff = 8;      // doubles per diffusion row record
constexpr int kMaxRanks = 8;
constexpr int kRecPad = 8;       // rows of padding either side of the streamed record table (Params::prec)

// row indices of the packed latitude table (each row `nrows` long, indexed by local latitude)
enum Tab { T_C = 0, T_TG, T_AC, T_F, T_TGMIDY, T_CMIDY, T_DY, T_DY1, T_DYC, T_DY1C, T_AY, T_BY, T_CY, T_FB, T_FD, T_COUNT };
// (T_FB, T_FD: latitude factor and addend of a separable Coriolis parameter, see Params::fa)

// slots of a row record (shared memory, one record per latitude of the block's chunk)
enum Rec {
    RC_C = 0,   // cos(theta_j)
    RC_AC,      // a cos(theta_j)
    RC_KX,      // x-face metric factor   STRICT: tan(theta_j)      FAST: hdt*0.5*tan/a
    RC_FKX,     // x-face Coriolis        STRICT: f_j               FAST: hdt*f_j
    RC_CDX,     // FAST: dt/dx_j
    RC_CXU,     // FAST: 0.5*dt/dx_j
    RC_KY,      // y-face (j|j+1) metric  STRICT: tgMidy_j          FAST: hdt*0.5*tgMidy/a
    RC_FKY,     // y-face Coriolis        STRICT: 0.5*(f_j+f_j+1)   FAST: hdt*0.5*(f_j+f_j+1)
    RC_CM,      // cMidy_j
    RC_CY1F,    // STRICT: hdt/dy1_j      FAST: dt/dy1_j
    RC_CYF,     // STRICT: hdt/dy_j       FAST: dt/dy_j
    RC_CY1U,    // STRICT: dt/dy1c_j      FAST: 0.5*dt/dy1c_j
    RC_CYU,     // STRICT: dt/dyc_j       FAST: 0.5*dt/dyc_j
    RC_KC,      // update metric factor   STRICT: tan(theta_j)      FAST: dt/64*tan/a
    RC_FD,      // update Coriolis        STRICT: f_j               FAST: dt/8*f_j
    RC_DY1S     // dy1_{j-1} + dy1_j (topography term)
};

// Device-resident loop state (numpy.py:494-496, 500, 524-532).  Three slots rotate with
// the launch index L: kernel L reads slot L%3, writes time/steps and accumulates the new
// CFL maxima in slot (L+1)%3 and clears the maxima of slot (L+2)%3.
struct Ctrl {  // same layout as swb_status (include/swb200.h): mirrored to the host by a plain copy
    double time;
    unsigned long long ex_bits;  // bit pattern of a non-negative double (or NaN): ordered as uint64
    unsigned long long ey_bits;
    double last_dt;
    long long nsteps;
    int done;
    int current;  // nsteps & 1: the buffer set holding the current state
    int error;
    int reserved;
};

// Multi-GPU exchange block (one per context, peers write into it): per launch slot and per
// rank the CFL maxima of that rank's band plus an arrival stamp.
struct XSlot {
    unsigned long long ex_bits, ey_bits;
    unsigned long long stamp;  // launch index + 1 of the step that produced the maxima
    unsigned long long pad;
};
struct XBlock {
    XSlot slot[3][kMaxRanks];
};

// What one call of `march_step` covers: warps strip0 .. strip0+W-1, latitudes [ja, jb).  `lead`: the
// block that advances the loop state (time, step count); `prologue`: this call starts a time step
// (the streaming kernel marches several pieces per step).
struct Tile {
    int strip0, ja, jb;
    bool lead, prologue;
};

struct Params {
    void *buf[2][3];    // two buffer sets {h,u,v} (fp64 or fp32); set (nsteps & 1) holds the current state
    const void *f2d;    // (nrows, pitch) or null
    const void *hs;     // (nrows, pitch) or null
    const double *phi;  // M+3 (+ padding)
    // separable Coriolis parameter (FM == 2), f[i, j] = fs_scale * ((fa[i] * tab[T_FB][j]) * fs_sa + tab[T_FD][j]) in exactly
    // this order: Williamson case 2 (ic.py:126-133 of the reference: 2 Omega (-cos(phi) cos(theta) sin(alpha) + sin(theta) cos(alpha)))
    // from two 1-D tables instead of an (M+3, N) array
    const double *fa;   // M+3 (+ padding): the longitude factor
    double fs_scale, fs_sa;
    const double *tab;  // T_COUNT x nrows
    const double *pre;  // nrows x kRowRec: dt-independent stage of the row records (prerec_kernel)
    const double *pred; // nrows x kRowDiff: the same for the FAST diffusion records, or null
    Ctrl *ctrl;         // 3 slots
    int *err;           // sticky device-side error word (0 or a negative SWB_E_*)
    unsigned int *done_ctr;  // grid-barrier arrival counter of resident_kernel (zeroed before each launch)
    double *log;        // 2 x log_cap: dt then time, or null
    int log_cap;
    int M, nrows, pitch;    // nrows = local latitudes (halo / pole rows included)
    int jl_first, jl_last;  // local latitude rows updated by this context (inclusive)
    int pole_s, pole_n;     // local row 0 / nrows-1 is the global copy row 0 / N-1
    int chunk;              // latitudes marched by one warp
    int n_big, chunk_small; // march_kernel: block rows 0 .. n_big-1 march `chunk` latitudes, the rest `chunk_small` (the short
                            // marches are dispatched last and shorten the tail of the launch); n_big == 0: every block row `chunk`
    int nstrips;            // warps needed along longitude
    int launch;             // launch index L (mod 3 cycle)
    int par_guess;          // march_kernel: the buffer set the host expects to hold the current state (launches since swb_cfl_init, mod 2)
    int nsteps;             // resident_kernel: time steps run by this one launch
    int pre_smem;           // resident_kernel: the dt-independent row records are kept in shared memory (short marches)
    long long *dbg;         // SWB_RES_TIMING builds: per block 8 clock64 stamps of one step (dev instrumentation)
    int fixed;              // 1: use fixed_dt, buffers 0 -> 1, no bookkeeping
    int force_rare;         // tuning / test knob (SWB_FORCE_RARE=1): every row takes the boundary-row body
    double fixed_dt;
    double courant, final_time, dxmin, dymin, g, a, nu, dphi;
    // multi-GPU (world > 1): peers' exchange blocks and neighbour halo rows
    int rank, world;
    unsigned long long stamp;       // launch sequence number (monotone, never wraps in practice)
    XBlock *xlocal;                 // this context's exchange block
    XBlock *xpeer[kMaxRanks];       // every rank's exchange block as mapped here (self included)
    void *south[2][3];              // south neighbour's buffer sets (peer mapped) or null
    void *north[2][3];
    int south_row, north_row;       // row index, in the neighbour's arrays, of the halo row we fill
    int south_pitch, north_pitch;
    int halo;                       // rows exchanged per side (1, or 2 with diffusion)
    long long spin_limit;           // clock64 ticks to wait for peers before flagging SWB_E_TIMEOUT
    // streaming resident kernel (resident_kernel<..., STREAM>): every block marches a contiguous range of
    // `stream_unit` rows of the (block column, latitude) space, its row records arriving with the tiles
    const double *prec;             // (nrows + 2 kRecPad) x (kRowRec [+ kRowDiff]) dt-independent records; points at local row 0
    long long stream_total;         // map 0: block columns x updated rows; map 1: number of tiles
    int stream_unit;                // rows per block range (map 0) / per tile (map 1), a multiple of 4
    int stream_rows;                // updated rows of this context (jl_last - jl_first + 1)
    int stream_gx;                  // block columns
    int stream_map;                 // 0: one contiguous range of the column-major (column, row) space per block
                                    // 1: tiles of stream_unit rows dealt round-robin, row-major over the columns (the
                                    //    blocks sweep the grid together: DRAM-page / TLB locality of a launch per step)
    unsigned long long *tlog;       // SWB_BAND_TIMING builds: kTlogCap x kTlogRow globaltimer stamps, indexed by stamp % kTlogCap
};

// SWB_BAND_TIMING (dev instrumentation of the launch-per-step band path): per step one row of
// nanosecond stamps -- [0] gather enter, [1] gather exit, [2] first block of the step kernel starts,
// [3] its last block ends, [4] fix-up kernel starts, [5] publish enters, [6] publish done (stamp
// released), [8 + r] the stamp of rank r was seen by the gather.  scripts/band_timing.py reads the dump.
constexpr int kTlogCap = 4096;
constexpr int kTlogRow = 16;

}  // namespace swb
