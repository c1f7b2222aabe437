/*
 * swb200.h -- C ABI of libswb200.so: the B200 (sm_100a) replacement for the body of the
 * time-stepping loop of ai2cm/sw_solver's numpy solver.
 *
 * The reference has no FFI; its operator API for this path is a set of Python functions
 * (src/sw_solver/numpy.py).  Each entry point below names the reference interface it
 * stands in for.  The Python host layer (sw_solver_b200/solver.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, a positive cudaError_t or a negative SWB_E_*
 *    code on failure; swb_last_error() returns a message for the calling thread;
 *  - no C++ types or exceptions cross this boundary; all pointers are plain;
 *  - state arrays are DEVICE pointers in the library's LATITUDE-MAJOR layout: C-contiguous
 *    [jl = local latitude 0..n_local-1][i = longitude 0..M+2], longitude contiguous,
 *    `pitch` elements between consecutive latitudes (the transpose of the reference's
 *    (M+3, N) array; swb_import_state / swb_export_state convert on the device).  pitch
 *    must be at least swb_required_pitch(M, dtype) -- the kernels read whole 64-column
 *    strips -- bases 16-byte aligned; the padding columns are never written.  State
 *    arrays hold doubles (SWB_F64) or floats (SWB_F32: FAST arithmetic, latitude-only
 *    Coriolis, flat topography; every metric / dt factor is still formed in fp64);
 *  - table pointers passed to swb_set_tables are HOST pointers (copied);
 *  - `stream` is a cudaStream_t passed as void*; calls only enqueue work unless stated;
 *  - the caller (PyTorch on the Python side) owns every state buffer and keeps it alive
 *    while it is bound; the library owns only its small scratch;
 *  - there is no CPU fallback: without a CUDA device swb_create fails.
 */
#ifndef SWB200_H
#define SWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB_ABI_VERSION 3

enum {
    SWB_E_INVALID = -1,  /* bad argument / configuration                    */
    SWB_E_STATE = -2,    /* call sequence error (e.g. state not bound)      */
    SWB_E_NODEVICE = -3, /* no CUDA device / wrong architecture             */
    SWB_E_TIMEOUT = -4,  /* a peer never signalled (multi-GPU)              */
    SWB_E_NONFINITE = -5,/* reserved                                        */
    SWB_E_DEPTH = -6     /* SWB_FAST met a non-positive half-step layer depth: its
                            results are not the reference's there (which switches
                            formulas, numpy.py:287-292, 321-322); rerun with
                            SWB_FAST_GUARDED or SWB_STRICT                  */
};

enum { SWB_F64 = 0, SWB_F32 = 1 };

/* Arithmetic mode.  STRICT evaluates every expression in the reference's operation order
 * with IEEE add/mul/div/sqrt and no FMA contraction: results are bit-identical to numpy.
 * FAST computes each face once with FMA contraction and shared reciprocals: results
 * agree with numpy to <= 1e-12 (normwise) over benchmark-length runs.  FAST assumes the
 * half-step layer depths stay positive (every physical run) and raises SWB_E_DEPTH in
 * swb_status.error otherwise; FAST_GUARDED evaluates the reference's `hMid > 0` selects. */
enum { SWB_STRICT = 0, SWB_FAST = 1, SWB_FAST_GUARDED = 2 };

typedef struct swb_ctx swb_ctx;

typedef struct swb_config {
    int32_t struct_bytes; /* sizeof(swb_config), for ABI checking                     */
    int32_t M;            /* num_lon_pts (state has M+3 longitude columns)            */
    int32_t N;            /* global num_lat_pts, already even (grid.py:83)            */
    int32_t j_begin;      /* global latitude index of local row 0 (0 on one GPU)     */
    int32_t n_local;      /* latitude rows held locally, halo/ghost rows included     */
    int32_t j_first;      /* first global latitude row this context UPDATES (>= 1)    */
    int32_t j_last;       /* last global latitude row this context updates (<= N-2)   */
    int32_t pitch;        /* elements between consecutive latitudes in state arrays   */
    int32_t dtype;        /* SWB_F64 | SWB_F32                                        */
    int32_t mode;         /* SWB_STRICT | SWB_FAST                                    */
    int32_t use_diffusion;/* numpy.py:541-544                                         */
    int32_t f_is_2d;      /* 0: Coriolis depends on latitude only (table); 1: array   */
    int32_t has_hs;       /* 0: flat topography (term skipped); 1: hs array bound     */
    int32_t device;       /* CUDA device ordinal                                      */
    int32_t log_capacity; /* steps of (dt, time) history kept on the device (0 = none)*/
    int32_t rank;         /* latitude-band index (0 on one GPU)                       */
    int32_t world;        /* number of latitude bands / GPUs (1 on one GPU)           */
    int32_t halo;         /* rows exchanged with each neighbour band (1; 2 with
                             diffusion); ignored when world == 1                      */
    int32_t f_sep;        /* 1: the Coriolis parameter is SEPARABLE,
                             f[i,j] = fsep_scale * ((fsep_lon[i] * fsep_lat_b[j]) * fsep_sa + fsep_lat_d[j])
                             evaluated in exactly this order -- Williamson case 2 of the reference
                             (ic.py:126-133) from 1-D tables: no (M+3, N) array, and the TMA / resident
                             kernels apply.  fp64, flat topography, f_is_2d == 0.                  */
    int32_t reserved0;
    double courant;       /* CFL number (numpy.py:413)                                */
    double final_time;    /* seconds (numpy.py:470-472)                               */
    double dxmin, dymin;  /* CartesianGrid.dxmin / dymin (grid.py:46-47)              */
    double g, a, nu;      /* EARTH_CONSTANTS (utils.py:28-33)                         */
    double dphi;          /* 2 pi / M (grid.py:86)                                    */
    double fsep_scale;    /* f_sep: the outer factor (2 Omega)                        */
    double fsep_sa;       /* f_sep: the factor of the product term (sin(alpha))       */
} swb_config;

/* Latitude tables of the local band (host pointers, each n_local long, indexed by local
 * row; entries that do not exist for a row -- e.g. face values of the last row -- are
 * ignored).  Replaces the 2-D metric arrays of LatLonGrid/CartesianGrid (grid.py:35-53,
 * 99-108) and DiffusionCoefficients' y-set (numpy.py:50-63). */
typedef struct swb_tables {
    const double *phi;    /* M+3 longitudes (grid.py:86-87)                           */
    const double *c;      /* cos(theta_j)                          LatLonGrid.c       */
    const double *tg;     /* tan(theta_j)                          LatLonGrid.tg      */
    const double *ac;     /* a*cos(theta_j): x[i,j] = ac[j]*phi[i] CartesianGrid.x    */
    const double *f_lat;  /* Coriolis f_j when f_is_2d == 0 (ic.py:46)                */
    const double *tgMidy; /* tan at face (j | j+1)                 LatLonGrid.tgMidy  */
    const double *cMidy;  /* cos at face (j | j+1)                 LatLonGrid.cMidy   */
    const double *dy;     /* y[j+1]-y[j]                           CartesianGrid.dy   */
    const double *dy1;    /* y1[j+1]-y1[j]                         CartesianGrid.dy1  */
    const double *dyc;    /* 0.5*(dy[j-1]+dy[j])                   CartesianGrid.dyc  */
    const double *dy1c;   /* 0.5*(dy1[j-1]+dy1[j])                 CartesianGrid.dy1c */
    const double *Ay, *By, *Cy; /* DiffusionCoefficients.Ay/By/Cy of row j (may be NULL
                                   when use_diffusion == 0)                           */
    const double *fsep_lon;   /* f_sep: M+3 longitude factors   (-cos(phi_i), ic.py:128)          */
    const double *fsep_lat_b; /* f_sep: n_local latitude factors (cos(theta_j))                   */
    const double *fsep_lat_d; /* f_sep: n_local latitude addends (sin(theta_j) * cos(alpha))      */
} swb_tables;

/* Small status record mirrored from the device (see swb_get_status). */
typedef struct swb_status {
    double time;       /* simulated seconds reached                                   */
    double eigenx;     /* CFL maxima of the CURRENT state (numpy.py:503-521); with
                          several bands: of this band's rows                          */
    double eigeny;
    double last_dt;    /* dt of the most recent step                                  */
    int64_t steps;     /* completed time steps (numpy.py:500 num_steps)               */
    int32_t done;      /* 1 once !(time < final_time) (numpy.py:496)                  */
    int32_t current;   /* which bound buffer set (0/1) holds the current state        */
    int32_t error;     /* 0, or a negative SWB_E_* raised on the device               */
    int32_t reserved;
} swb_status;

int swb_abi_version(void);
const char *swb_last_error(void);

/* Lifetime.  Replaces nothing in the reference (it keeps its state in Python locals,
 * numpy.py:464-494); holds what `solve` sets up before its loop. */
int swb_create(swb_ctx **out, const swb_config *cfg);
int swb_destroy(swb_ctx *ctx);

int swb_set_tables(swb_ctx *ctx, const swb_tables *tabs);

/* Smallest legal `pitch` for a grid with M longitudes and state type `dtype` (a multiple of
 * 16 elements). */
int swb_required_pitch(int M, int dtype);

/* Bind the two buffer sets the path ping-pongs between (set 0 holds the initial state),
 * plus the optional 2-D Coriolis array and topography (device pointers or NULL), all
 * (n_local, pitch) arrays in the latitude-major device layout.  The TMA-fed kernels read
 * h, u, v of a buffer set through one 3-D tensor map: they are used when u - h == v - u (a
 * positive multiple of 16 bytes) in both sets -- e.g. one (3, n_local, pitch) allocation per
 * set; any other placement runs the direct-load kernels. */
int swb_bind_state(swb_ctx *ctx, void *h0, void *u0, void *v0, void *h1, void *u1, void *v1,
                   const void *f2d, const void *hs);

/* Layout conversion on the device (numpy.py:488-492, 559-563 snapshots; the IC upload):
 * `ref` is a window of an array in the reference's layout -- element (i, jl) at
 * ref[i*ref_pitch + jl], i = 0..M+2, jl = 0..n_local-1 -- and `dev` an (n_local, pitch)
 * array in the device layout.  Both are device pointers. */
int swb_to_device_layout(swb_ctx *ctx, const void *ref, int ref_pitch, void *dev, void *stream);
int swb_from_device_layout(swb_ctx *ctx, const void *dev, void *ref, int ref_pitch, void *stream);

/* numpy.py:503-521 on the FULL ghosted initial state (ghost cells hold genuine initial
 * values at step 1); resets time and step count to zero. */
int swb_cfl_init(swb_ctx *ctx, void *stream);

/* Enqueue n iterations of the loop body numpy.py:496-549 (dt from the device-resident
 * CFL maxima, final-time clamp, Lax-Wendroff update, optional diffusion, boundary
 * conditions, CFL maxima of the new state).  Never blocks; iterations enqueued after the
 * final time has been reached do nothing. */
int swb_enqueue_steps(swb_ctx *ctx, int n, void *stream);

/* One Lax-Wendroff update (+ diffusion increment when configured) with a caller-given
 * dt, from buffer set 0 into buffer set 1: numpy.py:136-356 `lax_wendroff_update`,
 * 359-382 `compute_diffusion`, 385-406 `_apply_bcs`.  Does not touch time bookkeeping. */
int swb_step_fixed_dt(swb_ctx *ctx, double dt, void *stream);

/* Laplacian of one field (numpy.py:359-382 `compute_diffusion`): q is an (n_local, pitch)
 * device array, out receives the same shape with the interior filled. */
int swb_laplacian(swb_ctx *ctx, const void *q, void *out, void *stream);
/* Laplacian of an ALREADY EXTENDED field (numpy.py:66-133 `compute_laplacian`): the reference's
 * (M+5, N+2) argument, here latitude-major -- `qext` is an (N+2, pitch_ext) device array of
 * doubles, pitch_ext >= M+5, element (j+1, i+1) holding q[i, j] of the unextended field.  No
 * periodic wrap or row duplication is applied: whatever extension the caller built is used as is.
 * `out` as for swb_laplacian.  One band (world == 1) only. */
int swb_laplacian_ext(swb_ctx *ctx, const void *qext, int pitch_ext, void *out, void *stream);

/* max sqrt(u^2+v^2) over the full current state (numpy.py:553-554); result (double)
 * is written to *out_host after an internal stream synchronisation. */
int swb_max_speed(swb_ctx *ctx, double *out_host, void *stream);
/* The same reduction over buffer set `current` (0 or 1), enqueued only: the result, a double,
 * arrives in `out` (8 bytes of pinned host or device memory) in stream order.  With
 * swb_status_async in front of it a print (numpy.py:552-557) costs ONE synchronisation. */
int swb_max_speed_async(swb_ctx *ctx, int current, void *out, void *stream);

/* Enqueue an asynchronous copy of the status record as of all steps enqueued so far into
 * `host_out` (ideally pinned memory; it must stay valid until the stream reaches the
 * copy), or into the context's own pinned mirror when host_out is NULL. */
int swb_status_async(swb_ctx *ctx, swb_status *host_out, void *stream);
/* Read the context's pinned mirror without blocking. */
int swb_status_peek(swb_ctx *ctx, swb_status *out);
/* swb_status_async(NULL) + stream synchronise + peek. */
int swb_get_status(swb_ctx *ctx, swb_status *out, void *stream);

/* Copy the first `count` (dt, time) pairs of the step history to host arrays. */
int swb_read_log(swb_ctx *ctx, double *dt_out, double *time_out, int count, void *stream);

/* How long (seconds, ~2 GHz SM clock) a kernel waits for a peer band's stamp or for the grid barrier
 * before it raises SWB_E_TIMEOUT and ends the time loop (later steps become no-ops).  Default 10 s;
 * the environment variable SWB_SPIN_SECONDS sets the default of new contexts. */
int swb_set_spin_limit(swb_ctx *ctx, double seconds);

/* Number of kernels this context has launched so far (for bench.py's gpu_launches). */
int64_t swb_launch_count(const swb_ctx *ctx);

/* Non-zero if swb_enqueue_steps runs each batch of steps as ONE cooperative launch of the resident
 * kernel (the loop of numpy.py:496-549 then never leaves the device, a grid barrier replaces the
 * launch per step), else 0: 1 = one tile per block (grids whose blocks all fit the GPU at once),
 * 2 = the streaming variant (larger grids: every block marches several tiles per step, its row
 * records travelling with the tiles).  Decided by swb_bind_state; the environment variable
 * SWB_RESIDENT=0 keeps the launch-per-step kernel, SWB_STREAM=0 / 1 overrides the variant. */
int swb_resident(const swb_ctx *ctx);

/* Latitude bands (world > 1): allow the resident kernel for this band too -- the exchange with
 * the other bands (CFL maxima, stamps) then runs inside the launch, between two steps.  Call
 * after swb_link_peers and ONLY when every band of the globe runs on a GPU of its own (one
 * process per GPU): resident kernels of several bands on one GPU would wait for each other
 * without all being able to start.  A band that stays on the launch-per-step kernel (too large)
 * and a resident neighbour interoperate: same slots, same stamps. */
int swb_enable_resident_bands(swb_ctx *ctx);

/* ---- multi-GPU (one process per GPU, latitude bands) ------------------------------ */

#define SWB_IPC_HANDLE_BYTES 64

/* Export / open a CUDA IPC handle for the allocation containing `dptr`; `offset` is the
 * byte offset of dptr inside that allocation. */
int swb_ipc_export(const void *dptr, unsigned char handle[SWB_IPC_HANDLE_BYTES], uint64_t *offset);
int swb_ipc_open(const unsigned char handle[SWB_IPC_HANDLE_BYTES], int device, void **base);
int swb_ipc_close(void *base);

/* Device pointer and size of this context's exchange block (CFL slots + arrival flags)
 * that peers write into. */
int swb_exchange_block(swb_ctx *ctx, void **dptr, uint64_t *bytes);

/* Give the context its peers: for every band r, the (peer-mapped) address of its exchange
 * block (this context's own included); for the south (rank-1) and north (rank+1)
 * neighbours the addresses of their two buffer sets (h0,u0,v0,h1,u1,v1), their j_begin
 * and pitch.  Absent neighbours: NULL.  From then on every step stores the rows a
 * neighbour reads as halo straight into the neighbour's arrays (peer stores over NVLink)
 * and deposits this band's CFL maxima in every exchange block; the next step's prologue
 * waits for all bands' maxima -- no host call, no NCCL on the step path. */
int swb_link_peers(swb_ctx *ctx, void *const *exchange_blocks,
                   void *const south_bufs[6], int south_j_begin, int south_pitch,
                   void *const north_bufs[6], int north_j_begin, int north_pitch);

#ifdef __cplusplus
}
#endif
#endif /* SWB200_H */
