// swb_kernels.cuh -- device code of libswb200.so (sm_100a).
//
// The hot path is ONE kernel per time step (`step_kernel`) that replaces the body of the
// loop at src/sw_solver/numpy.py:496-549 of the reference:
//   * prologue: dt from the device-resident CFL maxima + final-time clamp (numpy.py:524-532)
//   * Lax-Wendroff half-step face values and full-step update      (numpy.py:185-354)
//   * optional Laplacian diffusion increment on the OLD fields       (numpy.py:541-544, 66-133)
//   * boundary conditions written in place of the ghost cells        (numpy.py:402-404)
//   * epilogue: CFL maxima of the NEW state for the next step        (numpy.py:503-521)
//
// Data layout (HBM): the reference's own -- state[i][j], i = longitude 0..M+2,
// j = latitude (contiguous).  A warp covers 32 consecutive latitudes and MARCHES along
// longitude: the x-face (i | i+1) computed at one iteration is the left face of the next
// one and stays in registers; y-face inputs/results travel between neighbouring lanes by
// warp shuffles.  Lanes 1..30 update cells (warps overlap by two latitudes), so every
// face is computed exactly once per warp.  All metric terms are per-lane registers
// derived from 1-D latitude tables; nothing 2-D but h, u, v is read from HBM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace swb {

constexpr int kWarpsPerBlock = 4;
constexpr int kCellsPerWarp = 30;  // lanes 1..30 update cells
constexpr int kMaxRanks = 8;

// row indices of the packed latitude table (each row n_local long)
enum Tab { T_C = 0, T_TG, T_AC, T_F, T_TGMIDY, T_CMIDY, T_DY, T_DY1, T_DYC, T_DY1C, T_AY, T_BY, T_CY, T_COUNT };

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

struct Params {
    double *buf[2][3];  // two buffer sets {h,u,v}; set (nsteps & 1) holds the current state
    const double *f2d;  // (M+3, pitch) or null
    const double *hs;   // (M+3, pitch) or null
    const double *phi;  // M+3
    const double *tab;  // T_COUNT x ncols
    Ctrl *ctrl;         // 3 slots
    double *log;        // 2 x log_cap: dt then time, or null
    int log_cap;
    int M, ncols, pitch;    // ncols = lead_pad + n_local: local columns [jl_lo, jl_hi] are real
    int jl_lo, jl_hi;
    int jl_first, jl_last;  // local latitude rows updated by this context (inclusive)
    int pole_s, pole_n;     // local row jl_lo / jl_hi is the global copy row 0 / N-1
    int chunk;              // longitudes marched by one warp
    int nstrips;            // warps needed along latitude
    int launch;             // launch index L
    int fixed;              // 1: use fixed_dt, buffers 0 -> 1, no bookkeeping
    double fixed_dt;
    double courant, final_time, dxmin, dymin, g, a, nu, dphi;
};

// ------------------------------------------------------------------------------------
// arithmetic primitives
// ------------------------------------------------------------------------------------
// STRICT: IEEE round-to-nearest, never contracted -> numpy's ufunc-by-ufunc evaluation.
// FAST:   plain operators (nvcc contracts a*b+c into DFMA) and explicit fma.
template <bool STRICT>
struct Op {
    static __device__ __forceinline__ double add(double a, double b) { return STRICT ? __dadd_rn(a, b) : a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return STRICT ? __dsub_rn(a, b) : a - b; }
    static __device__ __forceinline__ double mul(double a, double b) { return STRICT ? __dmul_rn(a, b) : a * b; }
    static __device__ __forceinline__ double div(double a, double b) { return STRICT ? __ddiv_rn(a, b) : a / b; }
};

// 1/x to ~1 ulp from the hardware seed (MUFU.RCP64H, ~2^-20) and one cubic Newton step:
// r(1 + e + e^2) with e = 1 - x r leaves a relative error e^3.
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}

// sqrt(x) for positive finite x from the hardware seed (MUFU.RSQ64H) and a coupled Newton
// iteration + one residual correction; no special-case branch (FAST mode only).
__device__ __forceinline__ double fast_sqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double s = x * y;
    double hy = 0.5 * y;
    const double e = fma(-s, hy, 0.5);
    s = fma(s, e, s);
    hy = fma(hy, e, hy);
    const double d = fma(-s, s, x);
    return fma(d, hy, s);
}

__device__ __forceinline__ unsigned long long nonneg_bits(double v)
{
    // |v| as an unsigned integer: monotone for finite values and +inf, and every NaN maps
    // above +inf, so an integer max reproduces numpy's NaN-propagating max.  (Integer
    // pipe only: the FP64 pipe is the scarce resource of this kernel.)
    return ((unsigned long long)(unsigned)(__double2hiint(v) & 0x7fffffff) << 32) | (unsigned)__double2loint(v);
}

__device__ __forceinline__ double shfl_down1(double x) { return __shfl_down_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double shfl_up1(double x) { return __shfl_up_sync(0xffffffffu, x, 1); }

// ------------------------------------------------------------------------------------
// per-cell products (numpy.py:185-188, 203, 220, 237, 254-255)
// ------------------------------------------------------------------------------------
struct Cell {
    double h, u, v;
    double hu, hv;   // h*u, h*v
    double Ux, Vx;   // hu*u + (hg*h)*h ; hu*v
    double hw;       // h*(v*c)          ("hv1")
    double Uy, Vy1;  // hu*(v*c) ; hv*(v*c)
    double pr;       // (hg*h)*h         ("Vy2", also the pressure part of Ux)
    double f;        // Coriolis at the cell (only when it is a 2-D field)
};

template <bool STRICT>
__device__ __forceinline__ Cell make_cell(double h, double u, double v, double c, double hg, double f)
{
    using O = Op<STRICT>;
    Cell q;
    q.h = h; q.u = u; q.v = v; q.f = f;
    q.hu = O::mul(h, u);
    q.hv = O::mul(h, v);
    q.pr = O::mul(O::mul(hg, h), h);
    q.Ux = STRICT ? O::add(O::mul(q.hu, u), q.pr) : fma(q.hu, u, q.pr);
    q.Vx = O::mul(q.hu, v);
    const double w = O::mul(v, c);
    q.hw = O::mul(h, w);
    q.Uy = O::mul(q.hu, w);
    q.Vy1 = O::mul(q.hv, w);
    return q;
}

// Face record.  STRICT stores the reference's quantities; FAST stores hM, huM, hvM and the
// fluxes DOUBLED (a factor 2 is exact), with the 0.5 folded into the update coefficients.
struct Face {
    double hM, huM, hvM;
    double q;           // huM / hM, unguarded (numpy.py:302-305)
    double F1, F2, F3;  // x: UxMid, VxMid, - ; y: UyMid, Vy1Mid, Vy2Mid
    double hvc;         // y only: hvMidy * cMidy
};

// x-face between (i, j) = `a` and (i+1, j) = `b`.  numpy.py:193-195, 203-217, 237-251,
// 287-291, 321.
//   STRICT: cdx = hdt/dx[i,j];  kt = tan(theta_j) (tgMidx);  fK = 0.5*(f_b+f_a)
//   FAST:   cdx = dt/dx_j;      kt = hdt*0.5*tan/a;          fK = hdt*f
template <bool STRICT, bool GUARD = true>
__device__ __forceinline__ Face x_face(const Cell &a, const Cell &b, double cdx, double kt, double fK,
                                       double hdt, double hg, double ra)
{
    using O = Op<STRICT>;
    Face F;
    if (STRICT) {
        const double P = O::add(fK, O::div(O::mul(O::mul(0.5, O::add(b.u, a.u)), kt), ra));  // ra = a
        const double K = O::mul(hdt, P);
        const double shu = O::mul(0.5, O::add(b.hu, a.hu));
        const double shv = O::mul(0.5, O::add(b.hv, a.hv));
        F.hM = O::sub(O::mul(0.5, O::add(b.h, a.h)), O::mul(cdx, O::sub(b.hu, a.hu)));
        F.huM = O::add(O::sub(shu, O::mul(cdx, O::sub(b.Ux, a.Ux))), O::mul(K, shv));
        F.hvM = O::sub(O::sub(shv, O::mul(cdx, O::sub(b.Vx, a.Vx))), O::mul(K, shu));
        F.q = O::div(F.huM, F.hM);
        const double pres = O::mul(O::mul(hg, F.hM), F.hM);
        const bool pos = F.hM > 0.0;
        F.F1 = pos ? O::add(O::div(O::mul(F.huM, F.huM), F.hM), pres) : pres;
        F.F2 = pos ? O::div(O::mul(F.hvM, F.huM), F.hM) : 0.0;
    } else {
        const double K = fma(b.u + a.u, kt, fK);
        const double shu = b.hu + a.hu, shv = b.hv + a.hv;
        F.hM = fma(-cdx, b.hu - a.hu, b.h + a.h);             // 2 hMidx
        F.huM = fma(K, shv, fma(-cdx, b.Ux - a.Ux, shu));     // 2 huMidx
        F.hvM = fma(-K, shu, fma(-cdx, b.Vx - a.Vx, shv));    // 2 hvMidx
        const double r = fast_rcp(F.hM);
        F.q = F.huM * r;
        const double pres = (hg * F.hM) * F.hM;               // hg here is g/4: 2*(g/2)*hM^2
        const bool pos = !GUARD || __double2hiint(F.hM) > 0 || (__double2hiint(F.hM) == 0 && __double2loint(F.hM) != 0);
        F.F1 = pos ? fma(F.huM, F.q, pres) : pres;            // 2 UxMid
        F.F2 = pos ? F.hvM * F.q : 0.0;                       // 2 VxMid
    }
    F.F3 = 0.0;
    F.hvc = 0.0;
    return F;
}

// y-face between (i, j) = `a` and (i, j+1) = `b`.  numpy.py:198-200, 220-234, 254-270,
// 292, 322-323.
//   STRICT: c1 = hdt/dy1_j, c2 = hdt/dy_j, kt = tgMidy_j, fK = 0.5*(f_b+f_a)
//   FAST:   c1 = dt/dy1_j,  c2 = dt/dy_j,  kt = hdt*0.5*tgMidy/a, fK = hdt*0.5*(f_b+f_a)
template <bool STRICT, bool GUARD = true>
__device__ __forceinline__ Face y_face(const Cell &a, const Cell &b, double c1, double c2, double kt, double fK,
                                       double cm, double hdt, double hg, double ra)
{
    using O = Op<STRICT>;
    Face F;
    if (STRICT) {
        const double P = O::add(fK, O::div(O::mul(O::mul(0.5, O::add(b.u, a.u)), kt), ra));
        const double K = O::mul(hdt, P);
        const double shu = O::mul(0.5, O::add(b.hu, a.hu));
        const double shv = O::mul(0.5, O::add(b.hv, a.hv));
        F.hM = O::sub(O::mul(0.5, O::add(b.h, a.h)), O::mul(c1, O::sub(b.hw, a.hw)));
        F.huM = O::add(O::sub(shu, O::mul(c1, O::sub(b.Uy, a.Uy))), O::mul(K, shv));
        F.hvM = O::sub(O::sub(O::sub(shv, O::mul(c1, O::sub(b.Vy1, a.Vy1))), O::mul(c2, O::sub(b.pr, a.pr))),
                       O::mul(K, shu));
        F.q = O::div(F.huM, F.hM);
        F.hvc = O::mul(F.hvM, cm);
        const bool pos = F.hM > 0.0;
        F.F1 = pos ? O::div(O::mul(F.hvc, F.huM), F.hM) : 0.0;
        F.F2 = pos ? O::mul(O::div(O::mul(F.hvM, F.hvM), F.hM), cm) : 0.0;
        F.F3 = O::mul(O::mul(hg, F.hM), F.hM);
    } else {
        const double K = fma(b.u + a.u, kt, fK);
        const double shu = b.hu + a.hu, shv = b.hv + a.hv;
        F.hM = fma(-c1, b.hw - a.hw, b.h + a.h);
        F.huM = fma(K, shv, fma(-c1, b.Uy - a.Uy, shu));
        F.hvM = fma(-K, shu, fma(-c2, b.pr - a.pr, fma(-c1, b.Vy1 - a.Vy1, shv)));
        const double r = fast_rcp(F.hM);
        F.q = F.huM * r;
        F.hvc = F.hvM * cm;
        const bool pos = !GUARD || __double2hiint(F.hM) > 0 || (__double2hiint(F.hM) == 0 && __double2loint(F.hM) != 0);
        F.F1 = pos ? F.hvc * F.q : 0.0;            // 2 UyMid
        F.F2 = pos ? F.hvc * (F.hvM * r) : 0.0;    // 2 Vy1Mid
        F.F3 = (hg * F.hM) * F.hM;                 // 2 Vy2Mid (hg = g/4)
    }
    return F;
}

// 3-point first-derivative weights on spacings d0 (minus side), d1 (plus side):
// numpy.py:32-34, 37-39, 42-44.
__device__ __forceinline__ void weights_strict(double d0, double d1, double &A, double &B, double &C)
{
    using O = Op<true>;
    A = O::div(O::sub(d1, d0), O::mul(d1, d0));
    B = O::div(d0, O::mul(d1, O::add(d1, d0)));
    C = O::div(-d1, O::mul(d0, O::add(d1, d0)));
}

__device__ __forceinline__ double second_diff_strict(double Am, double Bm, double Cm, double A0, double B0, double C0,
                                                     double Ap, double Bp, double Cp, double qmm, double qm, double q0,
                                                     double qp, double qpp)
{
    using O = Op<true>;
    // numpy.py:87-105: A*(A q + B q+ + C q-) + B*(A+ q+ + B+ q++ + C+ q) + C*(A- q- + B- q + C- q--)
    const double t0 = O::mul(A0, O::add(O::add(O::mul(A0, q0), O::mul(B0, qp)), O::mul(C0, qm)));
    const double t1 = O::mul(B0, O::add(O::add(O::mul(Ap, qp), O::mul(Bp, qpp)), O::mul(Cp, q0)));
    const double t2 = O::mul(C0, O::add(O::add(O::mul(Am, qm), O::mul(Bm, q0)), O::mul(Cm, qmm)));
    return O::add(O::add(t0, t1), t2);
}

// Everything a lane needs to evaluate the Laplacian (numpy.py:66-133 on the extension of
// numpy.py:376-381) at its latitude.
struct LapCtx {
    // y weights of rows j-1, j, j+1 (strict) or the five combined weights (fast)
    double Aym, Bym, Cym, Ay0, By0, Cy0, Ayp, Byp, Cyp;
    double w[5];
    double bx2;             // fast: 1 / (2 dx_j)^2
    int jm2, jm1, jp1, jp2;  // clamped neighbour columns (local indices)
};

template <bool STRICT>
__device__ __forceinline__ double laplacian_at(const double *__restrict__ q, const Params &p, const LapCtx &L,
                                               double ac, int i, int jl)
{
    const int M = p.M;
    const size_t P = (size_t)p.pitch;
    const int im2 = (i - 2 < 0) ? M - 1 : i - 2;  // numpy.py:376-378
    const int ip2 = (i + 2 > M + 2) ? 3 : i + 2;
    const double q0 = q[(size_t)i * P + jl];
    const double qxm = q[(size_t)(i - 1) * P + jl], qxp = q[(size_t)(i + 1) * P + jl];
    const double qxmm = q[(size_t)im2 * P + jl], qxpp = q[(size_t)ip2 * P + jl];
    const double qym = q[(size_t)i * P + L.jm1], qyp = q[(size_t)i * P + L.jp1];
    const double qymm = q[(size_t)i * P + L.jm2], qypp = q[(size_t)i * P + L.jp2];
    if (STRICT) {
        using O = Op<true>;
        double A[3], B[3], C[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            int k = i - 1 + s;
            k = (k == 0) ? M : ((k == M + 2) ? 2 : k);  // numpy.py:35, 40, 45
            const double xm = O::mul(ac, p.phi[k - 1]), x0 = O::mul(ac, p.phi[k]), xp = O::mul(ac, p.phi[k + 1]);
            weights_strict(O::sub(x0, xm), O::sub(xp, x0), A[s], B[s], C[s]);
        }
        const double qxx = second_diff_strict(A[0], B[0], C[0], A[1], B[1], C[1], A[2], B[2], C[2], qxmm, qxm, q0, qxp, qxpp);
        const double qyy = second_diff_strict(L.Aym, L.Bym, L.Cym, L.Ay0, L.By0, L.Cy0, L.Ayp, L.Byp, L.Cyp, qymm, qym, q0, qyp, qypp);
        return O::add(qxx, qyy);
    } else {
        const double qxx = L.bx2 * fma(-2.0, q0, qxpp + qxmm);
        const double qyy = fma(L.w[0], qymm, fma(L.w[1], qym, fma(L.w[2], q0, fma(L.w[3], qyp, L.w[4] * qypp))));
        return qxx + qyy;
    }
}

// ------------------------------------------------------------------------------------
// step prologue: dt from the device-resident CFL maxima (numpy.py:524-532)
// ------------------------------------------------------------------------------------
// Executed by thread 0 of every block (every block needs dt); block (0,0) also advances
// the loop state.  Returns through shared memory: dt and flags (bit 0 active, bit 1 parity).
__device__ __forceinline__ void step_prologue(const Params &p, double *s_dt, int *s_flags)
{
    Ctrl *const cur_slot = p.ctrl + (p.launch % 3);
    Ctrl *const new_slot = p.ctrl + ((p.launch + 1) % 3);
    if (p.fixed) {
        *s_dt = p.fixed_dt;
        *s_flags = 1;
        return;
    }
    const double time = cur_slot->time;
    const long long nsteps = cur_slot->nsteps;
    const unsigned long long exb = cur_slot->ex_bits, eyb = cur_slot->ey_bits;
    const double tx = __ddiv_rn(p.dxmin, __longlong_as_double((long long)exb));
    const double ty = __ddiv_rn(p.dymin, __longlong_as_double((long long)eyb));
    const double dtmax = (tx != tx) ? tx : ((ty != ty) ? ty : (tx < ty ? tx : ty));  // np.minimum propagates NaN
    double dt = __dmul_rn(p.courant, dtmax);
    const bool active = time < p.final_time;  // numpy.py:496 (false for NaN)
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
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        Ctrl *const clr_slot = p.ctrl + ((p.launch + 2) % 3);
        clr_slot->ex_bits = 0ULL;
        clr_slot->ey_bits = 0ULL;
        new_slot->time = tnew;
        new_slot->nsteps = nsteps + (active ? 1 : 0);
        new_slot->current = (int)((nsteps + (active ? 1 : 0)) & 1);
        new_slot->last_dt = active ? dt : cur_slot->last_dt;
        new_slot->done = !(tnew < p.final_time);
        new_slot->error = cur_slot->error;
        if (!active) {  // nothing will be accumulated: carry the maxima forward
            new_slot->ex_bits = exb;
            new_slot->ey_bits = eyb;
        } else if (p.log != nullptr && nsteps < (long long)p.log_cap) {
            p.log[nsteps] = dt;
            p.log[(size_t)p.log_cap + nsteps] = tnew;
        }
    }
}

// ------------------------------------------------------------------------------------
// one lane of the march: per-latitude constants + the per-row work
// ------------------------------------------------------------------------------------
template <bool STRICT, bool DIFF, bool F2D, bool HS>
struct Lane {
    using O = Op<STRICT>;
    int jl;                   // local latitude row of this lane (clamped into the array)
    double dt, hdt, hgCell, dnu;
    double c_j, ac_j;
    double hgK, raK;          // hg (FAST: g/4), a (FAST: 1/a)
    double kx, fKx;           // x-face Coriolis / metric factors
    double ky, fKy, cm_j;     // y-face
    double cy1f, cyf;         // y-face flux coefficients
    double cxu, cy1u, cyu;    // update coefficients (cxu: FAST only; STRICT forms dt/dxc per cell)
    double kc, fD;            // update Coriolis / metric
    double cdx_fast;
    double dy1s;              // HS: dy1[j-1] + dy1[j]
    double x_m, x_c;          // STRICT: x[i-1, j], x[i, j]
    LapCtx L;
    unsigned long long mx, my;  // running CFL maxima (bit patterns)

    __device__ __forceinline__ void init(const Params &p, int jl_, double dt_)
    {
        jl = jl_;
        dt = dt_;
        mx = my = 0ULL;
        const double *tab = p.tab + jl;
        const size_t TS = (size_t)p.ncols;
        c_j = tab[T_C * TS];
        const double tg_j = tab[T_TG * TS];
        ac_j = tab[T_AC * TS];
        const double f_j = F2D ? 0.0 : tab[T_F * TS];
        const double f_jp = F2D ? 0.0 : p.tab[T_F * TS + min(jl + 1, p.jl_hi)];
        cm_j = tab[T_CMIDY * TS];
        hdt = O::mul(0.5, dt);
        hgCell = O::mul(0.5, p.g);  // numpy.py:203, 255: 0.5*g
        dnu = O::mul(dt, p.nu);
        cxu = 0.0;
        cdx_fast = 0.0;
        x_m = x_c = 0.0;
        if (STRICT) {
            hgK = O::mul(0.5, p.g);
            raK = p.a;
            kx = tg_j;  // tgMidx == tan(theta_j): 0.5*(theta+theta) is exact (grid.py:107)
            fKx = f_j;  // 0.5*(f+f) is exact
            ky = tab[T_TGMIDY * TS];
            fKy = O::mul(0.5, O::add(f_jp, f_j));
            cy1f = O::div(hdt, tab[T_DY1 * TS]);
            cyf = O::div(hdt, tab[T_DY * TS]);
            cy1u = O::div(dt, tab[T_DY1C * TS]);
            cyu = O::div(dt, tab[T_DYC * TS]);
            kc = tg_j;
            fD = f_j;
        } else {
            const double ra = 1.0 / p.a;
            const double dxa = ac_j * p.dphi;  // analytic dx_j (the reference's differs by roundoff only)
            hgK = 0.25 * p.g;
            raK = ra;
            kx = hdt * 0.5 * tg_j * ra;
            fKx = hdt * f_j;
            ky = hdt * 0.5 * tab[T_TGMIDY * TS] * ra;
            fKy = hdt * 0.5 * (f_jp + f_j);
            cdx_fast = dt / dxa;
            cy1f = dt / tab[T_DY1 * TS];
            cyf = dt / tab[T_DY * TS];
            cxu = 0.5 * dt / dxa;
            cy1u = 0.5 * dt / tab[T_DY1C * TS];
            cyu = 0.5 * dt / tab[T_DYC * TS];
            kc = dt * 0.25 * 0.5 * 0.25 * tg_j * ra;  // (dt*0.25) * 0.5[doubled sums] * 0.25 * tg/a
            fD = dt * 0.25 * 0.5 * f_j;
        }
        dy1s = HS ? O::add(p.tab[T_DY1 * TS + max(jl - 1, p.jl_lo)], tab[T_DY1 * TS]) : 1.0;
        if (DIFF) {
            const int jm = max(jl - 1, p.jl_lo), jp = min(jl + 1, p.jl_hi);
            L.Aym = p.tab[T_AY * TS + jm]; L.Bym = p.tab[T_BY * TS + jm]; L.Cym = p.tab[T_CY * TS + jm];
            L.Ay0 = tab[T_AY * TS]; L.By0 = tab[T_BY * TS]; L.Cy0 = tab[T_CY * TS];
            L.Ayp = p.tab[T_AY * TS + jp]; L.Byp = p.tab[T_BY * TS + jp]; L.Cyp = p.tab[T_CY * TS + jp];
            // numpy.py:379-381: one duplicated latitude each side (only at the pole rows;
            // a band interior has real halo rows there)
            L.jm1 = jm;
            L.jp1 = jp;
            L.jm2 = max(jl - 2, p.jl_lo);
            L.jp2 = min(jl + 2, p.jl_hi);
            if (!STRICT) {
                L.w[0] = L.Cy0 * L.Cym;
                L.w[1] = L.Ay0 * L.Cy0 + L.Cy0 * L.Aym;
                L.w[2] = L.Ay0 * L.Ay0 + L.By0 * L.Cyp + L.Cy0 * L.Bym;
                L.w[3] = L.Ay0 * L.By0 + L.By0 * L.Ayp;
                L.w[4] = L.By0 * L.Byp;
                const double dxa = ac_j * p.dphi;
                L.bx2 = 0.25 / (dxa * dxa);
            }
        }
    }

    __device__ __forceinline__ Cell cell(double h, double u, double v, double f) const
    {
        return make_cell<STRICT>(h, u, v, c_j, hgCell, f);
    }

    // face (ia-1 | ia) that starts a march
    __device__ __forceinline__ Face first_face(const Params &p, const Cell &prev, const Cell &cur, int ia)
    {
        if (STRICT) {
            x_m = O::mul(ac_j, p.phi[ia - 1]);
            x_c = O::mul(ac_j, p.phi[ia]);
            const double cdx = O::div(hdt, O::sub(x_c, x_m));
            return x_face<true>(prev, cur, cdx, kx, F2D ? O::mul(0.5, O::add(cur.f, prev.f)) : fKx, hdt, hgK, raK);
        }
        return x_face<false>(prev, cur, cdx_fast, kx, F2D ? hdt * 0.5 * (cur.f + prev.f) : fKx, hdt, hgK, raK);
    }

    // Update cell (i, j) given its own products `cur`, the next longitude `nxt` and the
    // left face `x0`; returns the right face (the next iteration's left face).
    // numpy.py:193-354 (+541-544 with DIFF).  `count` says whether this lane's result is a
    // real cell (takes part in the CFL maxima).
    __device__ __forceinline__ Face advance(const Params &p, const double *__restrict__ H, const double *__restrict__ U,
                                            const double *__restrict__ V, const Cell &cur, const Cell &nxt,
                                            const Face &x0, int i, bool count, double &hn, double &un, double &vn)
    {
        // x-face (i | i+1)
        Face x1;
        double dx0 = 0.0, dx1 = 0.0;
        if (STRICT) {
            const double x_p = O::mul(ac_j, p.phi[i + 1]);
            dx0 = O::sub(x_c, x_m);
            dx1 = O::sub(x_p, x_c);
            x_m = x_c;
            x_c = x_p;
            x1 = x_face<true>(cur, nxt, O::div(hdt, dx1), kx, F2D ? O::mul(0.5, O::add(nxt.f, cur.f)) : fKx, hdt, hgK, raK);
        } else {
            x1 = x_face<false>(cur, nxt, cdx_fast, kx, F2D ? hdt * 0.5 * (nxt.f + cur.f) : fKx, hdt, hgK, raK);
        }

        // y-face (j | j+1): the northern cell comes from lane+1
        Cell up;
        up.h = shfl_down1(cur.h);   up.u = shfl_down1(cur.u);
        up.hu = shfl_down1(cur.hu); up.hv = shfl_down1(cur.hv);
        up.hw = shfl_down1(cur.hw); up.Uy = shfl_down1(cur.Uy);
        up.Vy1 = shfl_down1(cur.Vy1); up.pr = shfl_down1(cur.pr);
        double fy = fKy;
        if (F2D) {
            const double fup = shfl_down1(cur.f);
            fy = STRICT ? O::mul(0.5, O::add(fup, cur.f)) : hdt * 0.5 * (fup + cur.f);
        }
        const Face y1 = y_face<STRICT>(cur, up, cy1f, cyf, ky, fy, cm_j, hdt, hgK, raK);

        // y-face (j-1 | j) was computed by lane-1
        Face y0;
        y0.huM = shfl_up1(y1.huM); y0.hvM = shfl_up1(y1.hvM); y0.q = shfl_up1(y1.q);
        y0.F1 = shfl_up1(y1.F1);   y0.F2 = shfl_up1(y1.F2);   y0.F3 = shfl_up1(y1.F3);
        y0.hvc = shfl_up1(y1.hvc);
        y0.hM = HS ? shfl_up1(y1.hM) : 0.0;

        // full-step update of cell (i, j): numpy.py:275-354
        const size_t P = (size_t)p.pitch;
        if (STRICT) {
            const double cx = O::div(dt, O::mul(0.5, O::add(dx0, dx1)));
            hn = O::sub(O::sub(cur.h, O::mul(cx, O::sub(x1.huM, x0.huM))), O::mul(cy1u, O::sub(y1.hvc, y0.hvc)));
            const double qs = O::add(O::add(O::add(x0.q, x1.q), y0.q), y1.q);
            const double fc = F2D ? cur.f : fD;
            const double Fc = O::add(fc, O::div(O::mul(O::mul(0.25, qs), kc), raK));
            const double D = O::mul(O::mul(dt, Fc), 0.25);
            const double shv = O::add(O::add(O::add(x0.hvM, x1.hvM), y0.hvM), y1.hvM);
            const double shu = O::add(O::add(O::add(x0.huM, x1.huM), y0.huM), y1.huM);
            double hun = O::add(O::sub(O::sub(cur.hu, O::mul(cx, O::sub(x1.F1, x0.F1))), O::mul(cy1u, O::sub(y1.F1, y0.F1))),
                                O::mul(D, shv));
            double hvn = O::sub(O::sub(O::sub(O::sub(cur.hv, O::mul(cx, O::sub(x1.F2, x0.F2))), O::mul(cy1u, O::sub(y1.F2, y0.F2))),
                                       O::mul(cyu, O::sub(y1.F3, y0.F3))),
                                O::mul(D, shu));
            if (HS) {  // numpy.py:310-317, 342-349
                const double hsum = O::add(O::add(O::add(x0.hM, x1.hM), y0.hM), y1.hM);
                const double gq = O::mul(O::mul(O::mul(dt, p.g), 0.25), hsum);
                const size_t k = (size_t)i * P + jl;
                const double dhx = O::sub(p.hs[k + P], p.hs[k - P]);
                const double dhy = O::sub(p.hs[(size_t)i * P + min(jl + 1, p.jl_hi)], p.hs[(size_t)i * P + max(jl - 1, p.jl_lo)]);
                hun = O::sub(hun, O::div(O::mul(gq, dhx), O::add(dx0, dx1)));
                hvn = O::sub(hvn, O::div(O::mul(gq, dhy), dy1s));
            }
            un = O::div(hun, hn);
            vn = O::div(hvn, hn);
        } else {
            hn = fma(-cy1u, y1.hvc - y0.hvc, fma(-cxu, x1.huM - x0.huM, cur.h));
            const double qs = (x0.q + x1.q) + (y0.q + y1.q);
            const double D = F2D ? fma(qs, kc, dt * 0.125 * cur.f) : fma(qs, kc, fD);  // 0.5 * dt * Fc * 0.25
            const double shv = (x0.hvM + x1.hvM) + (y0.hvM + y1.hvM);
            const double shu = (x0.huM + x1.huM) + (y0.huM + y1.huM);
            double hun = fma(D, shv, fma(-cy1u, y1.F1 - y0.F1, fma(-cxu, x1.F1 - x0.F1, cur.hu)));
            double hvn = fma(-D, shu,
                             fma(-cyu, y1.F3 - y0.F3, fma(-cy1u, y1.F2 - y0.F2, fma(-cxu, x1.F2 - x0.F2, cur.hv))));
            if (HS) {
                const double hsum = (x0.hM + x1.hM) + (y0.hM + y1.hM);  // doubled
                const double gq = dt * p.g * 0.125 * hsum;
                const size_t k = (size_t)i * P + jl;
                const double dhx = p.hs[k + P] - p.hs[k - P];
                const double dhy = p.hs[(size_t)i * P + min(jl + 1, p.jl_hi)] - p.hs[(size_t)i * P + max(jl - 1, p.jl_lo)];
                hun -= gq * dhx / (2.0 * ac_j * p.dphi);
                hvn -= gq * dhy / dy1s;
            }
            const double r = fast_rcp(hn);
            un = hun * r;
            vn = hvn * r;
        }
        if (DIFF) {  // numpy.py:541-544: old fields; u, v (not hu, hv)
            const double lh = laplacian_at<STRICT>(H, p, L, ac_j, i, jl);
            const double lu = laplacian_at<STRICT>(U, p, L, ac_j, i, jl);
            const double lv = laplacian_at<STRICT>(V, p, L, ac_j, i, jl);
            if (STRICT) {
                hn = O::add(hn, O::mul(dnu, lh));
                un = O::add(un, O::mul(dnu, lu));
                vn = O::add(vn, O::mul(dnu, lv));
            } else {
                hn = fma(dnu, lh, hn);
                un = fma(dnu, lu, un);
                vn = fma(dnu, lv, vn);
            }
        }
        if (count) {
            // CFL maxima of the new state: max(|q-c|, |q|, |q+c|) == |q| + c for c >= 0
            const double cs = STRICT ? __dsqrt_rn(O::mul(p.g, fabs(hn))) : sqrt(p.g * fabs(hn));
            const unsigned long long bx = nonneg_bits(O::add(fabs(un), cs));
            const unsigned long long by = nonneg_bits(O::add(fabs(vn), cs));
            mx = bx > mx ? bx : mx;
            my = by > my ? by : my;
        }
        return x1;
    }

    // Branch-free form of `advance` for the hot configuration (FAST arithmetic, latitude-only
    // Coriolis, flat topography, no diffusion).  Assumes positive layer depths: the
    // reference's `hMid > 0` selects (numpy.py:287-292, 321-322) always take their first
    // branch there, so they are not evaluated.  The CFL maxima are accumulated for every lane
    // and row; lanes / rows that are not real cells are masked when published.
    __device__ __forceinline__ Face advance_fast(const Cell &cur, const Cell &nxt, const Face &x0, double g,
                                                 double &hn, double &un, double &vn)
    {
        static_assert(!STRICT && !DIFF && !F2D && !HS, "advance_fast serves the plain FAST variant");
        const Face x1 = x_face<false, false>(cur, nxt, cdx_fast, kx, fKx, hdt, hgK, raK);
        Cell up;
        up.h = shfl_down1(cur.h);   up.u = shfl_down1(cur.u);
        up.hu = shfl_down1(cur.hu); up.hv = shfl_down1(cur.hv);
        up.hw = shfl_down1(cur.hw); up.Uy = shfl_down1(cur.Uy);
        up.Vy1 = shfl_down1(cur.Vy1); up.pr = shfl_down1(cur.pr);
        const Face y1 = y_face<false, false>(cur, up, cy1f, cyf, ky, fKy, cm_j, hdt, hgK, raK);
        Face y0;
        y0.huM = shfl_up1(y1.huM); y0.hvM = shfl_up1(y1.hvM); y0.q = shfl_up1(y1.q);
        y0.F1 = shfl_up1(y1.F1);   y0.F2 = shfl_up1(y1.F2);   y0.F3 = shfl_up1(y1.F3);
        y0.hvc = shfl_up1(y1.hvc);

        hn = fma(-cy1u, y1.hvc - y0.hvc, fma(-cxu, x1.huM - x0.huM, cur.h));
        const double qs = (x0.q + x1.q) + (y0.q + y1.q);
        const double D = fma(qs, kc, fD);  // 0.5 * dt * Fc * 0.25
        const double shv = (x0.hvM + x1.hvM) + (y0.hvM + y1.hvM);
        const double shu = (x0.huM + x1.huM) + (y0.huM + y1.huM);
        const double hun = fma(D, shv, fma(-cy1u, y1.F1 - y0.F1, fma(-cxu, x1.F1 - x0.F1, cur.hu)));
        const double hvn = fma(-D, shu, fma(-cyu, y1.F3 - y0.F3, fma(-cy1u, y1.F2 - y0.F2, fma(-cxu, x1.F2 - x0.F2, cur.hv))));
        const double r = fast_rcp(hn);
        un = hun * r;
        vn = hvn * r;
        // |q| + sqrt(g |h|) is non-negative (or NaN, which compares above everything as an
        // unsigned pattern whatever its sign bit): no masking needed before the integer max
        const double cs = fast_sqrt(g * fabs(hn));
        const unsigned long long bx = (unsigned long long)__double_as_longlong(fabs(un) + cs);
        const unsigned long long by = (unsigned long long)__double_as_longlong(fabs(vn) + cs);
        mx = bx > mx ? bx : mx;
        my = by > my ? by : my;
        return x1;
    }

    // warp-reduce the running maxima and publish them (one atomic pair per warp)
    __device__ __forceinline__ void publish(const Params &p, int lane, bool valid_lane = true)
    {
        if (p.fixed) return;
        if (!valid_lane) mx = my = 0ULL;
        Ctrl *const new_slot = p.ctrl + ((p.launch + 1) % 3);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long ox = __shfl_xor_sync(0xffffffffu, mx, s);
            const unsigned long long oy = __shfl_xor_sync(0xffffffffu, my, s);
            mx = ox > mx ? ox : mx;
            my = oy > my ? oy : my;
        }
        if (lane == 0) {
            atomicMax(&new_slot->ex_bits, mx);
            atomicMax(&new_slot->ey_bits, my);
        }
    }
};

// Extra copies of a freshly computed cell: periodic wrap in longitude and zero-gradient
// pole rows (numpy.py:402-404).  `main` says whether (i, jl) itself is still to be written.
__device__ __forceinline__ void put_cell(const Params &p, double *__restrict__ Hn, double *__restrict__ Un,
                                         double *__restrict__ Vn, int i, int jl, double hn, double un, double vn, bool main)
{
    const size_t P = (size_t)p.pitch;
    const bool south = p.pole_s && (jl == p.jl_lo + 1);
    const bool north = p.pole_n && (jl == p.jl_hi - 1);
    auto put = [&](int row, bool centre) {
        const size_t k = (size_t)row * P + jl;
        if (centre) { Hn[k] = hn; Un[k] = un; Vn[k] = vn; }
        if (south) { Hn[k - 1] = hn; Un[k - 1] = un; Vn[k - 1] = vn; }
        if (north) { Hn[k + 1] = hn; Un[k + 1] = un; Vn[k + 1] = vn; }
    };
    if (main || south || north) put(i, main);
    if (i == p.M) put(0, true);
    if (i == 2) put(p.M + 2, true);
}

// ------------------------------------------------------------------------------------
// the fused step kernel, direct-load form (all variants)
// ------------------------------------------------------------------------------------
template <bool STRICT, bool DIFF, bool F2D, bool HS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) step_kernel(const __grid_constant__ Params p)
{
    __shared__ double s_dt;
    __shared__ int s_flags;
    if (threadIdx.x == 0) step_prologue(p, &s_dt, &s_flags);
    __syncthreads();
    const int flags = s_flags;
    if (!(flags & 1)) return;
    const int par = (flags >> 1) & 1;

    const double *__restrict__ H = p.buf[par][0];
    const double *__restrict__ U = p.buf[par][1];
    const double *__restrict__ V = p.buf[par][2];
    double *__restrict__ Hn = p.buf[par ^ 1][0];
    double *__restrict__ Un = p.buf[par ^ 1][1];
    double *__restrict__ Vn = p.buf[par ^ 1][2];

    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (strip >= p.nstrips) return;

    const size_t P = (size_t)p.pitch;
    const int jl_raw = p.jl_first - 1 + kCellsPerWarp * strip + lane;
    const int jl = min(jl_raw, p.jl_hi);  // lanes past the array replay the last row
    const bool upd = (lane >= 1) && (lane <= kCellsPerWarp) && (jl_raw <= p.jl_last);
    const int ia = 1 + blockIdx.y * p.chunk;     // first longitude updated
    const int ib = min(ia + p.chunk, p.M + 2);   // one past the last

    Lane<STRICT, DIFF, F2D, HS> ln;
    ln.init(p, jl, s_dt);
    auto load_cell = [&](int i) {
        const size_t k = (size_t)i * P + jl;
        return ln.cell(H[k], U[k], V[k], F2D ? p.f2d[k] : 0.0);
    };
    Cell cur;
    Face x0;
    {
        const Cell prev = load_cell(ia - 1);
        cur = load_cell(ia);
        x0 = ln.first_face(p, prev, cur, ia);
    }
    for (int i = ia; i < ib; ++i) {
        const Cell nxt = load_cell(i + 1);
        double hn, un, vn;
        x0 = ln.advance(p, H, U, V, cur, nxt, x0, i, upd, hn, un, vn);
        if (upd) put_cell(p, Hn, Un, Vn, i, jl, hn, un, vn, true);
        cur = nxt;
    }
    ln.publish(p, lane);
}

// ------------------------------------------------------------------------------------
// the fused step kernel, TMA form (flat topography, latitude-only Coriolis, no diffusion)
// ------------------------------------------------------------------------------------
// Every warp runs its own asynchronous pipeline: lane 0 issues cp.async.bulk.tensor loads
// of {34 latitudes x R longitudes} boxes of h, u, v into a ring of S shared-memory stages
// (completion on one mbarrier per stage) and bulk tensor stores of the {30 x R} result
// boxes; lanes read / write shared memory only.  Loads of the rows far ahead of the march
// are in flight while the FP64 pipe works, without costing registers.
struct TmaMaps {
    CUtensorMap in[2][3];   // box {34, R} over (ncols, M+3)
    CUtensorMap out[2][3];  // box {30, R} over (jl_last+1, M+2): clips garbage rows / halo columns
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((unsigned long long)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

template <int R, int S>
struct TmaLayout {
    static constexpr int kInCols = 34;  // lane 0's column is odd; boxes must start on 16-byte (even) columns
    static constexpr int kInTile = ((R * kInCols * 8 + 127) / 128) * 128;  // one field, one stage
    static constexpr int kOutTile = ((R * 30 * 8 + 127) / 128) * 128;
    static constexpr int kIn = S * 3 * kInTile;
    static constexpr int kOut = 2 * 3 * kOutTile;
    static constexpr int kBars = 128;
    static constexpr int kWarpBytes = kIn + kOut + kBars;
    static constexpr int kBlockBytes = kWarpsPerBlock * kWarpBytes;
};

template <bool STRICT, int R, int S>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, STRICT ? 3 : 4) step_kernel_tma(const __grid_constant__ Params p,
                                                                       const __grid_constant__ TmaMaps tm)
{
    using LY = TmaLayout<R, S>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ double s_dt;
    __shared__ int s_flags;
    if (threadIdx.x == 0) step_prologue(p, &s_dt, &s_flags);
    __syncthreads();
    const int flags = s_flags;
    if (!(flags & 1)) return;
    const int par = (flags >> 1) & 1;

    const double *__restrict__ H = p.buf[par][0];
    const double *__restrict__ U = p.buf[par][1];
    const double *__restrict__ V = p.buf[par][2];
    double *__restrict__ Hn = p.buf[par ^ 1][0];
    double *__restrict__ Un = p.buf[par ^ 1][1];
    double *__restrict__ Vn = p.buf[par ^ 1][2];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * kWarpsPerBlock + warp;
    if (strip >= p.nstrips) return;

    const size_t P = (size_t)p.pitch;
    const int c0 = p.jl_first - 1 + kCellsPerWarp * strip;  // latitude held by lane 0
    const int jl_raw = c0 + lane;
    const int jl = min(jl_raw, p.jl_hi);
    const bool upd = (lane >= 1) && (lane <= kCellsPerWarp) && (jl_raw <= p.jl_last);
    const int ia = 1 + blockIdx.y * p.chunk;
    const int ib = min(ia + p.chunk, p.M + 2);
    const int ntiles = (ib - ia + R - 1) / R;

    unsigned char *wbase = smem_raw + (size_t)warp * LY::kWarpBytes;
    const double *in_s = reinterpret_cast<const double *>(wbase);
    double *out_s = reinterpret_cast<double *>(wbase + LY::kIn);
    const uint32_t in_a = smem_u32(wbase), out_a = in_a + LY::kIn, bar_a = out_a + LY::kOut;

    const CUtensorMap *min_ = tm.in[par];
    const CUtensorMap *mout = tm.out[par ^ 1];

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(bar_a + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue_load = [&](int t) {  // lane 0 only: rows ia+1+tR .. of h, u, v into stage t % S
        const int s = t % S;
        const uint32_t bar = bar_a + 8 * s;
        mbar_expect_tx(bar, 3 * R * LY::kInCols * 8);
#pragma unroll
        for (int k = 0; k < 3; ++k) tma_load_2d(in_a + (s * 3 + k) * LY::kInTile, &min_[k], c0 - 1, ia + 1 + t * R, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < S - 1; ++t)
            if (t < ntiles) issue_load(t);
    }

    Lane<STRICT, false, false, false> ln;
    ln.init(p, jl, s_dt);
    Cell cur;
    Face x0;
    {
        const size_t k0 = (size_t)(ia - 1) * P + jl, k1 = k0 + P;
        const Cell prev = ln.cell(H[k0], U[k0], V[k0], 0.0);
        cur = ln.cell(H[k1], U[k1], V[k1], 0.0);
        x0 = ln.first_face(p, prev, cur, ia);
    }
    const bool edge = p.pole_s && (jl == p.jl_lo + 1) || p.pole_n && (jl == p.jl_hi - 1);
    const bool cell_lane = (lane >= 1) && (lane <= kCellsPerWarp);
    const bool warp_has_edge = __any_sync(0xffffffffu, edge && upd);
    constexpr int kOutF = LY::kOutTile / 8;  // doubles per output field tile

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        mbar_wait(bar_a + 8 * s, (uint32_t)((t / S) & 1));
        const int ob = t & 1;
        if (t >= 2) {  // the store that last read out buffer `ob` must have drained it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        const double *ti = in_s + (size_t)s * 3 * (LY::kInTile / 8) + lane + 1;
        double *to = out_s + (size_t)ob * 3 * kOutF + (lane - 1);
        const int i0 = ia + t * R;
        const bool full = (i0 + R <= ib);  // warp-uniform
        bool done_fast = false;
        if constexpr (!STRICT) {
          if (full) {
            done_fast = true;
            // hot path: R rows in one branch-free basic block
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const Cell nxt = ln.cell(ti[rr * LY::kInCols], ti[LY::kInTile / 8 + rr * LY::kInCols],
                                         ti[2 * (LY::kInTile / 8) + rr * LY::kInCols], 0.0);
                double hn, un, vn;
                x0 = ln.advance_fast(cur, nxt, x0, p.g, hn, un, vn);
                if (cell_lane) {
                    to[rr * 30] = hn;
                    to[kOutF + rr * 30] = un;
                    to[2 * kOutF + rr * 30] = vn;
                }
                cur = nxt;
            }
          }
        }
        if (!done_fast) {
#pragma unroll 1
            for (int rr = 0; rr < R; ++rr) {
                const int i = i0 + rr;
                const Cell nxt = ln.cell(ti[rr * LY::kInCols], ti[LY::kInTile / 8 + rr * LY::kInCols],
                                         ti[2 * (LY::kInTile / 8) + rr * LY::kInCols], 0.0);
                double hn, un, vn;
                const bool live = upd && (i < ib);
                x0 = ln.advance(p, H, U, V, cur, nxt, x0, i, live, hn, un, vn);
                if (cell_lane) {
                    to[rr * 30] = hn;
                    to[kOutF + rr * 30] = un;
                    to[2 * kOutF + rr * 30] = vn;
                }
                cur = nxt;
            }
        }
        // periodic-wrap rows and pole copies (numpy.py:402-404): rare, read back from the tile
        const bool wrap_tile = (i0 <= p.M && p.M < i0 + R) || (i0 <= 2 && 2 < i0 + R);
        if (warp_has_edge || wrap_tile) {
#pragma unroll 1
            for (int rr = 0; rr < R; ++rr) {
                const int i = i0 + rr;
                if (upd && i < ib && (edge || i == p.M || i == 2))
                    put_cell(p, Hn, Un, Vn, i, jl, to[rr * 30], to[kOutF + rr * 30], to[2 * kOutF + rr * 30], false);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) tma_store_2d(&mout[k], out_a + (ob * 3 + k) * LY::kOutTile, c0 + 1, i0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (t + S - 1 < ntiles) issue_load(t + S - 1);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    ln.publish(p, lane, upd);
}

// ------------------------------------------------------------------------------------
// K0: CFL maxima over the FULL ghosted initial state (numpy.py:503-521), into slot 0.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cfl_init_kernel(const double *__restrict__ H, const double *__restrict__ U,
                                                       const double *__restrict__ V, int nrows, int jl0, int jl1,
                                                       int pitch, double g, Ctrl *slot)
{
    unsigned long long mx = 0ULL, my = 0ULL;
    const int width = jl1 - jl0;
    const long long total = (long long)nrows * width;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / width), j = jl0 + (int)(k % width);
        const size_t a = (size_t)i * pitch + j;
        const double cs = __dsqrt_rn(__dmul_rn(g, fabs(H[a])));
        // the reference's three-way maximum, evaluated literally for step 1
        const double u = U[a], v = V[a];
        const unsigned long long bx0 = nonneg_bits(__dsub_rn(u, cs)), bx1 = nonneg_bits(u), bx2 = nonneg_bits(__dadd_rn(u, cs));
        const unsigned long long by0 = nonneg_bits(__dsub_rn(v, cs)), by1 = nonneg_bits(v), by2 = nonneg_bits(__dadd_rn(v, cs));
        unsigned long long bx = bx1 > bx2 ? bx1 : bx2; bx = bx0 > bx ? bx0 : bx;
        unsigned long long by = by1 > by2 ? by1 : by2; by = by0 > by ? by0 : by;
        mx = bx > mx ? bx : mx;
        my = by > my ? by : my;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long ox = __shfl_xor_sync(0xffffffffu, mx, s);
        const unsigned long long oy = __shfl_xor_sync(0xffffffffu, my, s);
        mx = ox > mx ? ox : mx;
        my = oy > my ? oy : my;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&slot->ex_bits, mx);
        atomicMax(&slot->ey_bits, my);
    }
}

// K5: max sqrt(u*u + v*v) over the full state (numpy.py:553-554).
__global__ void __launch_bounds__(256) max_speed_kernel(const double *__restrict__ U, const double *__restrict__ V,
                                                        int nrows, int jl0, int jl1, int pitch, unsigned long long *out)
{
    unsigned long long m = 0ULL;
    const int width = jl1 - jl0;
    const long long total = (long long)nrows * width;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / width), j = jl0 + (int)(k % width);
        const size_t a = (size_t)i * pitch + j;
        const double s = __dsqrt_rn(__dadd_rn(__dmul_rn(U[a], U[a]), __dmul_rn(V[a], V[a])));
        const unsigned long long b = nonneg_bits(s);
        m = b > m ? b : m;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, s);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// Operator-level Laplacian (numpy.py:359-382 `compute_diffusion`), strict arithmetic.
__global__ void __launch_bounds__(128) laplacian_kernel(const double *__restrict__ q, double *__restrict__ out, Params p)
{
    const int jl = p.jl_first + blockIdx.x * blockDim.x + threadIdx.x;
    const int i = 1 + blockIdx.y;
    if (jl > p.jl_last || i > p.M + 1) return;
    const size_t TS = (size_t)p.ncols;
    LapCtx L;
    const int jm = max(jl - 1, p.jl_lo), jp = min(jl + 1, p.jl_hi);
    L.Aym = p.tab[T_AY * TS + jm]; L.Bym = p.tab[T_BY * TS + jm]; L.Cym = p.tab[T_CY * TS + jm];
    L.Ay0 = p.tab[T_AY * TS + jl]; L.By0 = p.tab[T_BY * TS + jl]; L.Cy0 = p.tab[T_CY * TS + jl];
    L.Ayp = p.tab[T_AY * TS + jp]; L.Byp = p.tab[T_BY * TS + jp]; L.Cyp = p.tab[T_CY * TS + jp];
    L.jm1 = jm; L.jp1 = jp; L.jm2 = max(jl - 2, p.jl_lo); L.jp2 = min(jl + 2, p.jl_hi);
    out[(size_t)i * p.pitch + jl] = laplacian_at<true>(q, p, L, p.tab[T_AC * TS + jl], i, jl);
}

}  // namespace swb
