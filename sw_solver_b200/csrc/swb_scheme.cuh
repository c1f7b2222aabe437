// swb_scheme.cuh -- the arithmetic of the scheme: rounding-controlled primitives, named load paths,
// per-cell products, half-step face values, the full-step update, diffusion weights and stars
// (numpy.py:12-133, 185-354 of the reference).  Device functions only; see swb_kernels.cuh.
#pragma once
#include "swb_types.cuh"

namespace swb {

// ------------------------------------------------------------------------------------
// arithmetic primitives
// ------------------------------------------------------------------------------------
// STRICT: IEEE round-to-nearest, never contracted -> numpy's ufunc-by-ufunc evaluation.
// FAST:   plain operators (nvcc contracts a*b+c into DFMA) and explicit fma.
// T is the arithmetic type of the state: double, or float in the optional fp32 mode (FAST only).
template <bool STRICT, typename T = double>
struct Op {
    static __device__ __forceinline__ T add(T a, T b) { return STRICT ? (T)__dadd_rn(a, b) : a + b; }
    static __device__ __forceinline__ T sub(T a, T b) { return STRICT ? (T)__dsub_rn(a, b) : a - b; }
    static __device__ __forceinline__ T mul(T a, T b) { return STRICT ? (T)__dmul_rn(a, b) : a * b; }
    static __device__ __forceinline__ T div(T a, T b) { return STRICT ? (T)__ddiv_rn(a, b) : a / b; }
};

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };
template <typename T> using vec2_t = typename Vec2<T>::type;
__device__ __forceinline__ double2 ldg2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ float2 ldg2(const float *p) { return *reinterpret_cast<const float2 *>(p); }
// How the state is read.  LD_PLAIN: the compiler's choice (read-only data path when it can prove
// the arrays constant -- march_kernel).  resident_kernel rewrites the state while it runs, so it
// names the path: LD_CG = at the L2 (single-use rows, control words), LD_CA = through the L1 (the
// diffusion star: every row is re-read by four other latitudes) -- coherent because the grid
// barrier's acquire invalidates the L1 (LDG.STRONG.GPU + CCTL.IVALL) before the next step reads.
enum LdPath { LD_PLAIN = 0, LD_CG = 1, LD_CA = 2 };
template <int LD, typename T>
__device__ __forceinline__ typename Vec2<T>::type ld2(const T *p)
{
    using V = typename Vec2<T>::type;
    if (LD == LD_CG) return __ldcg(reinterpret_cast<const V *>(p));
    if (LD == LD_CA) return __ldca(reinterpret_cast<const V *>(p));
    return *reinterpret_cast<const V *>(p);
}
template <int LD, typename T>
__device__ __forceinline__ T ld1(const T *p)
{
    if (LD == LD_CG) return __ldcg(p);
    if (LD == LD_CA) return __ldca(p);
    return *p;
}
__device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
template <typename T> struct BitsOf;
template <> struct BitsOf<double> { using type = unsigned long long; };
template <> struct BitsOf<float> { using type = unsigned int; };
template <typename T> using bits_t = typename BitsOf<T>::type;

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

// fp32: MUFU.RCP (~1 ulp, biased) + one Newton step -- the residual is then far below half an
// ulp, so r is the correctly rounded reciprocal almost always and errors do not pile up
// coherently over thousands of steps; MUFU.SQRT alone serves the CFL estimate.
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned long long nonneg_bits(double v)
{
    // |v| as an unsigned integer: monotone for finite values and +inf, and every NaN maps
    // above +inf, so an integer max reproduces numpy's NaN-propagating max.  (Integer
    // pipe only: the FP64 pipe is the scarce resource of this kernel.)
    return ((unsigned long long)(unsigned)(__double2hiint(v) & 0x7fffffff) << 32) | (unsigned)__double2loint(v);
}

__device__ __forceinline__ unsigned int nonneg_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
// the same ordering key widened to the fp64 pattern the control slots hold
__device__ __forceinline__ unsigned long long widen_bits(unsigned long long b) { return b; }
__device__ __forceinline__ unsigned long long widen_bits(unsigned int b) { return nonneg_bits((double)__uint_as_float(b)); }

__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned int umax64(unsigned int a, unsigned int b) { return a > b ? a : b; }
__device__ __forceinline__ int hi_word(double x) { return __double2hiint(x); }
__device__ __forceinline__ int hi_word(float x) { return __float_as_int(x); }

template <typename T> __device__ __forceinline__ T shfl_down1(T x) { return __shfl_down_sync(0xffffffffu, x, 1); }
template <typename T> __device__ __forceinline__ T shfl_up1(T x) { return __shfl_up_sync(0xffffffffu, x, 1); }

// ------------------------------------------------------------------------------------
// per-cell products (numpy.py:185-188, 203, 220, 237, 254-255)
// ------------------------------------------------------------------------------------
template <typename T>
struct Cell {
    T h, u, v;
    T hu, hv;   // h*u, h*v
    T Ux, Vx;   // hu*u + (hg*h)*h ; hu*v
    T hw;       // h*(v*c)          ("hv1")
    T Uy, Vy1;  // hu*(v*c) ; hv*(v*c)
    T pr;       // (hg*h)*h         ("Vy2", also the pressure part of Ux)
    T f;        // Coriolis at the cell (only when it is a 2-D field)
};

template <bool STRICT, typename T>
__device__ __forceinline__ Cell<T> make_cell(T h, T u, T v, T c, T hg, T f)
{
    using O = Op<STRICT, T>;
    Cell<T> q;
    q.h = h; q.u = u; q.v = v; q.f = f;
    q.hu = O::mul(h, u);
    q.hv = O::mul(h, v);
    q.pr = O::mul(O::mul(hg, h), h);
    q.Ux = STRICT ? O::add(O::mul(q.hu, u), q.pr) : fma(q.hu, u, q.pr);
    q.Vx = O::mul(q.hu, v);
    const T w = O::mul(v, c);
    q.hw = O::mul(h, w);
    q.Uy = O::mul(q.hu, w);
    q.Vy1 = O::mul(q.hv, w);
    return q;
}

// Face record.  STRICT stores the reference's quantities; FAST stores hM, huM, hvM and the
// fluxes DOUBLED (a factor 2 is exact), with the 0.5 folded into the update coefficients.
template <typename T>
struct Face {
    T hM, huM, hvM;
    T q;           // huM / hM, unguarded (numpy.py:302-305)
    T F1, F2, F3;  // x: UxMid, VxMid, - ; y: UyMid, Vy1Mid, Vy2Mid
    T hvc;         // y only: hvMidy * cMidy
};

// hM > 0 on the integer pipe (false for -0, +0 and negative values; NaN is not an issue:
// a NaN state stops the loop through the CFL maxima)
__device__ __forceinline__ bool positive(double x) { return __double_as_longlong(x) > 0; }
__device__ __forceinline__ bool positive(float x) { return __float_as_int(x) > 0; }

// x-face between (i, j) = `a` and (i+1, j) = `b`.  numpy.py:193-195, 203-217, 237-251,
// 287-291, 321.
//   STRICT: cdx = hdt/dx[i,j];  kt = tan(theta_j) (tgMidx);  fK = 0.5*(f_b+f_a)
//   FAST:   cdx = dt/dx_j;      kt = hdt*0.5*tan/a;          fK = hdt*f
// GUARD evaluates the reference's `hMid > 0` selects; without it they are assumed true
// (positive layer depth) and the caller watches the sign of hM instead.
template <bool STRICT, bool GUARD, typename T>
__device__ __forceinline__ Face<T> x_face(const Cell<T> &a, const Cell<T> &b, T cdx, T kt, T fK, T hdt, T hg, T ra)
{
    using O = Op<STRICT, T>;
    Face<T> F;
    if (STRICT) {
        const T P = O::add(fK, O::div(O::mul(O::mul(T(0.5), O::add(b.u, a.u)), kt), ra));  // ra = a
        const T K = O::mul(hdt, P);
        const T shu = O::mul(T(0.5), O::add(b.hu, a.hu));
        const T shv = O::mul(T(0.5), O::add(b.hv, a.hv));
        F.hM = O::sub(O::mul(T(0.5), O::add(b.h, a.h)), O::mul(cdx, O::sub(b.hu, a.hu)));
        F.huM = O::add(O::sub(shu, O::mul(cdx, O::sub(b.Ux, a.Ux))), O::mul(K, shv));
        F.hvM = O::sub(O::sub(shv, O::mul(cdx, O::sub(b.Vx, a.Vx))), O::mul(K, shu));
        F.q = O::div(F.huM, F.hM);
        const T pres = O::mul(O::mul(hg, F.hM), F.hM);
        const bool pos = F.hM > 0.0;
        F.F1 = pos ? O::add(O::div(O::mul(F.huM, F.huM), F.hM), pres) : pres;
        F.F2 = pos ? O::div(O::mul(F.hvM, F.huM), F.hM) : T(0);
    } else {
        const T K = fma(b.u + a.u, kt, fK);
        const T shu = b.hu + a.hu, shv = b.hv + a.hv;
        F.hM = fma(-cdx, b.hu - a.hu, b.h + a.h);             // 2 hMidx
        F.huM = fma(K, shv, fma(-cdx, b.Ux - a.Ux, shu));     // 2 huMidx
        F.hvM = fma(-K, shu, fma(-cdx, b.Vx - a.Vx, shv));    // 2 hvMidx
        const T r = fast_rcp(F.hM);
        F.q = F.huM * r;
        const T pres = (hg * F.hM) * F.hM;               // hg here is g/4: 2*(g/2)*hM^2
        const T qg = (!GUARD || positive(F.hM)) ? F.q : T(0);
        F.F1 = fma(F.huM, qg, pres);                          // 2 UxMid
        F.F2 = F.hvM * qg;                                    // 2 VxMid
    }
    F.F3 = T(0);
    F.hvc = T(0);
    return F;
}

// y-face between (i, j) = `a` and (i, j+1) = `b`.  numpy.py:198-200, 220-234, 254-270,
// 292, 322-323.
//   STRICT: c1 = hdt/dy1_j, c2 = hdt/dy_j, kt = tgMidy_j, fK = 0.5*(f_b+f_a)
//   FAST:   c1 = dt/dy1_j,  c2 = dt/dy_j,  kt = hdt*0.5*tgMidy/a, fK = hdt*0.5*(f_b+f_a)
template <bool STRICT, bool GUARD, typename T>
__device__ __forceinline__ Face<T> y_face(const Cell<T> &a, const Cell<T> &b, T c1, T c2, T kt, T fK, T cm, T hdt, T hg, T ra)
{
    using O = Op<STRICT, T>;
    Face<T> F;
    if (STRICT) {
        const T P = O::add(fK, O::div(O::mul(O::mul(T(0.5), O::add(b.u, a.u)), kt), ra));
        const T K = O::mul(hdt, P);
        const T shu = O::mul(T(0.5), O::add(b.hu, a.hu));
        const T shv = O::mul(T(0.5), O::add(b.hv, a.hv));
        F.hM = O::sub(O::mul(T(0.5), O::add(b.h, a.h)), O::mul(c1, O::sub(b.hw, a.hw)));
        F.huM = O::add(O::sub(shu, O::mul(c1, O::sub(b.Uy, a.Uy))), O::mul(K, shv));
        F.hvM = O::sub(O::sub(O::sub(shv, O::mul(c1, O::sub(b.Vy1, a.Vy1))), O::mul(c2, O::sub(b.pr, a.pr))),
                       O::mul(K, shu));
        F.q = O::div(F.huM, F.hM);
        F.hvc = O::mul(F.hvM, cm);
        const bool pos = F.hM > 0.0;
        F.F1 = pos ? O::div(O::mul(F.hvc, F.huM), F.hM) : T(0);
        F.F2 = pos ? O::mul(O::div(O::mul(F.hvM, F.hvM), F.hM), cm) : T(0);
        F.F3 = O::mul(O::mul(hg, F.hM), F.hM);
    } else {
        const T K = fma(b.u + a.u, kt, fK);
        const T shu = b.hu + a.hu, shv = b.hv + a.hv;
        F.hM = fma(-c1, b.hw - a.hw, b.h + a.h);
        F.huM = fma(K, shv, fma(-c1, b.Uy - a.Uy, shu));
        F.hvM = fma(-K, shu, fma(-c2, b.pr - a.pr, fma(-c1, b.Vy1 - a.Vy1, shv)));
        const T r = fast_rcp(F.hM);
        F.q = F.huM * r;
        F.hvc = F.hvM * cm;
        const T t = F.hvM * r;
        const bool pos = !GUARD || positive(F.hM);
        F.F1 = F.hvc * (pos ? F.q : T(0));          // 2 UyMid
        F.F2 = F.hvc * (pos ? t : T(0));            // 2 Vy1Mid
        F.F3 = (hg * F.hM) * F.hM;                 // 2 Vy2Mid (hg = g/4)
    }
    return F;
}

// Full-step update of one cell from its four faces (numpy.py:275-354).
//   STRICT: cx = dt/dxc[i,j], cy1u = dt/dy1c_j, cyu = dt/dyc_j, kc = tan(theta_j), fc = f
//   FAST:   cx = 0.5*dt/dx_j, cy1u, cyu halved, kc = dt/64*tan/a, fc = dt/8*f
// hsx = hs[i+1,j]-hs[i-1,j], hsy = hs[i,j+1]-hs[i,j-1], dxs = dx[i-1,j]+dx[i,j], dy1s = dy1_{j-1}+dy1_j.
template <bool STRICT, bool HS, typename T>
__device__ __forceinline__ void update_cell(const Cell<T> &cur, const Face<T> &x0, const Face<T> &x1, const Face<T> &y0,
                                            const Face<T> &y1, T cx, T cy1u, T cyu, T kc, T fc, T dt, T g, T ra, T hsx, T hsy,
                                            T dxs, T dy1s, T &hn, T &un, T &vn)
{
    using O = Op<STRICT, T>;
    if (STRICT) {
        hn = O::sub(O::sub(cur.h, O::mul(cx, O::sub(x1.huM, x0.huM))), O::mul(cy1u, O::sub(y1.hvc, y0.hvc)));
        const T qs = O::add(O::add(O::add(x0.q, x1.q), y0.q), y1.q);
        const T Fc = O::add(fc, O::div(O::mul(O::mul(T(0.25), qs), kc), ra));
        const T D = O::mul(O::mul(dt, Fc), T(0.25));
        const T shv = O::add(O::add(O::add(x0.hvM, x1.hvM), y0.hvM), y1.hvM);
        const T shu = O::add(O::add(O::add(x0.huM, x1.huM), y0.huM), y1.huM);
        T hun = O::add(O::sub(O::sub(cur.hu, O::mul(cx, O::sub(x1.F1, x0.F1))), O::mul(cy1u, O::sub(y1.F1, y0.F1))),
                            O::mul(D, shv));
        T hvn = O::sub(O::sub(O::sub(O::sub(cur.hv, O::mul(cx, O::sub(x1.F2, x0.F2))), O::mul(cy1u, O::sub(y1.F2, y0.F2))),
                                   O::mul(cyu, O::sub(y1.F3, y0.F3))),
                            O::mul(D, shu));
        if (HS) {  // numpy.py:310-317, 342-349
            const T hsum = O::add(O::add(O::add(x0.hM, x1.hM), y0.hM), y1.hM);
            const T gq = O::mul(O::mul(O::mul(dt, g), T(0.25)), hsum);
            hun = O::sub(hun, O::div(O::mul(gq, hsx), dxs));
            hvn = O::sub(hvn, O::div(O::mul(gq, hsy), dy1s));
        }
        un = O::div(hun, hn);
        vn = O::div(hvn, hn);
    } else if (sizeof(T) == 4) {
        // fp32: increment form.  h' = h + dh, (hu)' = hu + dhu  =>  u' = u + (dhu - u dh) / h':
        // the increments are ~1e-3 of the fields, so their rounding errors stay far below an
        // ulp of u, v instead of the ~2 ulp per step the quotient (hu)'/h' costs in fp32.
        const T dh = fma(-cy1u, y1.hvc - y0.hvc, -cx * (x1.huM - x0.huM));
        hn = cur.h + dh;
        const T qs = (x0.q + x1.q) + (y0.q + y1.q);
        const T D = fma(qs, kc, fc);
        const T shv = (x0.hvM + x1.hvM) + (y0.hvM + y1.hvM);
        const T shu = (x0.huM + x1.huM) + (y0.huM + y1.huM);
        const T dhu = fma(D, shv, fma(-cy1u, y1.F1 - y0.F1, -cx * (x1.F1 - x0.F1)));
        const T dhv = fma(-D, shu, fma(-cyu, y1.F3 - y0.F3, fma(-cy1u, y1.F2 - y0.F2, -cx * (x1.F2 - x0.F2))));
        const T r = fast_rcp(hn);
        un = fma(fma(-cur.u, dh, dhu), r, cur.u);
        vn = fma(fma(-cur.v, dh, dhv), r, cur.v);
    } else {
        hn = fma(-cy1u, y1.hvc - y0.hvc, fma(-cx, x1.huM - x0.huM, cur.h));
        const T qs = (x0.q + x1.q) + (y0.q + y1.q);
        const T D = fma(qs, kc, fc);  // 0.5 * dt * Fc * 0.25
        const T shv = (x0.hvM + x1.hvM) + (y0.hvM + y1.hvM);
        const T shu = (x0.huM + x1.huM) + (y0.huM + y1.huM);
        T hun = fma(D, shv, fma(-cy1u, y1.F1 - y0.F1, fma(-cx, x1.F1 - x0.F1, cur.hu)));
        T hvn = fma(-D, shu, fma(-cyu, y1.F3 - y0.F3, fma(-cy1u, y1.F2 - y0.F2, fma(-cx, x1.F2 - x0.F2, cur.hv))));
        if (HS) {
            const T hsum = (x0.hM + x1.hM) + (y0.hM + y1.hM);  // doubled
            const T gq = dt * g * T(0.125) * hsum;
            hun -= gq * hsx / dxs;
            hvn -= gq * hsy / dy1s;
        }
        const T r = fast_rcp(hn);
        un = hun * r;
        vn = hvn * r;
    }
}

// ------------------------------------------------------------------------------------
// fp32, packed: a lane's two cells (A in the low half, B in the high half) through Blackwell's two-wide
// FP32 instructions (FFMA2 / FADD2 / FMUL2: fma / add / sub / mul .rn.f32x2).  They round element for
// element like the scalar instructions (scripts/probes/f32x2_probe.cu: 0 mismatches in 2 x 12 G results),
// and the functions below perform EXACTLY the operations of make_cell / x_face / y_face / update_cell with
// T = float, in the same order -- so a row computed here and a row computed by the scalar boundary body
// are the same bits (test: SWB_FORCE_RARE=1 against the default).  The fp32 kernel is bound by instruction
// issue (profiles/r2_fp32_ncu.md): halving its arithmetic instructions is the way up.
// PTX has no way to negate an operand of fma.rn.f32x2, although the instruction (FFMA2) can: the one route to it is ptxas's
// contraction of mul.rn.f32x2 + sub.rn.f32x2 into `FFMA2 d, -a, b, c` (CUDA 12.9; micro-test in scripts/probes/f32x2_probe.cu).
// `pnfma` below spells -a * b + c that way; the row coefficients that enter the scheme negated are STORED negated instead
// (rec_slot, swb_step.cuh).  The bit-identity test named above pins the contraction: were it ever not applied, the hot rows
// would differ from the boundary rows by a rounding and the test would fail.
// ------------------------------------------------------------------------------------
struct P2 {
    unsigned long long v;
};
__device__ __forceinline__ P2 pk(float lo, float hi)
{
    P2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ P2 pk(float2 q) { return pk(q.x, q.y); }
__device__ __forceinline__ P2 pk(double2) { return P2{0ULL}; }  // (never called: lets the discarded packed branch of the fp64 kernels parse)
__device__ __forceinline__ P2 pdup(float c) { return pk(c, c); }
__device__ __forceinline__ float plo(P2 a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    (void)hi;
    return lo;
}
__device__ __forceinline__ float phi(P2 a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    (void)lo;
    return hi;
}
__device__ __forceinline__ P2 padd(P2 a, P2 b) { P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 psub(P2 a, P2 b) { P2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 pmul(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2 pfma(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ P2 pnfma(P2 a, P2 b, P2 c) { return psub(c, pmul(a, b)); }  // -a * b + c, ONE rounding (contracted by ptxas, see above)
__device__ __forceinline__ P2 prcp(P2 x)  // fast_rcp(float) on both halves
{
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(plo(x)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(phi(x)));
    const P2 rr = pk(r0, r1);
    return pfma(rr, pnfma(x, rr, pdup(1.0f)), rr);
}
// both halves: positive(x) ? v : 0
__device__ __forceinline__ P2 psel_pos(P2 x, P2 v)
{
    return pk(positive(plo(x)) ? plo(v) : 0.0f, positive(phi(x)) ? phi(v) : 0.0f);
}

struct Cell2 {
    P2 h, u, v, hu, hv, Ux, Vx, hw, Uy, Vy1, pr;
};
struct Face2 {
    P2 hM, huM, hvM, q, F1, F2, F3, hvc;
};
__device__ __forceinline__ Cell2 pack_cells(const Cell<float> &a, const Cell<float> &b)
{
    Cell2 q;
    q.h = pk(a.h, b.h); q.u = pk(a.u, b.u); q.v = pk(a.v, b.v); q.hu = pk(a.hu, b.hu); q.hv = pk(a.hv, b.hv);
    q.Ux = pk(a.Ux, b.Ux); q.Vx = pk(a.Vx, b.Vx); q.hw = pk(a.hw, b.hw); q.Uy = pk(a.Uy, b.Uy); q.Vy1 = pk(a.Vy1, b.Vy1); q.pr = pk(a.pr, b.pr);
    return q;
}
__device__ __forceinline__ void unpack_cells(const Cell2 &q, Cell<float> &a, Cell<float> &b)
{
    a.h = plo(q.h); b.h = phi(q.h); a.u = plo(q.u); b.u = phi(q.u); a.v = plo(q.v); b.v = phi(q.v);
    a.hu = plo(q.hu); b.hu = phi(q.hu); a.hv = plo(q.hv); b.hv = phi(q.hv); a.Ux = plo(q.Ux); b.Ux = phi(q.Ux);
    a.Vx = plo(q.Vx); b.Vx = phi(q.Vx); a.hw = plo(q.hw); b.hw = phi(q.hw); a.Uy = plo(q.Uy); b.Uy = phi(q.Uy);
    a.Vy1 = plo(q.Vy1); b.Vy1 = phi(q.Vy1); a.pr = plo(q.pr); b.pr = phi(q.pr);
    a.f = 0.0f; b.f = 0.0f;
}
__device__ __forceinline__ Face2 pack_faces(const Face<float> &a, const Face<float> &b)
{
    Face2 F;
    F.hM = pk(a.hM, b.hM); F.huM = pk(a.huM, b.huM); F.hvM = pk(a.hvM, b.hvM); F.q = pk(a.q, b.q);
    F.F1 = pk(a.F1, b.F1); F.F2 = pk(a.F2, b.F2); F.F3 = pk(a.F3, b.F3); F.hvc = pk(a.hvc, b.hvc);
    return F;
}
__device__ __forceinline__ void unpack_faces(const Face2 &F, Face<float> &a, Face<float> &b)
{
    a.hM = plo(F.hM); b.hM = phi(F.hM); a.huM = plo(F.huM); b.huM = phi(F.huM); a.hvM = plo(F.hvM); b.hvM = phi(F.hvM);
    a.q = plo(F.q); b.q = phi(F.q); a.F1 = plo(F.F1); b.F1 = phi(F.F1); a.F2 = plo(F.F2); b.F2 = phi(F.F2);
    a.F3 = plo(F.F3); b.F3 = phi(F.F3); a.hvc = plo(F.hvc); b.hvc = phi(F.hvc);
}

// make_cell<false, float> for both cells of the lane (c, hg: the row's cos(theta) and g/2, in both halves)
__device__ __forceinline__ Cell2 make_cell2(P2 h, P2 u, P2 v, P2 c, P2 hg)
{
    Cell2 q;
    q.h = h; q.u = u; q.v = v;
    q.hu = pmul(h, u);
    q.hv = pmul(h, v);
    q.pr = pmul(pmul(hg, h), h);
    q.Ux = pfma(q.hu, u, q.pr);
    q.Vx = pmul(q.hu, v);
    const P2 w = pmul(v, c);
    q.hw = pmul(h, w);
    q.Uy = pmul(q.hu, w);
    q.Vy1 = pmul(q.hv, w);
    return q;
}
// x_face<false, GUARD, float>: the faces right of `a`'s cells (ncdx = -dt/dx_j; hg = g/4)
template <bool GUARD>
__device__ __forceinline__ Face2 x_face2(const Cell2 &a, const Cell2 &b, P2 ncdx, P2 kt, P2 fK, P2 hg)
{
    Face2 F;
    const P2 K = pfma(padd(b.u, a.u), kt, fK);
    const P2 shu = padd(b.hu, a.hu), shv = padd(b.hv, a.hv);
    F.hM = pfma(ncdx, psub(b.hu, a.hu), padd(b.h, a.h));
    F.huM = pfma(K, shv, pfma(ncdx, psub(b.Ux, a.Ux), shu));
    F.hvM = pnfma(K, shu, pfma(ncdx, psub(b.Vx, a.Vx), shv));
    const P2 r = prcp(F.hM);
    F.q = pmul(F.huM, r);
    const P2 pres = pmul(pmul(hg, F.hM), F.hM);
    const P2 qg = GUARD ? psel_pos(F.hM, F.q) : F.q;
    F.F1 = pfma(F.huM, qg, pres);
    F.F2 = pmul(F.hvM, qg);
    F.F3 = pdup(0.0f);
    F.hvc = pdup(0.0f);
    return F;
}
// y_face<false, GUARD, float>: the faces above `a`'s cells (nc1 = -dt/dy1_j, nc2 = -dt/dy_j)
template <bool GUARD>
__device__ __forceinline__ Face2 y_face2(const Cell2 &a, const Cell2 &b, P2 nc1, P2 nc2, P2 kt, P2 fK, P2 cm, P2 hg)
{
    Face2 F;
    const P2 K = pfma(padd(b.u, a.u), kt, fK);
    const P2 shu = padd(b.hu, a.hu), shv = padd(b.hv, a.hv);
    F.hM = pfma(nc1, psub(b.hw, a.hw), padd(b.h, a.h));
    F.huM = pfma(K, shv, pfma(nc1, psub(b.Uy, a.Uy), shu));
    F.hvM = pnfma(K, shu, pfma(nc2, psub(b.pr, a.pr), pfma(nc1, psub(b.Vy1, a.Vy1), shv)));
    const P2 r = prcp(F.hM);
    F.q = pmul(F.huM, r);
    F.hvc = pmul(F.hvM, cm);
    const P2 t = pmul(F.hvM, r);
    F.F1 = pmul(F.hvc, GUARD ? psel_pos(F.hM, F.q) : F.q);
    F.F2 = pmul(F.hvc, GUARD ? psel_pos(F.hM, t) : t);
    F.F3 = pmul(pmul(hg, F.hM), F.hM);
    return F;
}
// update_cell<false, false, float> (the increment form): ncx = -0.5 dt/dx_j, ncy1u, ncyu likewise negated
__device__ __forceinline__ void update_cell2(const Cell2 &cur, const Face2 &x0, const Face2 &x1, const Face2 &y0, const Face2 &y1, P2 ncx, P2 ncy1u,
                                             P2 ncyu, P2 kc, P2 fc, P2 &hn, P2 &un, P2 &vn)
{
    const P2 dh = pfma(ncy1u, psub(y1.hvc, y0.hvc), pmul(ncx, psub(x1.huM, x0.huM)));
    hn = padd(cur.h, dh);
    const P2 qs = padd(padd(x0.q, x1.q), padd(y0.q, y1.q));
    const P2 D = pfma(qs, kc, fc);
    const P2 shv = padd(padd(x0.hvM, x1.hvM), padd(y0.hvM, y1.hvM));
    const P2 shu = padd(padd(x0.huM, x1.huM), padd(y0.huM, y1.huM));
    const P2 dhu = pfma(D, shv, pfma(ncy1u, psub(y1.F1, y0.F1), pmul(ncx, psub(x1.F1, x0.F1))));
    const P2 dhv = pnfma(D, shu, pfma(ncyu, psub(y1.F3, y0.F3), pfma(ncy1u, psub(y1.F2, y0.F2), pmul(ncx, psub(x1.F2, x0.F2)))));
    const P2 r = prcp(hn);
    un = pfma(pnfma(cur.u, dh, dhu), r, cur.u);
    vn = pfma(pnfma(cur.v, dh, dhv), r, cur.v);
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

struct W3 {  // weights of three consecutive points (minus, centre, plus)
    double A[3], B[3], C[3];
};

__device__ __forceinline__ double second_diff_strict(const W3 &w, double qmm, double qm, double q0, double qp, double qpp)
{
    using O = Op<true>;
    // numpy.py:87-105: A*(A q + B q+ + C q-) + B*(A+ q+ + B+ q++ + C+ q) + C*(A- q- + B- q + C- q--)
    const double t0 = O::mul(w.A[1], O::add(O::add(O::mul(w.A[1], q0), O::mul(w.B[1], qp)), O::mul(w.C[1], qm)));
    const double t1 = O::mul(w.B[1], O::add(O::add(O::mul(w.A[2], qp), O::mul(w.B[2], qpp)), O::mul(w.C[2], q0)));
    const double t2 = O::mul(w.C[1], O::add(O::add(O::mul(w.A[0], qm), O::mul(w.B[0], q0)), O::mul(w.C[0], qmm)));
    return O::add(O::add(t0, t1), t2);
}

// longitude weights of columns i-1, i, i+1 (numpy.py:32-45; column 0 uses column M's,
// column M+2 uses column 2's), rebuilt from x = ac_j * phi_i
__device__ __forceinline__ W3 x_weights_strict(const double *__restrict__ phi, double ac, int i, int M)
{
    using O = Op<true>;
    W3 w;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        int k = i - 1 + s;
        k = (k == 0) ? M : ((k == M + 2) ? 2 : k);
        const double xm = O::mul(ac, phi[k - 1]), x0 = O::mul(ac, phi[k]), xp = O::mul(ac, phi[k + 1]);
        weights_strict(O::sub(x0, xm), O::sub(xp, x0), w.A[s], w.B[s], w.C[s]);
    }
    return w;
}

// latitude weights of rows j-1, j, j+1 (numpy.py:50-63), clamped to the local array
__device__ __forceinline__ W3 y_weights(const double *__restrict__ tab, int nrows, int jl)
{
    W3 w;
    const int jm = max(jl - 1, 0), jp = min(jl + 1, nrows - 1);
    const int idx[3] = {jm, jl, jp};
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        w.A[s] = tab[(size_t)T_AY * nrows + idx[s]];
        w.B[s] = tab[(size_t)T_BY * nrows + idx[s]];
        w.C[s] = tab[(size_t)T_CY * nrows + idx[s]];
    }
    return w;
}

// the five stencil values of q along each axis around (jl, i), with the reference's
// extension (numpy.py:376-381): periodic in longitude, duplicated rows in latitude
template <typename T>
struct Star {
    T q0, xm, xp, xmm, xpp, ym, yp, ymm, ypp;
};
template <typename T, int LD = LD_PLAIN>
__device__ __forceinline__ Star<T> load_star(const T *__restrict__ q, int pitch, int nrows, int M, int jl, int i)
{
    const size_t P = (size_t)pitch;
    const int im2 = (i - 2 < 0) ? M - 1 : i - 2;  // numpy.py:376-378
    const int ip2 = (i + 2 > M + 2) ? 3 : i + 2;
    const int jm1 = max(jl - 1, 0), jp1 = min(jl + 1, nrows - 1);
    const int jm2 = max(jl - 2, 0), jp2 = min(jl + 2, nrows - 1);
    Star<T> s;
    const T *row = q + (size_t)jl * P;
    s.q0 = ld1<LD>(row + i);
    s.xm = ld1<LD>(row + i - 1); s.xp = ld1<LD>(row + i + 1);
    s.xmm = ld1<LD>(row + im2);  s.xpp = ld1<LD>(row + ip2);
    s.ym = ld1<LD>(q + (size_t)jm1 * P + i); s.yp = ld1<LD>(q + (size_t)jp1 * P + i);
    s.ymm = ld1<LD>(q + (size_t)jm2 * P + i); s.ypp = ld1<LD>(q + (size_t)jp2 * P + i);
    return s;
}

// The same for a lane's pair of cells (jl, colA), (jl, colA+1) away from every boundary: six
// 16-byte loads per field instead of eighteen scalar ones; the centre values come from
// registers.
template <typename T, int LD = LD_PLAIN>
__device__ __forceinline__ void load_star_pair(const T *__restrict__ q, int pitch, int jl, int colA, T q0A, T q0B,
                                               Star<T> &sA, Star<T> &sB)
{
    const size_t P = (size_t)pitch;
    const T *row = q + (size_t)jl * P + colA;
    const vec2_t<T> L = ld2<LD>(row - 2), Rr = ld2<LD>(row + 2);
    const vec2_t<T> m2 = ld2<LD>(row - 2 * P), m1 = ld2<LD>(row - P), p1 = ld2<LD>(row + P), p2 = ld2<LD>(row + 2 * P);
    sA.q0 = q0A; sA.xmm = L.x; sA.xm = L.y; sA.xp = q0B;  sA.xpp = Rr.x;
    sB.q0 = q0B; sB.xmm = L.y; sB.xm = q0A; sB.xp = Rr.x; sB.xpp = Rr.y;
    sA.ymm = m2.x; sA.ym = m1.x; sA.yp = p1.x; sA.ypp = p2.x;
    sB.ymm = m2.y; sB.ym = m1.y; sB.yp = p1.y; sB.ypp = p2.y;
}

// FAST diffusion row record: w[0..4] combined latitude weights, [5] = 1/(2 dx_j)^2, [6] = dt*nu
// (Tried in round 2 and reverted: dt nu folded into the six weights and the update written as FMA chains --
// 8 FP64 instructions per cell and field when the chains start from the Lax-Wendroff result, 9 when only the
// last addition waits for it, against 10 here.  Neither was faster: 7200 x 3600 with diffusion 456 -> 449 / 456 us
// per step, and the resident 1440 x 720 step went from 25.9 to 27.6 / 27.9 us -- the kernel is bound by latency at
// 255 registers, not by its instruction count.)
template <typename T>
__device__ __forceinline__ T laplacian_fast(const Star<T> &s, const T *__restrict__ rd)
{
    const T qxx = rd[5] * fma(T(-2.0), s.q0, s.xpp + s.xmm);
    const T qyy = fma(rd[0], s.ymm, fma(rd[1], s.ym, fma(rd[2], s.q0, fma(rd[3], s.yp, rd[4] * s.ypp))));
    return qxx + qyy;
}

}  // namespace swb
