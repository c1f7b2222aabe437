/*
 * sw_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Scalar, per-grid-point restatement of the hot path of ai2cm/sw_solver's numpy solver:
 * the body of the `while time < final_time` loop at src/sw_solver/numpy.py:496-549
 * (CFL time step, two-step Lax-Wendroff update, optional Laplacian diffusion, boundary
 * conditions).  Every function cites the reference lines it follows.  All arrays are
 * C-contiguous float64 in the reference's own layout: state (M+3, N) indexed [i][j] with
 * latitude j contiguous, metric arrays with the shapes grid.py gives them.
 *
 * Arithmetic contract: numpy evaluates each ufunc separately in IEEE double, left to
 * right, with no fused multiply-add.  This file must therefore be compiled with
 * `-ffp-contract=off` and without -ffast-math (see oracle/Makefile); with that it is
 * bit-identical to the reference (pinned by tests/test_oracle.py against
 * the tests/golden fixtures, which were produced by importing the reference itself).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call into this file.  The product (sw_solver_b200/) never does.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int M;              /* num_lon_pts; state has M+3 longitude columns              */
    int N;              /* num_lat_pts (already made even, grid.py:83)                */
    double g, a, nu;    /* utils.py:28-33                                             */
    /* LatLonGrid (grid.py:99-108) */
    const double *c;      /* (M+3, N)   cos(theta)                                    */
    const double *cMidy;  /* (M+1, N-1) cos at y-faces                                */
    const double *tg;     /* (M+1, N-2) tan(theta) at interior cells                  */
    const double *tgMidx; /* (M+2, N-2) tan at x-faces                                */
    const double *tgMidy; /* (M+1, N-1) tan at y-faces                                */
    /* CartesianGrid (grid.py:35-53) */
    const double *dx;     /* (M+2, N)                                                 */
    const double *dy;     /* (M+3, N-1)                                               */
    const double *dy1;    /* (M+3, N-1)                                               */
    const double *dxc;    /* (M+1, N-2)                                               */
    const double *dyc;    /* (M+1, N-2)                                               */
    const double *dy1c;   /* (M+1, N-2)                                               */
    /* DiffusionCoefficients (numpy.py:27-63); may be NULL when diffusion is off */
    const double *Ax, *Bx, *Cx; /* (M+3, N-2) */
    const double *Ay, *By, *Cy; /* (M+1, N)   */
} swo_grid;

/* half-step values and fluxes living on one cell face */
typedef struct {
    double hM, huM, hvM; /* mid-point h, hu, hv                                       */
    double q;            /* huM / hM (unguarded, numpy.py:302-305)                    */
    double F1, F2, F3;   /* x-face: UxMid, VxMid, -  ;  y-face: UyMid, Vy1Mid, Vy2Mid */
    double hvc;          /* y-face only: hvMidy * cMidy (numpy.py:281-282)            */
} swo_face;

#define S(arr, i, j) ((arr)[(size_t)(i) * (size_t)N + (size_t)(j)])

/* x-face between columns i and i+1 at interior row j (1 <= j <= N-2), i in [0, M+1].
 * numpy.py:193-195 (hMidx), 203-217 (huMidx), 237-251 (hvMidx), 287-291 (UxMid),
 * 321 (VxMid). */
static swo_face x_face(const swo_grid *G, double dt, const double *f, const double *h,
                       const double *u, const double *v, int i, int j)
{
    const int N = G->N;
    const double hg = 0.5 * G->g;
    const double hdt = 0.5 * dt;
    const double h0 = S(h, i, j), h1 = S(h, i + 1, j);
    const double u0 = S(u, i, j), u1 = S(u, i + 1, j);
    const double v0 = S(v, i, j), v1 = S(v, i + 1, j);
    const double hu0 = h0 * u0, hu1 = h1 * u1;   /* numpy.py:186 */
    const double hv0 = h0 * v0, hv1 = h1 * v1;   /* numpy.py:187 */
    const double Ux0 = hu0 * u0 + hg * h0 * h0;  /* numpy.py:203 */
    const double Ux1 = hu1 * u1 + hg * h1 * h1;
    const double Vx0 = hu0 * v0, Vx1 = hu1 * v1; /* numpy.py:237 */
    const double cdx = hdt / S(G->dx, i, j);
    const double tgm = G->tgMidx[(size_t)i * (size_t)(N - 2) + (size_t)(j - 1)];
    const double P = 0.5 * (S(f, i + 1, j) + S(f, i, j)) + 0.5 * (u1 + u0) * tgm / G->a;
    swo_face F;
    F.hM = 0.5 * (h1 + h0) - cdx * (hu1 - hu0);
    F.huM = 0.5 * (hu1 + hu0) - cdx * (Ux1 - Ux0) + hdt * P * (0.5 * (hv1 + hv0));
    F.hvM = 0.5 * (hv1 + hv0) - cdx * (Vx1 - Vx0) - hdt * P * (0.5 * (hu1 + hu0));
    F.q = F.huM / F.hM;
    const double pres = hg * F.hM * F.hM;
    F.F1 = (F.hM > 0.0) ? (F.huM * F.huM / F.hM + pres) : pres;
    F.F2 = (F.hM > 0.0) ? (F.hvM * F.huM / F.hM) : 0.0;
    F.F3 = 0.0;
    F.hvc = 0.0;
    return F;
}

/* y-face between rows j and j+1 at interior column i (1 <= i <= M+1), j in [0, N-2].
 * numpy.py:198-200 (hMidy), 220-234 (huMidy), 254-270 (hvMidy), 292 (UyMid),
 * 322-323 (Vy1Mid, Vy2Mid). */
static swo_face y_face(const swo_grid *G, double dt, const double *f, const double *h,
                       const double *u, const double *v, int i, int j)
{
    const int N = G->N;
    const double hg = 0.5 * G->g;
    const double hdt = 0.5 * dt;
    const double h0 = S(h, i, j), h1 = S(h, i, j + 1);
    const double u0 = S(u, i, j), u1 = S(u, i, j + 1);
    const double v0 = S(v, i, j), v1 = S(v, i, j + 1);
    const double w0 = v0 * S(G->c, i, j), w1 = v1 * S(G->c, i, j + 1); /* numpy.py:185 */
    const double hu0 = h0 * u0, hu1 = h1 * u1;
    const double hv0 = h0 * v0, hv1 = h1 * v1;
    const double hw0 = h0 * w0, hw1 = h1 * w1;       /* hv1, numpy.py:188 */
    const double Uy0 = hu0 * w0, Uy1 = hu1 * w1;     /* numpy.py:220      */
    const double Vy10 = hv0 * w0, Vy11 = hv1 * w1;   /* numpy.py:254      */
    const double Vy20 = hg * h0 * h0, Vy21 = hg * h1 * h1; /* numpy.py:255 */
    const size_t kf = (size_t)i * (size_t)(N - 1) + (size_t)j;
    const double cdy1 = hdt / G->dy1[kf];
    const double cdy = hdt / G->dy[kf];
    const size_t km = (size_t)(i - 1) * (size_t)(N - 1) + (size_t)j;
    const double cm = G->cMidy[km];
    const double P = 0.5 * (S(f, i, j + 1) + S(f, i, j)) + 0.5 * (u1 + u0) * G->tgMidy[km] / G->a;
    swo_face F;
    F.hM = 0.5 * (h1 + h0) - cdy1 * (hw1 - hw0);
    F.huM = 0.5 * (hu1 + hu0) - cdy1 * (Uy1 - Uy0) + hdt * P * (0.5 * (hv1 + hv0));
    F.hvM = 0.5 * (hv1 + hv0) - cdy1 * (Vy11 - Vy10) - cdy * (Vy21 - Vy20)
            - hdt * P * (0.5 * (hu1 + hu0));
    F.q = F.huM / F.hM;
    F.hvc = F.hvM * cm;
    F.F1 = (F.hM > 0.0) ? (F.hvM * cm * F.huM / F.hM) : 0.0;
    F.F2 = (F.hM > 0.0) ? (F.hvM * F.hvM / F.hM * cm) : 0.0;
    F.F3 = hg * F.hM * F.hM;
    return F;
}

/* (D_x D_x + D_y D_y) q at interior cell (i, j): numpy.py:87-131 applied to the extension
 * of numpy.py:376-381 (column -1 <- column M-1, column M+3 <- column 3; row -1 <- row 0,
 * row N <- row N-1). */
static double laplacian(const swo_grid *G, const double *q, int i, int j)
{
    const int M = G->M, N = G->N;
    const size_t nx = (size_t)(N - 2);
#define QX(ii) S(q, ((ii) < 0 ? M - 1 : ((ii) > M + 2 ? 3 : (ii))), j)
#define QY(jj) S(q, i, ((jj) < 0 ? 0 : ((jj) > N - 1 ? N - 1 : (jj))))
#define CX(arr, ii) ((arr)[(size_t)(ii) * nx + (size_t)(j - 1)])
#define CY(arr, jj) ((arr)[(size_t)(i - 1) * (size_t)N + (size_t)(jj)])
    const double qxx =
        CX(G->Ax, i) * (CX(G->Ax, i) * QX(i) + CX(G->Bx, i) * QX(i + 1) + CX(G->Cx, i) * QX(i - 1))
        + CX(G->Bx, i) * (CX(G->Ax, i + 1) * QX(i + 1) + CX(G->Bx, i + 1) * QX(i + 2) + CX(G->Cx, i + 1) * QX(i))
        + CX(G->Cx, i) * (CX(G->Ax, i - 1) * QX(i - 1) + CX(G->Bx, i - 1) * QX(i) + CX(G->Cx, i - 1) * QX(i - 2));
    const double qyy =
        CY(G->Ay, j) * (CY(G->Ay, j) * QY(j) + CY(G->By, j) * QY(j + 1) + CY(G->Cy, j) * QY(j - 1))
        + CY(G->By, j) * (CY(G->Ay, j + 1) * QY(j + 1) + CY(G->By, j + 1) * QY(j + 2) + CY(G->Cy, j + 1) * QY(j))
        + CY(G->Cy, j) * (CY(G->Ay, j - 1) * QY(j - 1) + CY(G->By, j - 1) * QY(j) + CY(G->Cy, j - 1) * QY(j - 2));
#undef QX
#undef QY
#undef CX
#undef CY
    return qxx + qyy;
}

/* One Lax-Wendroff update (numpy.py:136-356) optionally followed by the diffusion
 * increment (numpy.py:541-544).  Outputs are the (M+1, N-2) interior arrays; inputs are
 * not modified (tests/numpy_test.py:33-80 of the reference). */
void swo_lw_update(const swo_grid *G, double dt, const double *f, const double *hs,
                   const double *h, const double *u, const double *v, int use_diffusion,
                   double *hnew, double *unew, double *vnew)
{
    const int M = G->M, N = G->N;
    const size_t ni = (size_t)(N - 2);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 1; i <= M + 1; ++i) {
        for (int j = 1; j <= N - 2; ++j) {
            const swo_face x0 = x_face(G, dt, f, h, u, v, i - 1, j);
            const swo_face x1 = x_face(G, dt, f, h, u, v, i, j);
            const swo_face y0 = y_face(G, dt, f, h, u, v, i, j - 1);
            const swo_face y1 = y_face(G, dt, f, h, u, v, i, j);
            const size_t k = (size_t)(i - 1) * ni + (size_t)(j - 1);
            const double hc = S(h, i, j), uc = S(u, i, j), vc = S(v, i, j);
            const double cx = dt / G->dxc[k];
            const double cy1 = dt / G->dy1c[k];
            const double cy = dt / G->dyc[k];
            /* numpy.py:275-284 */
            const double hn = hc - cx * (x1.huM - x0.huM) - cy1 * (y1.hvc - y0.hvc);
            /* numpy.py:298-309 / 330-341 */
            const double Fc = S(f, i, j) + 0.25 * (x0.q + x1.q + y0.q + y1.q) * G->tg[k] / G->a;
            const double hsum = x0.hM + x1.hM + y0.hM + y1.hM;
            const double dxs = S(G->dx, i - 1, j) + S(G->dx, i, j);
            const double dy1s = G->dy1[(size_t)i * (size_t)(N - 1) + (size_t)(j - 1)]
                                + G->dy1[(size_t)i * (size_t)(N - 1) + (size_t)j];
            /* numpy.py:293-318 */
            const double hun = hc * uc - cx * (x1.F1 - x0.F1) - cy1 * (y1.F1 - y0.F1)
                               + dt * Fc * 0.25 * (x0.hvM + x1.hvM + y0.hvM + y1.hvM)
                               - dt * G->g * 0.25 * hsum * (S(hs, i + 1, j) - S(hs, i - 1, j)) / dxs;
            /* numpy.py:324-350 */
            const double hvn = hc * vc - cx * (x1.F2 - x0.F2) - cy1 * (y1.F2 - y0.F2)
                               - cy * (y1.F3 - y0.F3)
                               - dt * Fc * 0.25 * (x0.huM + x1.huM + y0.huM + y1.huM)
                               - dt * G->g * 0.25 * hsum * (S(hs, i, j + 1) - S(hs, i, j - 1)) / dy1s;
            /* numpy.py:353-354 */
            double ho = hn, uo = hun / hn, vo = hvn / hn;
            if (use_diffusion) { /* numpy.py:541-544: old fields, applied to h, u, v */
                const double dn = dt * G->nu;
                ho = ho + dn * laplacian(G, h, i, j);
                uo = uo + dn * laplacian(G, u, i, j);
                vo = vo + dn * laplacian(G, v, i, j);
            }
            hnew[k] = ho;
            unew[k] = uo;
            vnew[k] = vo;
        }
    }
}

/* Boundary conditions, numpy.py:385-406 (in place on the (M+3, N) array q). */
void swo_apply_bcs(int M, int N, double *q, const double *qnew)
{
    const size_t ni = (size_t)(N - 2);
    for (int i = 0; i <= M + 2; ++i) {
        /* row i of concat(qnew[-2:-1], qnew, qnew[1:2]) (numpy.py:402) */
        const int src = (i == 0) ? M - 1 : ((i == M + 2) ? 1 : i - 1);
        for (int j = 1; j <= N - 2; ++j) S(q, i, j) = qnew[(size_t)src * ni + (size_t)(j - 1)];
        S(q, i, 0) = S(q, i, 1);         /* numpy.py:403 */
        S(q, i, N - 1) = S(q, i, N - 2); /* numpy.py:404 */
    }
}

static inline double nanmax(double a, double b)
{ /* np.maximum / ndarray.max propagate NaN */
    if (isnan(a)) return a;
    if (isnan(b)) return b;
    return a > b ? a : b;
}

/* Flux-Jacobian eigenvalue maxima over the FULL ghosted arrays, numpy.py:503-521. */
void swo_cfl_eigen(int M, int N, double g, const double *h, const double *u, const double *v,
                   double *eigenx, double *eigeny)
{
    double ex = -INFINITY, ey = -INFINITY;
    const size_t n = (size_t)(M + 3) * (size_t)N;
    for (size_t k = 0; k < n; ++k) {
        const double cs = sqrt(g * fabs(h[k]));
        const double ax = nanmax(fabs(u[k] - cs), nanmax(fabs(u[k]), fabs(u[k] + cs)));
        const double ay = nanmax(fabs(v[k] - cs), nanmax(fabs(v[k]), fabs(v[k] + cs)));
        ex = nanmax(ex, ax);
        ey = nanmax(ey, ay);
    }
    *eigenx = ex;
    *eigeny = ey;
}

/* dt from the CFL condition, numpy.py:524-525 (np.minimum propagates NaN). */
double swo_cfl_dt(double courant, double dxmin, double dymin, double eigenx, double eigeny)
{
    const double tx = dxmin / eigenx, ty = dymin / eigeny;
    double dtmax;
    if (isnan(tx)) dtmax = tx;
    else if (isnan(ty)) dtmax = ty;
    else dtmax = tx < ty ? tx : ty;
    return courant * dtmax;
}

/* One full loop body, numpy.py:503-549.  h, u, v updated in place; work arrays hold the
 * (M+1, N-2) interior update.  Returns dt; *time is advanced with the final-time clamp
 * of numpy.py:528-532. */
double swo_step(const swo_grid *G, double courant, double dxmin, double dymin,
                double final_time, double *time, const double *f, const double *hs,
                double *h, double *u, double *v, int use_diffusion,
                double *wh, double *wu, double *wv)
{
    double ex, ey;
    swo_cfl_eigen(G->M, G->N, G->g, h, u, v, &ex, &ey);
    double dt = swo_cfl_dt(courant, dxmin, dymin, ex, ey);
    if (*time + dt > final_time) {
        dt = final_time - *time;
        *time = final_time;
    } else {
        *time += dt;
    }
    swo_lw_update(G, dt, f, hs, h, u, v, use_diffusion, wh, wu, wv);
    swo_apply_bcs(G->M, G->N, h, wh);
    swo_apply_bcs(G->M, G->N, u, wu);
    swo_apply_bcs(G->M, G->N, v, wv);
    return dt;
}

int swo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
