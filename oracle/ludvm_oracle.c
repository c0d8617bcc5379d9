/*
 * ludvm_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle + timed CPU baseline).
 *
 * A scalar C restatement of the vortex-velocity hot path of jcatalang/LUDVM, written to be BIT-IDENTICAL
 * to the reference's numpy arithmetic (same operation order, numpy's pairwise-summation tree, np.trapz's
 * formula, LAPACK's 2x2 solve).  Compile with -ffp-contract=off (see oracle/Makefile): a fused
 * multiply-add anywhere except where noted breaks bit parity, and the problem is chaotic.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (ludvm_b200/) never does.
 *
 * Pinning: checked against the unmodified reference run in the development container by
 * tests/test_oracle_vs_reference.py, and against the committed fixtures under tests/golden/ (generated
 * from the reference by tests/golden/make_golden.py) everywhere else.
 *
 * Reference citations are to /root/reference/LUDVM.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * numpy reductions
 * ---------------------------------------------------------------------------------------------- */

/* numpy's DOUBLE_pairwise_sum for a contiguous array. */
static double pw(const double *a, long n)
{
    if (n < 8) {
        double r = -0.0;
        for (long i = 0; i < n; i++) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
        long i, m = n - (n % 8);
        for (i = 8; i < m; i += 8) {
            r0 += a[i + 0]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
            r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
        }
        double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
        for (; i < n; i++) res += a[i];
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return pw(a, n2) + pw(a + n2, n - n2);
}

/* np.sum / np.add.reduce of a contiguous float64 vector: identity 0.0 plus the pairwise tree. */
ORACLE_API double oracle_np_sum(const double *a, long n)
{
    return 0.0 + pw(a, n);
}

/* np.trapz(y, x) = sum( d * (y[1:] + y[:-1]) / 2.0 ), d = diff(x)   (SURVEY.md Appendix A.2). */
ORACLE_API double oracle_np_trapz(const double *y, const double *x, long n)
{
    double t[512];
    double *buf = (n - 1 <= 512) ? t : (double *)malloc(sizeof(double) * (size_t)(n - 1));
    for (long j = 0; j + 1 < n; j++) {
        double d = x[j + 1] - x[j];
        buf[j] = d * (y[j + 1] + y[j]) / 2.0;
    }
    double r = 0.0 + pw(buf, n - 1);
    if (buf != t) free(buf);
    return r;
}

/* np.linalg.solve for a 2x2 system as executed by scipy-openblas' dgesv:
 * partial pivoting, reciprocal scaling of the sub-diagonal, unfused trailing update, fused
 * forward/back substitution. */
ORACLE_API void oracle_solve2x2(const double *A, const double *b, double *x)
{
    double a00 = A[0], a01 = A[1], a10 = A[2], a11 = A[3], b0 = b[0], b1 = b[1];
    if (fabs(a10) > fabs(a00)) {
        double t;
        t = a00; a00 = a10; a10 = t;
        t = a01; a01 = a11; a11 = t;
        t = b0; b0 = b1; b1 = t;
    }
    double l = a10 * (1.0 / a00);
    double u11 = a11 - l * a01;
    double y1 = fma(-l, b0, b1);
    double x1 = y1 / u11;
    double x0 = fma(-a01, x1, b0) / a00;
    x[0] = x0;
    x[1] = x1;
}

/* ------------------------------------------------------------------------------------------------
 * induced_velocity  (LUDVM.py:549-570)
 * ---------------------------------------------------------------------------------------------- */

/* One target row: the per-pair terms in the reference's order (LUDVM.py:565-568), then the row sum
 * of LUDVM.py:569.  g_stride = 0 reproduces broadcasting of a length-1 circulation (LUDVM.py:751). */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void iv_row(const double *g, long g_stride, const double *xw, const double *zw, long nw,
                   double xp, double zp, double vc4, double *tu, double *tw, double *u, double *w)
{
    const double two_pi = 2 * M_PI;
    for (long j = 0; j < nw; j++) {
        double dx = xp - xw[j];
        double dz = zp - zw[j];
        double r2 = dx * dx + dz * dz;
        double den = two_pi * sqrt(r2 * r2 + vc4);
        double ku = dz / den;
        double kw = dx / den;
        double gj = g[j * g_stride];
        tu[j] = gj * ku;
        tw[j] = -(gj * kw);
    }
    *u = 0.0 + pw(tu, nw);
    *w = 0.0 + pw(tw, nw);
}

/* vc4 is the Python value `v_core**4` (0.0 for viscous=False).  nthreads <= 0: all cores. */
ORACLE_API int oracle_induced_velocity(const double *g, long ng, const double *xw, const double *zw, long nw,
                                       const double *xp, const double *zp, long np_, double vc4,
                                       double *u, double *w, int nthreads)
{
    if (nw < 0 || np_ < 0 || (ng != nw && ng != 1)) return -1;
    long gs = (ng == nw) ? 1 : 0;
    if (nw == 0) {
        for (long k = 0; k < np_; k++) { u[k] = 0.0; w[k] = 0.0; }
        return 0;
    }
#ifdef _OPENMP
    int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
    if ((double)np_ * (double)nw < 1.0e5) nt = 1;
#pragma omp parallel num_threads(nt)
#endif
    {
        double *tu = (double *)malloc(sizeof(double) * (size_t)nw * 2);
        double *tw = tu + nw;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
        for (long k = 0; k < np_; k++) iv_row(g, gs, xw, zw, nw, xp[k], zp[k], vc4, tu, tw, &u[k], &w[k]);
        free(tu);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * time_loop  (LUDVM.py:597-1171)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    /* sizes */
    long nt, P, Nc, nfree;
    int method;              /* 0 = Faure (LUDVM.py:741, :916), 1 = Ramesh (LUDVM.py:683, :807) */
    /* scalars, all evaluated by the host in Python exactly as the reference does */
    double dt, Uinf, chord, rho, piv, lespcrit, vc4, ic;
    double maxerror, epsilon;
    long maxiter;
    double a0_init, a1_init; /* fourier[0,0,:2]  (LUDVM.py:645-647) */
    /* per-step kinematics [nt] */
    const double *cos_a, *sin_a, *alpha_dot, *h_dot;
    const double *gp;        /* path['airfoil_gamma_points'] [nt,2,P] */
    const double *le, *te;   /* path['airfoil'][:, :, 0] and [:, :, -1], each [nt,2] */
    /* section tables [P] */
    const double *detadx_p, *eta_p, *x_p, *theta_p, *dtheta, *cos_tp, *sin_tp;
    const double *cosn, *sinn; /* [Nc,P]: np.cos(n*theta_panel), np.sin(n*theta_panel) */
    /* free vortices */
    const double *free_g, *free_xz; /* [nfree], [2,nfree] */
} oracle_sim_in;

typedef struct {
    double *path_tev, *path_lev; /* [nt,2,nt-1] */
    double *path_free;           /* [nt,2,nfree] */
    double *g_tev, *g_lev, *g_bound; /* [nt-1] */
    double *g_airfoil, *gamma_airfoil, *Gamma_airfoil; /* [nt-1,P] */
    double *fourier;             /* [nt,2,Nc] */
    double *lesp, *lesp_prev, *lev_shed; /* [nt] */
    double *Fn, *Fs, *L, *D, *T, *M;     /* [nt] */
    long *itev, *ilev;           /* final self.itev / self.ilev (LUDVM.py:1129-1130) */
} oracle_sim_out;

typedef struct {
    const oracle_sim_in *in;
    const oracle_sim_out *out;
    long nv;        /* nt-1 */
    double *g, *xw, *zw;          /* concatenated wake, capacity 2*nv + nfree */
    double *tu, *tw;              /* row temporaries */
    double *u1, *w1, *W, *T1, *T2, *T3, *tmp; /* [P] */
} sim_ws;

#define TEVX(i) (o->path_tev + ((i) * 2 + 0) * nv)
#define TEVZ(i) (o->path_tev + ((i) * 2 + 1) * nv)
#define LEVX(i) (o->path_lev + ((i) * 2 + 0) * nv)
#define LEVZ(i) (o->path_lev + ((i) * 2 + 1) * nv)
#define FREX(i) (o->path_free + ((i) * 2 + 0) * in->nfree)
#define FREZ(i) (o->path_free + ((i) * 2 + 1) * in->nfree)

/* np.append(np.append(TEV[:nT], LEV[:nL]), FREE) for circulation and row-i coordinates. */
static long gather_wake(sim_ws *s, long i, long nT, long nL)
{
    const oracle_sim_in *in = s->in;
    const oracle_sim_out *o = s->out;
    long nv = s->nv, n = 0;
    for (long j = 0; j < nT; j++, n++) { s->g[n] = o->g_tev[j]; s->xw[n] = TEVX(i)[j]; s->zw[n] = TEVZ(i)[j]; }
    for (long j = 0; j < nL; j++, n++) { s->g[n] = o->g_lev[j]; s->xw[n] = LEVX(i)[j]; s->zw[n] = LEVZ(i)[j]; }
    for (long j = 0; j < in->nfree; j++, n++) { s->g[n] = in->free_g[j]; s->xw[n] = FREX(i)[j]; s->zw[n] = FREZ(i)[j]; }
    return n;
}

static void iv_serial(sim_ws *s, const double *g, long gs, const double *xw, const double *zw, long nw,
                      const double *xp, const double *zp, long np_, double *u, double *w)
{
    for (long k = 0; k < np_; k++) {
        if (nw == 0) { u[k] = 0.0; w[k] = 0.0; continue; }
        iv_row(g, gs, xw, zw, nw, xp[k], zp[k], s->in->vc4, s->tu, s->tw, &u[k], &w[k]);
    }
}

/* airfoil_downwash (LUDVM.py:572-595) for the gathered wake of length nw. */
static void downwash(sim_ws *s, long i, long nw, double *W)
{
    const oracle_sim_in *in = s->in;
    long P = in->P;
    const double *xa = in->gp + (i * 2 + 0) * P, *za = in->gp + (i * 2 + 1) * P;
    double ca = in->cos_a[i], sa = in->sin_a[i], ad = in->alpha_dot[i], hd = in->h_dot[i];
    iv_serial(s, s->g, 1, s->xw, s->zw, nw, xa, za, P, s->u1, s->w1);
    double s1 = in->Uinf * ca + hd * sa;
    double us = in->Uinf * sa;
    double hc = hd * ca;
    for (long j = 0; j < P; j++) {
        double u = s->u1[j] * ca - s->w1[j] * sa;
        double w = s->u1[j] * sa + s->w1[j] * ca;
        W[j] = in->detadx_p[j] * (s1 + u - ad * in->eta_p[j]) - us - ad * (in->x_p[j] - in->piv) + hc - w;
    }
}

/* unit-strength influence of one vortex on the panel normal (LUDVM.py:749-754, :925-934). */
static void unit_influence(sim_ws *s, long i, double xv, double zv, double *T)
{
    const oracle_sim_in *in = s->in;
    long P = in->P;
    const double *xa = in->gp + (i * 2 + 0) * P, *za = in->gp + (i * 2 + 1) * P;
    double one = 1.0, ca = in->cos_a[i], sa = in->sin_a[i];
    iv_serial(s, &one, 0, &xv, &zv, 1, xa, za, P, s->u1, s->w1);
    for (long j = 0; j < P; j++) {
        double ut = s->u1[j] * ca - s->w1[j] * sa;
        double un = s->u1[j] * sa + s->w1[j] * ca;
        T[j] = in->detadx_p[j] * ut - un;
    }
}

static double trapz_theta(sim_ws *s, const double *y) { return oracle_np_trapz(y, s->in->theta_p, s->in->P); }

/* np.trapz(T*(np.cos(theta_panel)-1), theta_panel)  (LUDVM.py:756-757, :936-938) */
static double trapz_cm1(sim_ws *s, const double *T)
{
    for (long j = 0; j < s->in->P; j++) s->tmp[j] = T[j] * (s->in->cos_tp[j] - 1);
    return trapz_theta(s, s->tmp);
}

/* A0 = -1/pi*trapz(W/Uinf), An = 2/pi*trapz(W/Uinf*cos(n theta))  (LUDVM.py:694-695, :769-771) */
static double fourier_coeff(sim_ws *s, const double *W, long n)
{
    const oracle_sim_in *in = s->in;
    if (n == 0) {
        for (long j = 0; j < in->P; j++) s->tmp[j] = W[j] / in->Uinf;
        return (-1 / M_PI) * trapz_theta(s, s->tmp);
    }
    for (long j = 0; j < in->P; j++) s->tmp[j] = W[j] / in->Uinf * in->cosn[n * in->P + j];
    return (2 / M_PI) * trapz_theta(s, s->tmp);
}

/* Kelvin residual used by both Newton loops (LUDVM.py:697-699, :825-827). */
static double kelvin_f(sim_ws *s, double A0, double A1, long nT, long nL, double *cbound)
{
    const oracle_sim_in *in = s->in;
    const oracle_sim_out *o = s->out;
    double cb = in->Uinf * in->chord * M_PI * (A0 + A1 / 2);
    if (cbound) *cbound = cb;
    return cb + oracle_np_sum(o->g_tev, nT) + oracle_np_sum(o->g_lev, nL) + oracle_np_sum(in->free_g, in->nfree) - in->ic;
}

ORACLE_API int oracle_sim_run(const oracle_sim_in *in, const oracle_sim_out *o)
{
    const long nt = in->nt, P = in->P, Nc = in->Nc, nv = nt - 1, nfree = in->nfree;
    if (nt < 2 || P < 2 || Nc < 4 || nfree < 1) return -1;
    sim_ws ws, *s = &ws;
    memset(s, 0, sizeof(ws));
    s->in = in; s->out = o; s->nv = nv;
    long cap = (2 * nv + nfree > P ? 2 * nv + nfree : P) + 8;   /* the row temporaries also serve the P bound vortices as sources */
    s->g = (double *)calloc((size_t)cap * 5, sizeof(double));
    s->xw = s->g + cap; s->zw = s->xw + cap; s->tu = s->zw + cap; s->tw = s->tu + cap;
    s->u1 = (double *)calloc((size_t)P * 7, sizeof(double));
    s->w1 = s->u1 + P; s->W = s->w1 + P; s->T1 = s->W + P; s->T2 = s->T1 + P; s->T3 = s->T2 + P; s->tmp = s->T3 + P;
    double *uw = (double *)calloc((size_t)cap * 4, sizeof(double));
    double *ww = uw + cap, *uf = ww + cap, *wf = uf + cap;

    /* state allocation, LUDVM.py:610-654 (caller zero-fills every output) */
    for (long j = 0; j < nfree; j++) { FREX(0)[j] = in->free_xz[j]; FREZ(0)[j] = in->free_xz[nfree + j]; }
    o->fourier[0] = in->a0_init;
    o->fourier[1] = in->a1_init;
    for (long i = 0; i < nt; i++) o->lev_shed[i] = -1;
    double lespcrit = in->lespcrit;
    const double Uinf = in->Uinf, chord = in->chord, rho = in->rho, dt = in->dt;
    long itev = 0, ilev = 0;
    double sumfree = oracle_np_sum(in->free_g, nfree);

    for (long i = 1; i < nt; i++) {
        double ca = in->cos_a[i], sa = in->sin_a[i], hd = in->h_dot[i];
        double *F = o->fourier + i * 2 * Nc, *Fd = F + Nc, *Fprev = o->fourier + (i - 1) * 2 * Nc;
        const double *xa = in->gp + (i * 2 + 0) * P, *za = in->gp + (i * 2 + 1) * P;

        /* LUDVM.py:664-666 */
        memcpy(TEVX(i), TEVX(i - 1), sizeof(double) * (size_t)itev);
        memcpy(TEVZ(i), TEVZ(i - 1), sizeof(double) * (size_t)itev);
        memcpy(LEVX(i), LEVX(i - 1), sizeof(double) * (size_t)ilev);
        memcpy(LEVZ(i), LEVZ(i - 1), sizeof(double) * (size_t)ilev);
        memcpy(FREX(i), FREX(i - 1), sizeof(double) * (size_t)nfree);
        memcpy(FREZ(i), FREZ(i - 1), sizeof(double) * (size_t)nfree);

        /* TEV placement, LUDVM.py:672-681 */
        if (itev == 0) {
            TEVX(i)[0] = in->te[0] + 0.5 * Uinf * dt;
            TEVZ(i)[0] = in->te[1] + 0.0;
        } else {
            double tex = in->te[i * 2], tez = in->te[i * 2 + 1];
            TEVX(i)[itev] = tex + 1.0 / 3 * (TEVX(i)[itev - 1] - tex);
            TEVZ(i)[itev] = tez + 1.0 / 3 * (TEVZ(i)[itev - 1] - tez);
        }

        if (in->method == 1) {
            /* Ramesh 1-D Newton, LUDVM.py:683-739 */
            double f = 1, shed = -1;
            long niter = 1;
            while (fabs(f) > in->maxerror && niter < in->maxiter) {
                o->g_tev[itev] = shed;
                long nw = gather_wake(s, i, itev + 1, ilev + 1);
                downwash(s, i, nw, s->W);
                double A0 = fourier_coeff(s, s->W, 0), A1 = fourier_coeff(s, s->W, 1);
                f = kelvin_f(s, A0, A1, itev + 1, ilev + 1, NULL);
                o->g_tev[itev] = shed + in->epsilon;
                nw = gather_wake(s, i, itev + 1, ilev + 1);
                downwash(s, i, nw, s->W);
                A0 = fourier_coeff(s, s->W, 0); A1 = fourier_coeff(s, s->W, 1);
                double fdelta = kelvin_f(s, A0, A1, itev + 1, ilev + 1, NULL);
                double fprime = (fdelta - f) / in->epsilon;
                shed = shed - f / fprime;
                o->g_tev[itev] = shed;
                niter++;
            }
            long nw = gather_wake(s, i, itev + 1, ilev + 1);
            downwash(s, i, nw, s->W);
            F[0] = fourier_coeff(s, s->W, 0);
            F[1] = fourier_coeff(s, s->W, 1);
            o->g_bound[itev] = Uinf * chord * M_PI * (F[0] + F[1] / 2);
            for (long n = 2; n < Nc; n++) F[n] = fourier_coeff(s, s->W, n);
            for (long n = 0; n < Nc; n++) Fd[n] = (F[n] - Fprev[n]) / dt;
        } else {
            /* Faure closed form, LUDVM.py:741-773 */
            long nw = gather_wake(s, i, itev, ilev);
            downwash(s, i, nw, s->T1);
            unit_influence(s, i, TEVX(i)[itev], TEVZ(i)[itev], s->T2);
            double I1 = trapz_cm1(s, s->T1), I2 = trapz_cm1(s, s->T2);
            o->g_tev[itev] = -(I1 + oracle_np_sum(o->g_tev, itev) + oracle_np_sum(o->g_lev, ilev) + sumfree - in->ic) / (1 + I2);
            o->g_bound[itev] = I1 + o->g_tev[itev] * I2;
            for (long j = 0; j < P; j++) s->W[j] = s->T1[j] + o->g_tev[itev] * s->T2[j];
            for (long n = 0; n < Nc; n++) F[n] = fourier_coeff(s, s->W, n);
            for (long n = 0; n < Nc; n++) Fd[n] = (F[n] - Fprev[n]) / dt;
        }
        o->lesp_prev[itev] = F[0]; /* LUDVM.py:775 */

        /* LESP test and LEV shedding, LUDVM.py:781-966 */
        if (fabs(F[0]) >= fabs(lespcrit)) {
            double lev_guess = o->g_tev[itev], tev_guess = o->g_tev[itev];
            o->lev_shed[i] = (double)ilev;
            double lex = in->le[i * 2], lez = in->le[i * 2 + 1];
            if (ilev > 0 && o->lev_shed[i - 1] != -1) {
                LEVX(i)[ilev] = lex + 1.0 / 3 * (LEVX(i)[ilev - 1] - lex);
                LEVZ(i)[ilev] = lez + 1.0 / 3 * (LEVZ(i)[ilev - 1] - lez);
            } else {
                LEVX(i)[ilev] = lex;
                LEVZ(i)[ilev] = lez;
            }
            lespcrit = (F[0] < 0) ? -fabs(lespcrit) : fabs(lespcrit);

            if (in->method == 1) {
                /* Ramesh 2-D Newton, LUDVM.py:807-909 */
                double f1 = 0.1, f2 = 0.1;
                long niter = 1;
                while ((fabs(f1) > in->maxerror || fabs(f2) > in->maxerror) && niter < in->maxiter) {
                    double cbound, A0, A1;
                    o->g_tev[itev] = tev_guess; o->g_lev[ilev] = lev_guess;
                    long nw = gather_wake(s, i, itev + 1, ilev + 1);
                    downwash(s, i, nw, s->W);
                    A0 = fourier_coeff(s, s->W, 0); A1 = fourier_coeff(s, s->W, 1);
                    f1 = kelvin_f(s, A0, A1, itev + 1, ilev + 1, &cbound);
                    f2 = lespcrit - A0;
                    o->g_tev[itev] = tev_guess + in->epsilon; o->g_lev[ilev] = lev_guess;
                    nw = gather_wake(s, i, itev + 1, ilev + 1);
                    downwash(s, i, nw, s->W);
                    A0 = fourier_coeff(s, s->W, 0); A1 = fourier_coeff(s, s->W, 1);
                    double f1dT = kelvin_f(s, A0, A1, itev + 1, ilev + 1, NULL), f2dT = lespcrit - A0;
                    o->g_tev[itev] = tev_guess; o->g_lev[ilev] = lev_guess + in->epsilon;
                    nw = gather_wake(s, i, itev + 1, ilev + 1);
                    downwash(s, i, nw, s->W);
                    A0 = fourier_coeff(s, s->W, 0); A1 = fourier_coeff(s, s->W, 1);
                    double f1dL = kelvin_f(s, A0, A1, itev + 1, ilev + 1, NULL), f2dL = lespcrit - A0;
                    double J[4] = { (f1dL - f1) / in->epsilon, (f1dT - f1) / in->epsilon,
                                    (f2dL - f2) / in->epsilon, (f2dT - f2) / in->epsilon };
                    double rhs[2] = { f1, f2 }, sol[2];
                    oracle_solve2x2(J, rhs, sol);
                    lev_guess = lev_guess + (-sol[0]);
                    tev_guess = tev_guess + (-sol[1]);
                    o->g_tev[itev] = tev_guess; o->g_lev[ilev] = lev_guess;
                    o->g_bound[itev] = cbound;
                    niter++;
                }
                long nw = gather_wake(s, i, itev + 1, ilev + 1);
                downwash(s, i, nw, s->W);
                F[0] = fourier_coeff(s, s->W, 0);
                F[1] = fourier_coeff(s, s->W, 1);
                o->g_bound[itev] = Uinf * chord * M_PI * (F[0] + F[1] / 2);
                for (long n = 2; n < Nc; n++) F[n] = fourier_coeff(s, s->W, n);
            } else {
                /* Faure 2x2 linear system, LUDVM.py:916-961 */
                long nw = gather_wake(s, i, itev, ilev);
                downwash(s, i, nw, s->T1);
                unit_influence(s, i, TEVX(i)[itev], TEVZ(i)[itev], s->T2);
                unit_influence(s, i, LEVX(i)[ilev], LEVZ(i)[ilev], s->T3);
                double I1 = trapz_cm1(s, s->T1), I2 = trapz_cm1(s, s->T2), I3 = trapz_cm1(s, s->T3);
                double J1 = (-1 / M_PI) * trapz_theta(s, s->T1);
                double J2 = (-1 / M_PI) * trapz_theta(s, s->T2);
                double J3 = (-1 / M_PI) * trapz_theta(s, s->T3);
                double A[4] = { 1 + I2, 1 + I3, J2, J3 };
                double b[2] = { -(I1 + oracle_np_sum(o->g_tev, itev) + oracle_np_sum(o->g_lev, ilev) + sumfree - in->ic),
                                lespcrit - J1 };
                double sol[2];
                oracle_solve2x2(A, b, sol);
                o->g_tev[itev] = sol[0];
                o->g_lev[ilev] = sol[1];
                o->g_bound[itev] = I1 + sol[0] * I2 + sol[1] * I3;
                for (long j = 0; j < P; j++) s->W[j] = s->T1[j] + sol[0] * s->T2[j] + sol[1] * s->T3[j];
                F[0] = J1 + sol[0] * J2 + sol[1] * J3;
                for (long n = 1; n < Nc; n++) F[n] = fourier_coeff(s, s->W, n);
            }
        }
        o->lesp[itev] = F[0]; /* LUDVM.py:971 */

        /* bound-vortex distribution, LUDVM.py:986-1010 */
        double *dG = o->g_airfoil + itev * P, *gam = o->gamma_airfoil + itev * P, *Gam = o->Gamma_airfoil + itev * P;
        for (long j = 0; j < P; j++) {
            double term2 = 0;
            for (long n = 1; n < Nc; n++) term2 = F[n] * in->sinn[n * P + j] + term2;
            double term1 = F[0] * (1 + in->cos_tp[j]) / in->sin_tp[j];
            double gamma = 2 * Uinf * (term1 + term2);
            dG[j] = gamma * chord / 2 * in->sin_tp[j] * in->dtheta[j];
            gam[j] = gamma;
        }
        for (long j = 0; j < P; j++) Gam[j] = oracle_np_sum(dG, j + 1);

        /* loads, LUDVM.py:1035-1090 */
        long nw = gather_wake(s, i, itev + 1, ilev + 1);
        iv_serial(s, s->g, 1, s->xw, s->zw, nw, xa, za, P, s->u1, s->w1);
        for (long j = 0; j < P; j++) {
            double u = s->u1[j] * ca - s->w1[j] * sa;
            s->tmp[j] = u * gam[j];
            s->T3[j] = u * gam[j] * in->x_p[j];
        }
        double A0 = F[0], A1 = F[1], A2 = F[2], A0d = Fd[0], A1d = Fd[1], A2d = Fd[2], A3d = Fd[3];
        double vrel = Uinf * ca + hd * sa;
        o->Fn[i] = rho * M_PI * chord * Uinf * (vrel * (A0 + 0.5 * A1) + chord * (3.0 / 4 * A0d + 1.0 / 4 * A1d + 1.0 / 8 * A2d))
                   + rho * oracle_np_trapz(s->tmp, in->x_p, P);
        o->Fs[i] = rho * M_PI * chord * (Uinf * Uinf) * (A0 * A0);
        o->L[i] = o->Fn[i] * ca + o->Fs[i] * sa;
        o->D[i] = o->Fn[i] * sa - o->Fs[i] * ca;
        o->T[i] = -o->D[i];
        o->M[i] = in->piv * o->Fn[i]
                  - rho * M_PI * (chord * chord) * Uinf * (vrel * (1.0 / 4 * A0 + 1.0 / 4 * A1 - 1.0 / 8 * A2)
                        + chord * (7.0 / 16 * A0d + 3.0 / 16 * A1d + 1.0 / 16 * A2d - 1.0 / 64 * A3d))
                  - rho * oracle_np_trapz(s->T3, in->x_p, P);

        /* convection, LUDVM.py:1095-1127: every velocity at pre-update positions */
        long nT = itev + 1, nL = ilev + 1;
        iv_serial(s, s->g, 1, s->xw, s->zw, nw, s->xw, s->zw, nw, uw, ww);
        iv_serial(s, dG, 1, xa, za, P, s->xw, s->zw, nw, uf, wf);
        for (long j = 0; j < nT; j++) {
            TEVX(i)[j] = TEVX(i)[j] + dt * (uw[j] + uf[j]);
            TEVZ(i)[j] = TEVZ(i)[j] + dt * (ww[j] + wf[j]);
        }
        for (long j = 0; j < nL; j++) {
            LEVX(i)[j] = LEVX(i)[j] + dt * (uw[nT + j] + uf[nT + j]);
            LEVZ(i)[j] = LEVZ(i)[j] + dt * (ww[nT + j] + wf[nT + j]);
        }
        for (long j = 0; j < nfree; j++) {
            FREX(i)[j] = FREX(i)[j] + dt * (uw[nT + nL + j] + uf[nT + nL + j]);
            FREZ(i)[j] = FREZ(i)[j] + dt * (ww[nT + nL + j] + wf[nT + nL + j]);
        }
        *o->itev = itev;
        *o->ilev = ilev;

        /* LUDVM.py:1166-1169 */
        if (o->lev_shed[i] != -1) ilev++;
        itev++;
    }
    free(uw);
    free(s->u1);
    free(s->g);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * flowfield vorticity stencil  (LUDVM.py:1222-1292)
 * x, z: [nx,nz] meshgrid('ij'); u, w, ome: [ns,nx,nz]
 * ---------------------------------------------------------------------------------------------- */
ORACLE_API void oracle_vorticity(const double *x, const double *z, const double *u, const double *w,
                                 long ns, long nx, long nz, double *ome)
{
#define IX(i, j) ((i) * nz + (j))
    for (long s = 0; s < ns; s++) {
        const double *us = u + s * nx * nz, *wsl = w + s * nx * nz;
        double *os = ome + s * nx * nz;
        for (long i = 0; i < nx; i++) {
            long ip = (i + 1 < nx) ? i + 1 : i, im = (i > 0) ? i - 1 : i;
            for (long j = 0; j < nz; j++) {
                long jp = (j + 1 < nz) ? j + 1 : j, jm = (j > 0) ? j - 1 : j;
                double dx = x[IX(ip, j)] - x[IX(im, j)];
                double dz = z[IX(i, jp)] - z[IX(i, jm)];
                double dw = wsl[IX(ip, j)] - wsl[IX(im, j)];
                double du = us[IX(i, jp)] - us[IX(i, jm)];
                os[IX(i, j)] = dw / dx - du / dz;
            }
        }
    }
#undef IX
}
