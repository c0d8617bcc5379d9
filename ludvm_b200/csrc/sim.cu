// sim.cu -- the on-device time step of LUDVM.time_loop (LUDVM.py:597-1171), replayed as a CUDA graph.
//
// This translation unit is compiled with -fmad=false: the scalar phases of the step are written as plain C
// expressions in the reference's operation order and must not be contracted into FMAs.
// The fast pair arithmetic uses explicit fma() and is unaffected.
//
// One time step i (Faure method) = four kernels; the vortex state (x, z, Gamma of TEV | LEV | FREE) stays in
// HBM in one SoA with three fixed segments, addressed through the logical->physical map of SrcView so that the
// reference's np.append(TEV[:..], LEV[:..], FREE) ordering -- which fixes the shape of numpy's summation tree --
// is reproduced without moving data:
//
//   k_wake_on_foil   partial sums of the existing wake at the P gamma points           (LUDVM.py:743-746)
//   k_solve          1 CTA: TEV placement, T1/T2/T3, Kelvin + LESP 2x2 solve, Fourier coefficients, bound
//                    vortex distribution                                              (LUDVM.py:672-681, 741-1010)
//   k_conv_partials  partial sums of the updated wake at the gamma points (loads) and at every wake vortex
//                    (convection), plus the bound vortices on the wake                (LUDVM.py:1049-1054, 1095-1124)
//   k_finish         block 0: loads Fn, Fs, L, D, T, M; other blocks: fold partials, forward-Euler update in
//                    place, history snapshot                                          (LUDVM.py:1069-1090, 1108-1127)
//
// The step index is read from device memory (counters[0] + offset baked into the node), so one captured graph of
// K steps is replayed for the whole run; the number of vortices is read from device memory too, kernels are
// grid-stride over tasks, and graphs are cached per power-of-two bracket of the vortex-count upper bound.
#include <algorithm>
#include <map>

#include "biot_savart.cuh"

namespace ludvm {

#define SIM_DMAX 11          // at most 2^11 tree nodes per row are evaluated by separate warps
#define SOLVE_THREADS 256
#define SUM_NODES_MAX 4096   // block_np_sum: tree nodes held in shared memory

struct SimDev {
    int nt, P, Nc, nfree, nv, method, mode, store_history;
    int target_warps;  // parallelism target used to pick split depths (same value in every kernel)
    double dt, Uinf, chord, rho, piv, vc4, ic, sum_free, maxerror, epsilon;
    int maxiter;
    // tables
    const double *cos_a, *sin_a, *alpha_dot, *h_dot, *gp, *le, *te;
    const double *detadx_p, *eta_p, *x_p, *theta_p, *dtheta, *cos_tp, *sin_tp, *cosn, *sinn;
    // resident vortex state: [TEV nv | LEV nv | FREE nfree]
    double *wx, *wz, *wg;
    // results
    double *g_bound, *g_airfoil, *gamma_airfoil, *Gamma_airfoil, *fourier, *lesp, *lesp_prev, *lev_shed;
    double *Fn, *Fs, *L, *D, *T, *M;
    double *path_tev, *path_lev, *path_free;  // nullable
    // step bookkeeping
    long long *counters;  // [0] steps done, [1] itev of the last step, [2] ilev of the last step, [3] error flags
    int *ilev_arr;        // [nt+1]: ilev at the start of step i
    double *lespcrit_cur; // sign-carrying LESPcrit (LUDVM.py:802-805)
    // scratch
    double *pa_u, *pa_w;  // phase-C partials   [2^d1][P]
    double *pb_u, *pb_w;  // loads+convection partials [2^d2][P + Nw2]
    double *foil_u, *foil_w;  // bound vortices on the wake [Nw2]
};

__host__ __device__ __forceinline__ int ilog2_ceil_i(int v)
{
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

// Split depth for `nrows` target rows against n sources (identical on host and device, in every kernel).
__host__ __device__ __forceinline__ int sim_depth(int n, int nrows, int target_warps)
{
    int quads = (nrows + 3) >> 2;
    int want = ilog2_ceil_i(max(1, target_warps / max(1, quads)));
    int d = min(pw_max_depth(n), want);
    return min(d, SIM_DMAX);
}
__host__ __device__ __forceinline__ int sim_chunks(int n, int nrows, int target_warps)
{
    int quads = (nrows + 3) >> 2;
    int c = max(1, target_warps / max(1, quads));
    c = min(c, (n + 63) / 64);
    return min(max(c, 1), 1 << SIM_DMAX);
}

struct Step {
    int i, itev, ilev;
};

__device__ __forceinline__ bool step_begin(const SimDev &S, int s, Step &st)
{
    long long i = S.counters[0] + s + 1;
    if (i >= S.nt) return false;
    st.i = (int)i;
    st.itev = st.i - 1;
    st.ilev = S.ilev_arr[st.i];
    return true;
}

// wake = TEV[:nT] ++ LEV[:nL] ++ FREE
__device__ __forceinline__ SrcView wake_view(const SimDev &S, int nT, int nL)
{
    SrcView v;
    v.x = S.wx; v.z = S.wz; v.g = S.wg; v.vc4 = nullptr; v.vc4s = S.vc4;
    v.n0 = nT; v.n01 = nT + nL; v.n = nT + nL + S.nfree;
    v.o1 = S.nv; v.o2 = 2 * S.nv; v.gstride = 1;
    return v;
}

struct TgtGamma {  // the P gamma points of step i
    const double *x, *z;
    __device__ __forceinline__ void get(int r, double &xp, double &zp) const { xp = x[r]; zp = z[r]; }
};

struct TgtGammaWake {  // rows [0,P): gamma points; rows [P, P+Nw): wake vortices in logical order
    const double *gx, *gz;
    int P;
    SrcView W;
    __device__ __forceinline__ void get(int r, double &xp, double &zp) const
    {
        if (r < P) {
            xp = gx[r];
            zp = gz[r];
        } else {
            int p = W.phys(r - P);
            xp = W.x[p];
            zp = W.z[p];
        }
    }
};

struct TgtWake {
    SrcView W;
    __device__ __forceinline__ void get(int r, double &xp, double &zp) const
    {
        int p = W.phys(r);
        xp = W.x[p];
        zp = W.z[p];
    }
};

// ---------------------------------------------------------------------------------------------------
// kernel 1: existing wake on the gamma points (LUDVM.py:743-746 -> :584)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wake_on_foil(SimDev S, int s)
{
    Step st;
    if (!step_begin(S, s, st)) return;
    SrcView W = wake_view(S, st.itev, st.ilev);
    TgtGamma T{S.gp + ((size_t)st.i * 2 + 0) * S.P, S.gp + ((size_t)st.i * 2 + 1) * S.P};
    int lane = threadIdx.x & 31;
    long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    long nquads = (S.P + 3) >> 2;
    if (S.mode == LUDVM_EXACT_F64) {
        int d = sim_depth(W.n, S.P, S.target_warps);
        for (long t = gw; t < (nquads << d); t += nwarps) exact_rows_warp_task(W, T, S.P, d, t, lane, S.pa_u, S.pa_w);
    } else {
        int c = sim_chunks(W.n, S.P, S.target_warps);
        for (long t = gw; t < nquads * c; t += nwarps) fast_rows_warp_task(W, T, S.P, c, t, lane, S.pa_u, S.pa_w);
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel 2: the scalar phases (one CTA)
// ---------------------------------------------------------------------------------------------------

// np.sum(a[:n]) by the whole block: tree nodes at depth d are summed by one thread each, thread 0 folds them.
__device__ double block_np_sum(const double *a, int n, double *s_nodes, double *s_out)
{
    __syncthreads();
    int d = min(pw_max_depth(n), ilog2_ceil_i(SUM_NODES_MAX) - 0);
    while ((1 << d) > SUM_NODES_MAX) d--;
    int nn = 1 << d;
    auto f = [a](int j) { return a[j]; };
    for (int b = threadIdx.x; b < nn; b += blockDim.x) {
        int off, len;
        pw_node(n, d, b, off, len);
        s_nodes[b] = pw_seq(f, off, len);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double stck[PW_MAX_STACK];
        int sp = 0;
        for (int i = 0; i < nn; i++) {
            double v = s_nodes[i];
            for (int k = i; k & 1; k >>= 1) v = __dadd_rn(stck[--sp], v);
            stck[sp++] = v;
        }
        *s_out = 0.0 + stck[0];
    }
    __syncthreads();
    return *s_out;
}

// np.trapz(y, x) over P points by one thread.
template <class Y>
__device__ double trapz_seq(Y y, const double *x, int P)
{
    auto term = [&](int j) {
        double d = x[j + 1] - x[j];
        return d * (y(j + 1) + y(j)) / 2.0;
    };
    return 0.0 + pw_seq(term, 0, P - 1);
}

// unit-strength influence of a vortex at (xv, zv) on panel j (LUDVM.py:749-754): T = detadx*ut - un
__device__ __forceinline__ double unit_T(const SimDev &S, double xa, double za, double xv, double zv, double ca,
                                         double sa, double detadx)
{
    double tu, tw;
    pair_exact(xa, za, xv, zv, 1.0, S.vc4, tu, tw);
    double u = 0.0 + (-0.0 + tu), w = 0.0 + (-0.0 + tw);  // np.sum over one element
    double ut = u * ca - w * sa;
    double un = u * sa + w * ca;
    return detadx * ut - un;
}

// LAPACK dgesv on a 2x2 system
__device__ __forceinline__ void solve2x2(double a00, double a01, double a10, double a11, double b0, double b1,
                                         double &x0, double &x1)
{
    if (fabs(a10) > fabs(a00)) {
        double t;
        t = a00; a00 = a10; a10 = t;
        t = a01; a01 = a11; a11 = t;
        t = b0; b0 = b1; b1 = t;
    }
    double l = a10 * (1.0 / a00);
    double u11 = a11 - l * a01;
    double y1 = fma(-l, b0, b1);
    x1 = y1 / u11;
    x0 = fma(-a01, x1, b0) / a00;
}

#define PI_D 3.141592653589793

__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(SimDev S, int s)
{
    extern __shared__ double sm[];
    Step st;
    if (!step_begin(S, s, st)) return;
    const int P = S.P, Nc = S.Nc, tid = threadIdx.x, nth = blockDim.x;
    const int i = st.i, itev = st.itev, ilev = st.ilev, nv = S.nv;
    double *u1 = sm, *w1 = u1 + P, *T1 = w1 + P, *T2 = T1 + P, *T3 = T2 + P, *W = T3 + P, *dG = W + P;
    double *A = dG + P, *sc = A + Nc, *nodes = sc + 32;
    // sc[]: 0 xt, 1 zt, 2 sT, 3 sL, 4 I1, 5 I2, 6 I3, 7 J1, 8 J2, 9 J3, 10 gtev, 11 glev, 12 shed, 13 xl, 14 zl,
    //       15 lespcrit, 16 scratch
    const double ca = S.cos_a[i], sa = S.sin_a[i], ad = S.alpha_dot[i], hd = S.h_dot[i];
    const double *xa = S.gp + ((size_t)i * 2 + 0) * P, *za = S.gp + ((size_t)i * 2 + 1) * P;
    const double Uinf = S.Uinf, chord = S.chord, dt = S.dt;

    // fold the phase-C partials (LUDVM.py:584)
    {
        int n1 = itev + ilev + S.nfree;
        if (S.mode == LUDVM_EXACT_F64) {
            int d = sim_depth(n1, P, S.target_warps);
            for (int j = tid; j < P; j += nth) {
                u1[j] = exact_combine_row(S.pa_u, P, j, d);
                w1[j] = exact_combine_row(S.pa_w, P, j, d);
            }
        } else {
            int c = sim_chunks(n1, P, S.target_warps);
            for (int j = tid; j < P; j += nth) {
                u1[j] = fast_combine_row(S.pa_u, P, j, c);
                w1[j] = fast_combine_row(S.pa_w, P, j, c);
            }
        }
    }
    // TEV placement (LUDVM.py:672-681)
    if (tid == 0) {
        double xt, zt;
        if (itev == 0) {
            xt = S.te[0] + 0.5 * Uinf * dt;
            zt = S.te[1] + 0.0;
        } else {
            double tex = S.te[(size_t)i * 2], tez = S.te[(size_t)i * 2 + 1];
            xt = tex + 1.0 / 3 * (S.wx[itev - 1] - tex);
            zt = tez + 1.0 / 3 * (S.wz[itev - 1] - tez);
        }
        sc[0] = xt;
        sc[1] = zt;
        S.wx[itev] = xt;
        S.wz[itev] = zt;
        sc[15] = *S.lespcrit_cur;
    }
    // np.sum(circulation['TEV'][:itev]), np.sum(circulation['LEV'][:ilev])  (LUDVM.py:758-759)
    double sT = block_np_sum(S.wg, itev, nodes, &sc[2]);
    double sL = block_np_sum(S.wg + nv, ilev, nodes, &sc[3]);
    const double xt = sc[0], zt = sc[1];
    // T1 = airfoil_downwash(existing wake) (LUDVM.py:587-593), T2 = unit TEV influence (LUDVM.py:749-754)
    {
        double s1 = Uinf * ca + hd * sa, us = Uinf * sa, hc = hd * ca;
        for (int j = tid; j < P; j += nth) {
            double u = u1[j] * ca - w1[j] * sa;
            double w = u1[j] * sa + w1[j] * ca;
            T1[j] = S.detadx_p[j] * (s1 + u - ad * S.eta_p[j]) - us - ad * (S.x_p[j] - S.piv) + hc - w;
            T2[j] = unit_T(S, xa[j], za[j], xt, zt, ca, sa, S.detadx_p[j]);
        }
    }
    __syncthreads();
    // I1, I2 (LUDVM.py:756-757) on two different warps
    if (tid == 0) sc[4] = trapz_seq([&](int j) { return T1[j] * (S.cos_tp[j] - 1); }, S.theta_p, P);
    if (tid == 32) sc[5] = trapz_seq([&](int j) { return T2[j] * (S.cos_tp[j] - 1); }, S.theta_p, P);
    __syncthreads();
    if (tid == 0) {
        double I1 = sc[4], I2 = sc[5];
        sc[10] = -(I1 + sT + sL + S.sum_free - S.ic) / (1 + I2);  // LUDVM.py:758-760
        sc[11] = 0.0;
        sc[16] = I1 + sc[10] * I2;                                // circulation['bound'], LUDVM.py:762
    }
    __syncthreads();
    for (int j = tid; j < P; j += nth) W[j] = T1[j] + sc[10] * T2[j];  // LUDVM.py:767
    __syncthreads();
    // Fourier coefficients and their derivatives (LUDVM.py:769-773); thread n*8 keeps them on distinct warps
    double *F = S.fourier + (size_t)i * 2 * Nc;
    const double *Fprev = S.fourier + (size_t)(i - 1) * 2 * Nc;
    for (int n = tid; n < Nc; n += nth) {
        double a;
        if (n == 0) a = (-1 / PI_D) * trapz_seq([&](int j) { return W[j] / Uinf; }, S.theta_p, P);
        else {
            const double *cn = S.cosn + (size_t)n * P;
            a = (2 / PI_D) * trapz_seq([&](int j) { return W[j] / Uinf * cn[j]; }, S.theta_p, P);
        }
        A[n] = a;
        F[Nc + n] = (a - Fprev[n]) / dt;
    }
    __syncthreads();
    // LESP test (LUDVM.py:775-781)
    if (tid == 0) {
        S.lesp_prev[itev] = A[0];
        sc[12] = (fabs(A[0]) >= fabs(sc[15])) ? 1.0 : 0.0;
    }
    __syncthreads();
    const bool shed = sc[12] != 0.0;
    if (shed) {
        if (tid == 0) {  // LEV placement (LUDVM.py:784-805)
            double lex = S.le[(size_t)i * 2], lez = S.le[(size_t)i * 2 + 1], xl = lex, zl = lez;
            if (ilev > 0 && S.lev_shed[i - 1] != -1.0) {
                xl = lex + 1.0 / 3 * (S.wx[nv + ilev - 1] - lex);
                zl = lez + 1.0 / 3 * (S.wz[nv + ilev - 1] - lez);
            }
            sc[13] = xl;
            sc[14] = zl;
            sc[15] = (A[0] < 0) ? -fabs(sc[15]) : fabs(sc[15]);
            S.lev_shed[i] = (double)ilev;
        }
        __syncthreads();
        for (int j = tid; j < P; j += nth) T3[j] = unit_T(S, xa[j], za[j], sc[13], sc[14], ca, sa, S.detadx_p[j]);
        __syncthreads();
        // I3, J1, J2, J3 (LUDVM.py:936-942)
        if (tid == 0) sc[6] = trapz_seq([&](int j) { return T3[j] * (S.cos_tp[j] - 1); }, S.theta_p, P);
        if (tid == 32) sc[7] = (-1 / PI_D) * trapz_seq([&](int j) { return T1[j]; }, S.theta_p, P);
        if (tid == 64) sc[8] = (-1 / PI_D) * trapz_seq([&](int j) { return T2[j]; }, S.theta_p, P);
        if (tid == 96) sc[9] = (-1 / PI_D) * trapz_seq([&](int j) { return T3[j]; }, S.theta_p, P);
        __syncthreads();
        if (tid == 0) {  // LUDVM.py:945-959
            double I1 = sc[4], I2 = sc[5], I3 = sc[6], J1 = sc[7], J2 = sc[8], J3 = sc[9];
            double b1 = -(I1 + sT + sL + S.sum_free - S.ic), b2 = sc[15] - J1, x0, x1;
            solve2x2(1 + I2, 1 + I3, J2, J3, b1, b2, x0, x1);
            sc[10] = x0;
            sc[11] = x1;
            sc[16] = I1 + x0 * I2 + x1 * I3;
            A[0] = J1 + x0 * J2 + x1 * J3;
        }
        __syncthreads();
        for (int j = tid; j < P; j += nth) W[j] = T1[j] + sc[10] * T2[j] + sc[11] * T3[j];
        __syncthreads();
        for (int n = 1 + tid; n < Nc; n += nth) {  // LUDVM.py:960-961 (derivatives keep their values, :963-966)
            const double *cn = S.cosn + (size_t)n * P;
            A[n] = (2 / PI_D) * trapz_seq([&](int j) { return W[j] / Uinf * cn[j]; }, S.theta_p, P);
        }
        __syncthreads();
    }
    // commit the step's circulations and bookkeeping
    if (tid == 0) {
        S.wg[itev] = sc[10];
        if (shed) {
            S.wg[nv + ilev] = sc[11];
            S.wx[nv + ilev] = sc[13];
            S.wz[nv + ilev] = sc[14];
        } else if (ilev < nv) {  // untouched slot of row i: zero circulation at the origin
            S.wx[nv + ilev] = 0.0;
            S.wz[nv + ilev] = 0.0;
        }
        S.g_bound[itev] = sc[16];
        S.lesp[itev] = A[0];  // LUDVM.py:971
        S.ilev_arr[i + 1] = ilev + (shed ? 1 : 0);
        *S.lespcrit_cur = sc[15];
    }
    for (int n = tid; n < Nc; n += nth) F[n] = A[n];
    // bound-vortex distribution (LUDVM.py:986-1010)
    for (int j = tid; j < P; j += nth) {
        double term2 = 0;
        for (int n = 1; n < Nc; n++) term2 = A[n] * S.sinn[(size_t)n * P + j] + term2;
        double term1 = A[0] * (1 + S.cos_tp[j]) / S.sin_tp[j];
        double gamma = 2 * Uinf * (term1 + term2);
        double dg = gamma * chord / 2 * S.sin_tp[j] * S.dtheta[j];
        dG[j] = dg;
        S.g_airfoil[(size_t)itev * P + j] = dg;
        S.gamma_airfoil[(size_t)itev * P + j] = gamma;
    }
    __syncthreads();
    for (int j = tid; j < P; j += nth) {
        auto f = [&](int k) { return dG[k]; };
        S.Gamma_airfoil[(size_t)itev * P + j] = 0.0 + pw_seq(f, 0, j + 1);
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel 3: updated wake on (gamma points ++ wake) and bound vortices on the wake
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_conv_partials(SimDev S, int s)
{
    Step st;
    if (!step_begin(S, s, st)) return;
    const int P = S.P;
    SrcView W = wake_view(S, st.itev + 1, S.ilev_arr[st.i] + 1);
    const double *gx = S.gp + ((size_t)st.i * 2 + 0) * P, *gz = S.gp + ((size_t)st.i * 2 + 1) * P;
    TgtGammaWake TA{gx, gz, P, W};
    TgtWake TW{W};
    SrcView Fo = make_src(S.g_airfoil + (size_t)st.itev * P, 1, gx, gz, nullptr, S.vc4, P);
    const int nrows = P + W.n;
    int lane = threadIdx.x & 31;
    long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    long nquadsA = (nrows + 3) >> 2, nquadsW = (W.n + 3) >> 2;
    if (S.mode == LUDVM_EXACT_F64) {
        int d = sim_depth(W.n, nrows, S.target_warps);
        long nA = nquadsA << d;
        for (long t = gw; t < nA + nquadsW; t += nwarps) {
            if (t < nA) exact_rows_warp_task(W, TA, nrows, d, t, lane, S.pb_u, S.pb_w);
            else exact_rows_warp_task(Fo, TW, W.n, 0, t - nA, lane, S.foil_u, S.foil_w);
        }
    } else {
        int c = sim_chunks(W.n, nrows, S.target_warps);
        long nA = nquadsA * c;
        for (long t = gw; t < nA + nquadsW; t += nwarps) {
            if (t < nA) fast_rows_warp_task(W, TA, nrows, c, t, lane, S.pb_u, S.pb_w);
            else fast_rows_warp_task(Fo, TW, W.n, 1, t - nA, lane, S.foil_u, S.foil_w);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel 4: loads (block 0) and convection update + history (other blocks)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_finish(SimDev S, int s)
{
    extern __shared__ double sm[];
    Step st;
    if (!step_begin(S, s, st)) return;
    const int P = S.P, i = st.i, itev = st.itev, ilev = st.ilev, nv = S.nv, tid = threadIdx.x, nth = blockDim.x;
    const int nT = itev + 1, nL = ilev + 1;
    SrcView W = wake_view(S, nT, nL);
    const int nrows = P + W.n;
    const bool exact = S.mode == LUDVM_EXACT_F64;
    const int fold = exact ? sim_depth(W.n, nrows, S.target_warps) : sim_chunks(W.n, nrows, S.target_warps);

    if (blockIdx.x == 0) {  // loads, LUDVM.py:1035-1090
        double *ug = sm, *ugx = ug + P, *sc = ugx + P;
        const double ca = S.cos_a[i], sa = S.sin_a[i], hd = S.h_dot[i];
        const double *gam = S.gamma_airfoil + (size_t)itev * P;
        for (int j = tid; j < P; j += nth) {
            double u1 = exact ? exact_combine_row(S.pb_u, nrows, j, fold) : fast_combine_row(S.pb_u, nrows, j, fold);
            double w1 = exact ? exact_combine_row(S.pb_w, nrows, j, fold) : fast_combine_row(S.pb_w, nrows, j, fold);
            double u = u1 * ca - w1 * sa;
            ug[j] = u * gam[j];
            ugx[j] = u * gam[j] * S.x_p[j];
        }
        __syncthreads();
        if (tid == 0) sc[0] = trapz_seq([&](int j) { return ug[j]; }, S.x_p, P);
        if (tid == 32) sc[1] = trapz_seq([&](int j) { return ugx[j]; }, S.x_p, P);
        __syncthreads();
        if (tid == 0) {
            const double *F = S.fourier + (size_t)i * 2 * S.Nc, *Fd = F + S.Nc;
            const double rho = S.rho, chord = S.chord, Uinf = S.Uinf;
            double A0 = F[0], A1 = F[1], A2 = F[2], A0d = Fd[0], A1d = Fd[1], A2d = Fd[2], A3d = Fd[3];
            double vrel = Uinf * ca + hd * sa;
            double Fn = rho * PI_D * chord * Uinf *
                            (vrel * (A0 + 0.5 * A1) + chord * (3.0 / 4 * A0d + 1.0 / 4 * A1d + 1.0 / 8 * A2d)) +
                        rho * sc[0];
            double Fs = rho * PI_D * chord * (Uinf * Uinf) * (A0 * A0);
            double Lf = Fn * ca + Fs * sa;
            double Df = Fn * sa - Fs * ca;
            double Mo = S.piv * Fn -
                        rho * PI_D * (chord * chord) * Uinf *
                            (vrel * (1.0 / 4 * A0 + 1.0 / 4 * A1 - 1.0 / 8 * A2) +
                             chord * (7.0 / 16 * A0d + 3.0 / 16 * A1d + 1.0 / 16 * A2d - 1.0 / 64 * A3d)) -
                        rho * sc[1];
            S.Fn[i] = Fn; S.Fs[i] = Fs; S.L[i] = Lf; S.D[i] = Df; S.T[i] = -Df; S.M[i] = Mo;
            S.counters[1] = itev;
            S.counters[2] = ilev;
        }
        return;
    }
    // convection, LUDVM.py:1095-1127: x += dt*(u_wake + u_foil) for TEV[:nT], LEV[:nL], FREE, in place
    const double dt = S.dt;
    for (long r = (long)(blockIdx.x - 1) * nth + tid; r < W.n; r += (long)(gridDim.x - 1) * nth) {
        int row = P + (int)r;
        double uw, ww, uf, wf;
        if (exact) {
            uw = exact_combine_row(S.pb_u, nrows, row, fold);
            ww = exact_combine_row(S.pb_w, nrows, row, fold);
            uf = 0.0 + S.foil_u[r];
            wf = 0.0 + S.foil_w[r];
        } else {
            uw = fast_combine_row(S.pb_u, nrows, row, fold);
            ww = fast_combine_row(S.pb_w, nrows, row, fold);
            uf = S.foil_u[r];
            wf = S.foil_w[r];
        }
        int p = W.phys((int)r);
        double xn = S.wx[p] + dt * (uw + uf);
        double zn = S.wz[p] + dt * (ww + wf);
        S.wx[p] = xn;
        S.wz[p] = zn;
        if (S.store_history) {
            double *hx, *hz;
            if (r < nT) {
                hx = S.path_tev + ((size_t)i * 2) * nv + r;
                hz = hx + nv;
            } else if (r < nT + nL) {
                hx = S.path_lev + ((size_t)i * 2) * nv + (r - nT);
                hz = hx + nv;
            } else {
                hx = S.path_free + ((size_t)i * 2) * S.nfree + (r - nT - nL);
                hz = hx + S.nfree;
            }
            *hx = xn;
            *hz = zn;
        }
    }
}

__global__ void k_advance(SimDev S, int k)
{
    long long v = S.counters[0] + k;
    S.counters[0] = v > S.nt - 1 ? S.nt - 1 : v;
}

__global__ void k_fill(double *p, size_t n, double v)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) p[idx] = v;
}

}  // namespace ludvm

using namespace ludvm;

// ---------------------------------------------------------------------------------------------------
// host object
// ---------------------------------------------------------------------------------------------------
struct ludvm_sim {
    ludvm_ctx *ctx = nullptr;
    ludvm_sim_params p{};
    SimDev d{};
    std::vector<void *> allocs;
    long steps_enqueued = 0;
    int K = 50;
    std::map<int, cudaGraphExec_t> graphs;  // (bracket, length) -> instantiated graph
    size_t solve_smem = 0, finish_smem = 0;
    cudaStream_t cap_stream = nullptr;  // private stream used only to record graphs (the context's stream may be
                                        // the legacy default stream, which cannot be captured)
};

namespace ludvm {

static int dev_alloc(ludvm_sim *s, size_t n_doubles, double **out, bool zero = true)
{
    void *p = nullptr;
    size_t bytes = std::max<size_t>(n_doubles, 1) * sizeof(double);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(LUDVM_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    s->allocs.push_back(p);
    if (zero) CUDA_TRY(cudaMemsetAsync(p, 0, bytes, s->ctx->stream));
    *out = (double *)p;
    return LUDVM_OK;
}

static int upload(ludvm_sim *s, const double *host, size_t n, const double **out)
{
    double *d;
    int rc = dev_alloc(s, n, &d, false);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(d, host, n * sizeof(double), cudaMemcpyHostToDevice, s->ctx->stream));
    *out = d;
    return LUDVM_OK;
}

// Upper bound of the wake size reached within steps [first, first+K)
static long wake_upper(const ludvm_sim *s, long first_step)
{
    long i = std::min<long>(first_step + s->K, s->p.nt - 1);
    return 2 * i + 2 + s->p.nfree;
}

static int bracket_of(long n)
{
    int b = 6;
    while ((1L << b) < n) b++;
    return b;
}

static int build_graph(ludvm_sim *s, int bracket, int ksteps, cudaGraphExec_t *out)
{
    ludvm_ctx *ctx = s->ctx;
    const SimDev &D = s->d;
    const int nw = (int)std::min<long>(1L << bracket, 2L * D.nv + D.nfree + 2);  // wake-size upper bound
    const int sm = ctx->sm_count;
    // worst-case warp-task counts over every wake size n <= nw (split depth is monotone in n, capped by the target)
    const int dcap = std::min(pw_max_depth(nw), SIM_DMAX);
    const long quadsP = (D.P + 3) / 4;
    long t1 = std::min<long>(quadsP << dcap, std::max<long>(quadsP, 2L * D.target_warps));
    if (D.mode != LUDVM_EXACT_F64) t1 = quadsP * sim_chunks(nw, D.P, D.target_warps);
    int g1 = (int)std::max<long>(1, std::min<long>((t1 + 7) / 8, (long)sm * 8));
    const long quadsA = (D.P + nw + 3) / 4, quadsW = (nw + 3) / 4;
    long t3 = std::min<long>(quadsA << dcap, std::max<long>(quadsA, 2L * D.target_warps)) + quadsW;
    if (D.mode != LUDVM_EXACT_F64) t3 = std::max<long>(quadsA, (long)D.target_warps + quadsA) + quadsW;
    int g3 = (int)std::max<long>(1, std::min<long>((t3 + 7) / 8, (long)sm * 16));
    int g4 = 1 + (int)std::max<long>(1, std::min<long>((nw + 255) / 256, (long)sm * 8));

    cudaGraph_t graph;
    if (!s->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
    cudaStream_t cs = s->cap_stream;
    CUDA_TRY(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < ksteps; k++) {
        k_wake_on_foil<<<g1, 256, 0, cs>>>(D, k);
        k_solve<<<1, SOLVE_THREADS, s->solve_smem, cs>>>(D, k);
        k_conv_partials<<<g3, 256, 0, cs>>>(D, k);
        k_finish<<<g4, 256, s->finish_smem, cs>>>(D, k);
    }
    k_advance<<<1, 1, 0, cs>>>(D, ksteps);
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    if (e != cudaSuccess) return set_error(LUDVM_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return set_error(LUDVM_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    return LUDVM_OK;
}

}  // namespace ludvm

LUDVM_API int ludvm_sim_create(ludvm_ctx *ctx, const ludvm_sim_params *p, const ludvm_sim_tables *t, ludvm_sim **out)
{
    ARG_CHECK(ctx && p && t && out);
    *out = nullptr;
    ARG_CHECK(p->nt >= 2 && p->nt < (1 << 24) && p->P >= 2 && p->P <= 4096 && p->Nc >= 4 && p->Nc <= 512);
    ARG_CHECK(p->nfree >= 1 && p->nfree < (1 << 28));
    ARG_CHECK(p->mode == LUDVM_EXACT_F64 || p->mode == LUDVM_FAST_F64);
    if (p->method != LUDVM_METHOD_FAURE)
        return set_error(LUDVM_E_UNSUPPORTED, "only method='Faure' runs on the device in this build");
    ARG_CHECK(t->cos_a && t->sin_a && t->alpha_dot && t->h_dot && t->gp && t->le && t->te && t->detadx_p && t->eta_p &&
              t->x_p && t->theta_p && t->dtheta && t->cos_tp && t->sin_tp && t->cosn && t->sinn && t->free_g &&
              t->free_xz);
    DeviceGuard g(ctx->device);
    ludvm_sim *s = new ludvm_sim();
    s->ctx = ctx;
    s->p = *p;
    s->K = p->steps_per_graph > 0 ? p->steps_per_graph : 50;
    SimDev &D = s->d;
    const size_t nt = p->nt, P = p->P, Nc = p->Nc, nf = p->nfree, nv = nt - 1;
    D.nt = (int)nt; D.P = (int)P; D.Nc = (int)Nc; D.nfree = (int)nf; D.nv = (int)nv;
    D.method = p->method; D.mode = p->mode; D.store_history = p->store_history;
    D.target_warps = ctx->sm_count * 48;
    D.dt = p->dt; D.Uinf = p->Uinf; D.chord = p->chord; D.rho = p->rho; D.piv = p->piv; D.vc4 = p->vc4;
    D.ic = p->ic; D.sum_free = p->sum_free; D.maxerror = p->maxerror; D.epsilon = p->epsilon; D.maxiter = (int)p->maxiter;
    int rc = LUDVM_OK;
#define TRY(x) do { if ((rc = (x)) != LUDVM_OK) { ludvm_sim_destroy(s); return rc; } } while (0)
    TRY(upload(s, t->cos_a, nt, &D.cos_a));
    TRY(upload(s, t->sin_a, nt, &D.sin_a));
    TRY(upload(s, t->alpha_dot, nt, &D.alpha_dot));
    TRY(upload(s, t->h_dot, nt, &D.h_dot));
    TRY(upload(s, t->gp, nt * 2 * P, &D.gp));
    TRY(upload(s, t->le, nt * 2, &D.le));
    TRY(upload(s, t->te, nt * 2, &D.te));
    TRY(upload(s, t->detadx_p, P, &D.detadx_p));
    TRY(upload(s, t->eta_p, P, &D.eta_p));
    TRY(upload(s, t->x_p, P, &D.x_p));
    TRY(upload(s, t->theta_p, P, &D.theta_p));
    TRY(upload(s, t->dtheta, P, &D.dtheta));
    TRY(upload(s, t->cos_tp, P, &D.cos_tp));
    TRY(upload(s, t->sin_tp, P, &D.sin_tp));
    TRY(upload(s, t->cosn, Nc * P, &D.cosn));
    TRY(upload(s, t->sinn, Nc * P, &D.sinn));
    const size_t nstate = 2 * nv + nf;
    TRY(dev_alloc(s, nstate, &D.wx));
    TRY(dev_alloc(s, nstate, &D.wz));
    TRY(dev_alloc(s, nstate, &D.wg));
    CUDA_TRY(cudaMemcpyAsync(D.wg + 2 * nv, t->free_g, nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(D.wx + 2 * nv, t->free_xz, nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(D.wz + 2 * nv, t->free_xz + nf, nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRY(dev_alloc(s, nv, &D.g_bound));
    TRY(dev_alloc(s, nv * P, &D.g_airfoil));
    TRY(dev_alloc(s, nv * P, &D.gamma_airfoil));
    TRY(dev_alloc(s, nv * P, &D.Gamma_airfoil));
    TRY(dev_alloc(s, nt * 2 * Nc, &D.fourier));
    TRY(dev_alloc(s, nt, &D.lesp));
    TRY(dev_alloc(s, nt, &D.lesp_prev));
    TRY(dev_alloc(s, nt, &D.lev_shed));
    TRY(dev_alloc(s, nt, &D.Fn));
    TRY(dev_alloc(s, nt, &D.Fs));
    TRY(dev_alloc(s, nt, &D.L));
    TRY(dev_alloc(s, nt, &D.D));
    TRY(dev_alloc(s, nt, &D.T));
    TRY(dev_alloc(s, nt, &D.M));
    k_fill<<<ceil_div((long)nt, 256), 256, 0, ctx->stream>>>(D.lev_shed, nt, -1.0);  // LUDVM.py:654
    ctx->launches++;
    double f0[2] = {p->a0_init, p->a1_init};
    CUDA_TRY(cudaMemcpyAsync(D.fourier, f0, sizeof(f0), cudaMemcpyHostToDevice, ctx->stream));  // LUDVM.py:647
    if (p->store_history) {
        TRY(dev_alloc(s, nt * 2 * nv, &D.path_tev));
        TRY(dev_alloc(s, nt * 2 * nv, &D.path_lev));
        TRY(dev_alloc(s, nt * 2 * nf, &D.path_free));
        // row 0 of path['FREE'] holds the initial positions (LUDVM.py:618)
        CUDA_TRY(cudaMemcpyAsync(D.path_free, t->free_xz, 2 * nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    double *tmp;
    TRY(dev_alloc(s, 8, &tmp));
    D.counters = (long long *)tmp;
    TRY(dev_alloc(s, (nt + 2) / 2 + 2, &tmp));
    D.ilev_arr = (int *)tmp;
    TRY(dev_alloc(s, 1, &D.lespcrit_cur));
    CUDA_TRY(cudaMemcpyAsync(D.lespcrit_cur, &p->lespcrit, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // scratch: phase-C partials and loads+convection partials
    TRY(dev_alloc(s, ((size_t)1 << SIM_DMAX) * P, &D.pa_u, false));
    TRY(dev_alloc(s, ((size_t)1 << SIM_DMAX) * P, &D.pa_w, false));
    const size_t rows_max = P + nstate + 2;
    // rows * 2^d <= 4*quads * 2*target/quads = 8*target when split; rows when not
    size_t pb = std::max(rows_max, (size_t)16 * D.target_warps + rows_max);
    TRY(dev_alloc(s, pb, &D.pb_u, false));
    TRY(dev_alloc(s, pb, &D.pb_w, false));
    TRY(dev_alloc(s, nstate + 8, &D.foil_u, false));
    TRY(dev_alloc(s, nstate + 8, &D.foil_w, false));
#undef TRY
    s->solve_smem = (7 * P + Nc + 32 + SUM_NODES_MAX) * sizeof(double);
    s->finish_smem = (2 * P + 8) * sizeof(double);
    if (s->solve_smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->solve_smem);
        if (e != cudaSuccess) {
            ludvm_sim_destroy(s);
            return set_error(LUDVM_E_CUDA, "k_solve needs %zu bytes of shared memory: %s", s->solve_smem,
                             cudaGetErrorString(e));
        }
    }
    if (s->finish_smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(k_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->finish_smem));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the host tables may be freed by the caller after return
    *out = s;
    return LUDVM_OK;
}

LUDVM_API int ludvm_sim_run(ludvm_sim *s, long nsteps)
{
    ARG_CHECK(s != nullptr && nsteps >= 0);
    DeviceGuard g(s->ctx->device);
    long total = s->p.nt - 1;
    long todo = std::min(nsteps, total - s->steps_enqueued);
    // graphs of K unrolled steps (a shorter one for a tail), cached per (wake-size bracket, length)
    while (todo > 0) {
        int k = (int)std::min<long>(s->K, todo);
        int b = bracket_of(wake_upper(s, s->steps_enqueued));
        int key = b * 100000 + k;
        auto it = s->graphs.find(key);
        if (it == s->graphs.end()) {
            cudaGraphExec_t ge;
            int rc = build_graph(s, b, k, &ge);
            if (rc) return rc;
            it = s->graphs.emplace(key, ge).first;
        }
        CUDA_TRY(cudaGraphLaunch(it->second, s->ctx->stream));
        s->ctx->launches += 4L * k + 1;
        s->steps_enqueued += k;
        todo -= k;
    }
    return LUDVM_OK;
}

LUDVM_API int ludvm_sim_steps_done(ludvm_sim *s, long *out)
{
    ARG_CHECK(s && out);
    DeviceGuard g(s->ctx->device);
    long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, s->d.counters, sizeof(v), cudaMemcpyDeviceToHost, s->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    *out = (long)v;
    return LUDVM_OK;
}

static int field_info(ludvm_sim *s, int field, const void **ptr, size_t *bytes)
{
    const SimDev &D = s->d;
    const size_t nt = D.nt, nv = D.nv, P = D.P, Nc = D.Nc, nf = D.nfree, d8 = sizeof(double);
    switch (field) {
    case LUDVM_F_PATH_TEV: *ptr = D.path_tev; *bytes = nt * 2 * nv * d8; break;
    case LUDVM_F_PATH_LEV: *ptr = D.path_lev; *bytes = nt * 2 * nv * d8; break;
    case LUDVM_F_PATH_FREE: *ptr = D.path_free; *bytes = nt * 2 * nf * d8; break;
    case LUDVM_F_G_TEV: *ptr = D.wg; *bytes = nv * d8; break;
    case LUDVM_F_G_LEV: *ptr = D.wg + nv; *bytes = nv * d8; break;
    case LUDVM_F_G_BOUND: *ptr = D.g_bound; *bytes = nv * d8; break;
    case LUDVM_F_G_AIRFOIL: *ptr = D.g_airfoil; *bytes = nv * P * d8; break;
    case LUDVM_F_GAMMA_AIRFOIL: *ptr = D.gamma_airfoil; *bytes = nv * P * d8; break;
    case LUDVM_F_GAMMA_INT_AIRFOIL: *ptr = D.Gamma_airfoil; *bytes = nv * P * d8; break;
    case LUDVM_F_FOURIER: *ptr = D.fourier; *bytes = nt * 2 * Nc * d8; break;
    case LUDVM_F_LESP: *ptr = D.lesp; *bytes = nt * d8; break;
    case LUDVM_F_LESP_PREV: *ptr = D.lesp_prev; *bytes = nt * d8; break;
    case LUDVM_F_LEV_SHED: *ptr = D.lev_shed; *bytes = nt * d8; break;
    case LUDVM_F_FN: *ptr = D.Fn; *bytes = nt * d8; break;
    case LUDVM_F_FS: *ptr = D.Fs; *bytes = nt * d8; break;
    case LUDVM_F_L: *ptr = D.L; *bytes = nt * d8; break;
    case LUDVM_F_D: *ptr = D.D; *bytes = nt * d8; break;
    case LUDVM_F_T: *ptr = D.T; *bytes = nt * d8; break;
    case LUDVM_F_M: *ptr = D.M; *bytes = nt * d8; break;
    case LUDVM_F_COUNTERS: *ptr = D.counters; *bytes = 4 * sizeof(long long); break;
    case LUDVM_F_CUR_TEV: case LUDVM_F_CUR_LEV: case LUDVM_F_CUR_FREE: *ptr = nullptr; *bytes = 2 * (field == LUDVM_F_CUR_FREE ? nf : nv) * d8; break;
    default: return set_error(LUDVM_E_ARG, "unknown field %d", field);
    }
    return LUDVM_OK;
}

LUDVM_API int ludvm_sim_field_bytes(ludvm_sim *s, int field, size_t *out)
{
    ARG_CHECK(s && out);
    const void *p;
    return field_info(s, field, &p, out);
}

LUDVM_API int ludvm_sim_fetch(ludvm_sim *s, int field, void *dst, size_t bytes)
{
    ARG_CHECK(s && dst);
    DeviceGuard g(s->ctx->device);
    const void *p;
    size_t want;
    int rc = field_info(s, field, &p, &want);
    if (rc) return rc;
    if (bytes != want) return set_error(LUDVM_E_ARG, "field %d holds %zu bytes, caller passed %zu", field, want, bytes);
    cudaStream_t st = s->ctx->stream;
    if (field == LUDVM_F_CUR_TEV || field == LUDVM_F_CUR_LEV || field == LUDVM_F_CUR_FREE) {
        const SimDev &D = s->d;
        size_t off = field == LUDVM_F_CUR_TEV ? 0 : (field == LUDVM_F_CUR_LEV ? D.nv : 2 * (size_t)D.nv);
        size_t n = field == LUDVM_F_CUR_FREE ? D.nfree : D.nv;
        CUDA_TRY(cudaMemcpyAsync(dst, D.wx + off, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync((double *)dst + n, D.wz + off, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else {
        if (!p) return set_error(LUDVM_E_STATE, "field %d was not kept (store_history = 0)", field);
        CUDA_TRY(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return LUDVM_OK;
}

LUDVM_API int ludvm_sim_destroy(ludvm_sim *s)
{
    if (!s) return LUDVM_OK;
    DeviceGuard g(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    for (auto &kv : s->graphs) cudaGraphExecDestroy(kv.second);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    for (void *p : s->allocs) cudaFree(p);
    delete s;
    return LUDVM_OK;
}
