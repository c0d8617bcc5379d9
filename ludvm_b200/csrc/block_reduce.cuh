// block_reduce.cuh -- block-wide reductions in numpy's summation order for the once-per-step scalar phases of the
// time step (np.sum / np.trapz over the panels, folds of partial row sums).  These run on ONE CTA once per step, so
// what matters is latency: operands are first staged into shared memory by all threads (coalesced, many loads in
// flight), then 8-lane groups (or single threads for short rows) add them in the prescribed order.  Every routine
// is out of line so that each kernel carries one small copy (the solve kernel's code is fetched cold every step).
#pragma once
#include "common.cuh"

namespace ludvm {

#ifdef LUDVM_TRACE   // clock64 phase traces of the solve CTA (scripts/coop_trace.py); off in the shipped library
__device__ long long g_trace[64];
#define TRACE(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_trace[k] = clock64(); } while (0)
#else
#define TRACE(k) do { } while (0)
#endif

// np.sum(a[off:off+n]) as numpy's tree, by the calling 8-lane group.
__device__ __noinline__ double sum_group(const double *a, int off, int n)
{
    auto f = [a](int j) { return a[j]; };
    return pw_group(f, off, n, (int)(threadIdx.x & 7));
}
// The same when all four groups of every calling warp sum rows of the same length n (whole-warp shuffles).
__device__ __noinline__ double sum_group_uniform(const double *a, int n)
{
    auto f = [a](int j) { return a[j]; };
    return pw_group<true>(f, 0, n, (int)(threadIdx.x & 7));
}

// Fold of nn partials of one row, staged contiguously in shared memory, by the calling 8-lane group.
//   exact: the nn = 2^d node partials are the leaves of a perfect binary tree in index order; each lane folds a
//          contiguous subtree with the recursion's stack, an xor-butterfly closes the top levels, and numpy's
//          additive identity finishes (np.sum = 0.0 + pairwise).
//   fast:  nn chunk partials, any order.
__device__ __noinline__ double fold_group(const double *v, int nn, bool exact)
{
    const unsigned gm = group8_mask();
    const int lane8 = threadIdx.x & 7;
    if (exact) {
        const int per = max(1, nn >> 3), first = lane8 * per;
        double r = 0.0;
        if (first < nn) {
            double st[PW_MAX_STACK];
            int sp = 0;
            for (int i = 0; i < per; i++) {
                double x = v[first + i];
                for (int k = i; k & 1; k >>= 1) x = __dadd_rn(st[--sp], x);
                st[sp++] = x;
            }
            r = st[0];
        }
        for (int s = 1; s < 8 && s < nn; s <<= 1) r = __dadd_rn(r, __shfl_xor_sync(gm, r, s));
        return __dadd_rn(0.0, r);
    }
    double s = 0.0;
    for (int c = lane8; c < nn; c += 8) s += v[c];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) s += __shfl_xor_sync(gm, s, o);
    return s;
}

// Fold of NN = 2^k staged partials by one thread as the perfect binary tree in index order -- the tree numpy's
// recursion builds above the node level -- written as straight-line code (independent subtrees overlap).
template <int NN>
__device__ __forceinline__ double fold_tree(const double *v)
{
    if constexpr (NN == 1) return v[0];
    else return __dadd_rn(fold_tree<NN / 2>(v), fold_tree<NN / 2>(v + NN / 2));
}

__device__ __forceinline__ double fold_thread(const double *v, int nn, bool exact)
{
    if (exact) {  // nn is a power of two <= 64
        double r;
        switch (nn) {
        case 1: r = v[0]; break;
        case 2: r = fold_tree<2>(v); break;
        case 4: r = fold_tree<4>(v); break;
        case 8: r = fold_tree<8>(v); break;
        case 16: r = fold_tree<16>(v); break;
        case 32: r = fold_tree<32>(v); break;
        default: r = fold_tree<64>(v); break;
        }
        return __dadd_rn(0.0, r);
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = 0;
    for (; c + 3 < nn; c += 4) {
        s0 += v[c];
        s1 += v[c + 1];
        s2 += v[c + 2];
        s3 += v[c + 3];
    }
    for (; c < nn; c++) s0 += v[c];
    return (s0 + s1) + (s2 + s3);
}

// Block-wide fold of the partial sums of rows [0, nr) of two components (u, w): the partials [fold][nrows] are first
// staged into shared memory by all threads (one warp per partial index, lanes along the rows: coalesced, no index
// arithmetic), then each (component, row) is folded in the prescribed order -- by one thread when the row has at
// most 64 partials, by an 8-lane group otherwise.  `nfold` is the tree depth (exact) or the chunk count (fast).
// Ends with a block barrier.
__device__ __noinline__ void block_fold(const double *__restrict__ pu, const double *__restrict__ pw, int nrows, int nr,
                                        int nfold, bool exact, double *__restrict__ stage, int cap,
                                        double *__restrict__ out_u, double *__restrict__ out_w)
{
    const int tid = threadIdx.x, nth = blockDim.x, lane8 = tid & 7, grp = tid >> 3, ngrp = nth >> 3;
    const int nn = exact ? 1 << nfold : nfold, ld = nn | 1;  // odd row pitch: no bank conflicts while staging
    const int per = max(1, cap / ld);
    for (int t0 = 0; t0 < 2 * nr; t0 += per) {
        const int tend = min(t0 + per, 2 * nr), nb = tend - t0;   // see the note in block_trapz
        TRACE(30);
        __syncthreads();
        TRACE(31);
        // every thread first issues all of its loads (up to 8 independent ones), then stores: one round trip to L2 for
        // the whole batch (element e = f * nb + tl: consecutive threads read consecutive rows of one partial index)
        const int total = nn * nb;
        for (int e0 = 0; e0 < total; e0 += 8 * nth) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int e = e0 + k * nth + tid;
                if (e < total) {
                    const int f = e / nb, tl = e - f * nb, t = t0 + tl, c = t >= nr, r = t - c * nr;
                    v[k] = (c ? pw : pu)[(size_t)f * nrows + r];
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int e = e0 + k * nth + tid;
                if (e < total) {
                    const int f = e / nb, tl = e - f * nb;
                    stage[tl * ld + f] = v[k];
                }
            }
        }
        TRACE(32);
        __syncthreads();
        TRACE(33);
        if (nn <= 64) {
            for (int tl = tid; tl < nb; tl += nth) {
                double v = fold_thread(stage + tl * ld, nn, exact);
                int t = t0 + tl, c = t >= nr, r = t - c * nr;
                (c ? out_w : out_u)[r] = v;
            }
        } else {
            for (int tl = grp; tl < nb; tl += ngrp) {
                double v = fold_group(stage + tl * ld, nn, exact);
                if (lane8 == 0) {
                    int t = t0 + tl, c = t >= nr, r = t - c * nr;
                    (c ? out_w : out_u)[r] = v;
                }
            }
        }
        TRACE(34);
    }
    __syncthreads();
    TRACE(35);
}

// Block-wide np.trapz of nq integrands at once:
//   out[q] = np.trapz(a_q * b_q, x),  a_q = a0 + (q & amask) * astride,  b_q = b0 + (q >> bshift) * bstride,
// dx[j] = x[j+1] - x[j] (the same subtraction, tabulated).  A plain np.trapz(a, x) passes a table of ones for b
// (a * 1.0 is exact).  All threads evaluate the terms d*(y[1:]+y[:-1])/2.0 into shared memory (coalesced operand
// loads, everything in flight at once); then one 8-lane group per integrand adds them in numpy's pairwise order.
// The solve phase runs once per step on one CTA, so what matters is its latency: few dependent round trips and a
// small code footprint (one out-of-line copy serves every integral of the step).  Ends with a block barrier.
__device__ __noinline__ void block_trapz(const double *a0, int amask, int astride, const double *b0, int bshift,
                                         int bstride, const double *dx, int P, int nq, double *__restrict__ stage,
                                         int cap, double *out)
{
    const int n = P - 1, tid = threadIdx.x, nth = blockDim.x, lane8 = tid & 7, grp = tid >> 3, ngrp = nth >> 3;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    const int rows = max(1, cap / P);
    for (int q0 = 0; q0 < nq; q0 += rows) {
        // batch [q0, qend): written as min(q0 + rows, nq) - q0, NOT min(rows, nq - q0) -- when ptxas (12.9) clones this
        // function for a call site with a literal nq it folds `nq - q0` into VIADDMNMX(q0 + (-nq), rows), i.e. the
        // wrong sign, and the batch comes out empty (observed in k_finish: both load integrals stayed 0)
        const int qend = min(q0 + rows, nq), nb = qend - q0;
        TRACE(36);
        __syncthreads();
        TRACE(37);
        for (int q = warp; q < nb; q += nwarps) {   // one warp per integrand, lanes along the panels
            const int qq = q0 + q;
            const double *a = a0 + (qq & amask) * astride, *b = b0 + (size_t)(qq >> bshift) * bstride;
            double *srow = stage + q * P;
#pragma unroll 4
            for (int j = lane; j < n; j += 32) srow[j] = dx[j] * (a[j + 1] * b[j + 1] + a[j] * b[j]) / 2.0;
        }
        TRACE(38);
        __syncthreads();
        TRACE(39);
        // every group of every warp takes the same trips with the same length (a group past the end re-sums the last
        // integrand and drops the result), so the shuffles may name the whole warp: ~70 cycles less per shuffle than
        // the per-group masks, ~10 shuffles per sum
        for (int qb = 0; qb < nb; qb += ngrp) {
            const int q = qb + grp;
            double v = 0.0 + sum_group_uniform(stage + min(q, nb - 1) * P, n);
            if (lane8 == 0 && q < nb) out[q0 + q] = v;
        }
    }
    __syncthreads();
}

// np.sum(a[:n]) by the whole block: the tree nodes at depth d are summed by 8-lane groups, the top of the tree is
// folded level by level in shared memory.  Returns the sum on every thread.  Starts with a block barrier (so stores
// to `a` and to shared memory made before the call are visible inside and after it).
__device__ __noinline__ double block_np_sum(const double *a, int n, double *s_nodes, int max_nodes)
{
    __syncthreads();
    const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3, ngrp = blockDim.x >> 3;
    int d = pw_max_depth(n);
    while ((1 << d) > max_nodes) d--;
    const int nn = 1 << d;
    for (int b = grp; b < nn; b += ngrp) {
        int off, len;
        pw_node(n, d, b, off, len);
        double v = sum_group(a, off, len);
        if (lane8 == 0) s_nodes[b] = v;
    }
    __syncthreads();
    for (int stride = 1; stride < nn; stride <<= 1) {  // left + right, in place at the left child's slot
        for (int i = threadIdx.x * 2 * stride; i < nn; i += blockDim.x * 2 * stride)
            s_nodes[i] = __dadd_rn(s_nodes[i], s_nodes[i + stride]);
        __syncthreads();
    }
    return 0.0 + s_nodes[0];
}

}  // namespace ludvm
