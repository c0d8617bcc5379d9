// biot_savart.cu -- generic all-pairs entry points: ludvm_induced_velocity (LUDVM.py:549-570),
// ludvm_selfconv_step (LUDVM.py:1095-1127 without the aerofoil), ludvm_flowfield_* (LUDVM.py:1186-1298).
#include "biot_savart.cuh"

namespace ludvm {

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// Range scan behind the flag-free exact kernels (common.cuh, coord_in_safe_window): *bad |= 1 if any coordinate of the
// four arrays lies outside the window in which the branch-free division / square root need no range test.
__global__ void __launch_bounds__(256) k_range_scan(const double *a, int na, const double *b, int nb, const double *c,
                                                    int nc, const double *d, int nd, int *bad)
{
    const long tot = (long)na + nb + nc + nd;
    bool ok = true;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long)gridDim.x * blockDim.x) {
        double v;
        if (i < na) v = a[i];
        else if (i < (long)na + nb) v = b[i - na];
        else if (i < (long)na + nb + nc) v = c[i - na - nb];
        else v = d[i - na - nb - nc];
        ok = ok && coord_in_safe_window(v);
    }
    if (!ok) atomicOr(bad, 1);
}

// `bad`: nullptr = always evaluate the range words; otherwise the scan's verdict selects the instantiation (uniform).
template <int R, class Tgt>
__global__ void __launch_bounds__(256) k_exact_rows(SrcView S, Tgt T, int nrows, int d, double *pu, double *pw_,
                                                    const int *bad)
{
    int lane = threadIdx.x & 31;
    long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    long ntasks = (long)((nrows + 4 * R - 1) / (4 * R)) << d;
    if (bad && *bad == 0)
        for (long t = gw; t < ntasks; t += nwarps) exact_rows_warp_task<R, false>(S, T, nrows, d, t, lane, pu, pw_);
    else
        for (long t = gw; t < ntasks; t += nwarps) exact_rows_warp_task<R, true>(S, T, nrows, d, t, lane, pu, pw_);
}

template <class Tgt>
__global__ void __launch_bounds__(ET_THREADS, 3) k_exact_tiled(SrcView S, Tgt T, int nrows, int d, double *pu, double *pw_,
                                                               const int *bad)
{
    __shared__ __align__(16) double2 sxz[ET_TILE], sgv[ET_TILE];
    if (bad && *bad == 0) exact_tiled_block<false>(S, T, nrows, blockIdx.x, d, blockIdx.y, pu, pw_, sxz, sgv);
    else exact_tiled_block<true>(S, T, nrows, blockIdx.x, d, blockIdx.y, pu, pw_, sxz, sgv);
}

template <class Tgt>
__global__ void __launch_bounds__(256) k_fast_rows(SrcView S, Tgt T, int nrows, int nchunks, double *pu, double *pw_)
{
    int lane = threadIdx.x & 31;
    long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    long ntasks = (long)((nrows + 3) >> 2) * nchunks;
    for (long t = gw; t < ntasks; t += nwarps) fast_rows_warp_task(S, T, nrows, nchunks, t, lane, pu, pw_);
}

template <int R, class Tgt>
__global__ void __launch_bounds__(FT_THREADS, 2)
k_fast_tiled(SrcView S, Tgt T, int nrows, int chunk_len, double *pu, double *pw_)
{
    __shared__ double sx[FT_TILE], sz[FT_TILE], sg[FT_TILE], sv[FT_TILE];
    int c0 = blockIdx.y * chunk_len, c1 = min(S.n, c0 + chunk_len);
    size_t po = (size_t)blockIdx.y * nrows;
    fast_tiled_block<R>(S, T, nrows, blockIdx.x, c0, c1, pu + po, pw_ + po, sx, sz, sg, sv);
}

template <int R, class Tgt>
__global__ void __launch_bounds__(FT_THREADS, 2)
k_fast_tiled_tma(SrcView S, Tgt T, int nrows, int chunk_len, double *pu, double *pw_)
{
    __shared__ __align__(128) TmaTiles tiles;
    int c0 = blockIdx.y * chunk_len, c1 = min(S.n, c0 + chunk_len);
    size_t po = (size_t)blockIdx.y * nrows;
    fast_tiled_block_tma<R>(S, T, nrows, blockIdx.x, c0, c1, pu + po, pw_ + po, tiles);
}

template <int R, class Tgt>
__global__ void __launch_bounds__(FT_THREADS, 2)
k_fast32_tiled(SrcView S, Tgt T, int nrows, int chunk_len, double *pu, double *pw_)
{
    __shared__ float4 ssrc[FT_TILE];
    int c0 = blockIdx.y * chunk_len, c1 = min(S.n, c0 + chunk_len);
    size_t po = (size_t)blockIdx.y * nrows;
    fast32_tiled_block<R>(S, T, nrows, blockIdx.x, c0, c1, pu + po, pw_ + po, ssrc);
}

// fp32 with packed fp32x2 arithmetic: 8 rows (4 float2 pairs) per thread; scalar core radius only.
template <class Tgt>
__global__ void __launch_bounds__(FT_THREADS, 2)
k_fast32x2_tiled(SrcView S, Tgt T, int nrows, int chunk_len, double *pu, double *pw_)
{
    __shared__ Src32x2 ssrc[FT_TILE];
    int c0 = blockIdx.y * chunk_len, c1 = min(S.n, c0 + chunk_len);
    size_t po = (size_t)blockIdx.y * nrows;
    fast32x2_tiled_block<4>(S, T, nrows, blockIdx.x, c0, c1, pu + po, pw_ + po, ssrc);
}

// Whole row sums in one launch: warp-private bulk-copy pipelines, chunk partials folded through (distributed) shared
// memory, Euler update and peer stores in the epilogue (biot_savart.cuh, "fast fused").
template <int R, int UNROLL, int WARPS, int CL, int V, class Tgt>
__global__ void __launch_bounds__(32 * WARPS, 16 / WARPS)
k_fast_fused(SrcView S, Tgt T, int nrows, int chunk_len, int nchunks, FusedOut O)
{
    extern __shared__ __align__(128) unsigned char fw_raw[];
    fast_fused_block<R, UNROLL, WARPS, CL, V>(S, T, nrows, chunk_len, nchunks, O, *reinterpret_cast<FwSmem<WARPS> *>(fw_raw));
}

// Fold partials.  exact: nfold = tree depth d; fast: nfold = number of chunks.  Optional second partial set
// (the reference's `u_wake + u_foil`, LUDVM.py:1108, :1219) and optional forward-Euler update (LUDVM.py:1108-1127).
struct CombineArgs {
    const double *pu, *pw;   // [fold][nrows]
    const double *qu, *qw;   // second set or nullptr
    int nfold, qfold;
    int exact;
    int nrows;
    double *u, *w;           // nullable
    const double *x, *z;     // Euler: inputs (already offset to the shard's first row)
    double *xo, *zo;         // Euler: outputs (nullable)
    double dt;
    // fused all-gather: the updated rows are also stored into every peer's buffers (NVLink peer stores),
    // already offset to the shard's first row
    int npeers;
    double *xo_peer[LUDVM_MAX_PEERS], *zo_peer[LUDVM_MAX_PEERS];
};

__global__ void __launch_bounds__(256) k_combine(CombineArgs a)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.nrows) return;
    double u, w;
    if (a.exact) {
        u = exact_combine_row(a.pu, a.nrows, row, a.nfold);
        w = exact_combine_row(a.pw, a.nrows, row, a.nfold);
        if (a.qu) {
            u = __dadd_rn(u, exact_combine_row(a.qu, a.nrows, row, a.qfold));
            w = __dadd_rn(w, exact_combine_row(a.qw, a.nrows, row, a.qfold));
        }
    } else {
        u = fast_combine_row(a.pu, a.nrows, row, a.nfold);
        w = fast_combine_row(a.pw, a.nrows, row, a.nfold);
        if (a.qu) {
            u += fast_combine_row(a.qu, a.nrows, row, a.qfold);
            w += fast_combine_row(a.qw, a.nrows, row, a.qfold);
        }
    }
    if (a.u) {
        a.u[row] = u;
        a.w[row] = w;
    }
    if (a.xo || a.npeers) {
        double xn = __dadd_rn(a.x[row], __dmul_rn(a.dt, u));
        double zn = __dadd_rn(a.z[row], __dmul_rn(a.dt, w));
        if (a.xo) {
            a.xo[row] = xn;
            a.zo[row] = zn;
        }
        for (int p = 0; p < a.npeers; p++) {  // st.global on mapped peer pointers: the all-gather, fused
            a.xo_peer[p][row] = xn;
            a.zo_peer[p][row] = zn;
        }
    }
}

// Vorticity stencil, LUDVM.py:1222-1292: clamped neighbours reproduce the centred / one-sided variants.
__global__ void __launch_bounds__(256) k_vorticity(const double *x1, int nx, const double *z1, int nz, const double *u,
                                                   const double *w, int ns, double *ome)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long per = (long)nx * nz;
    if (idx >= per * ns) return;
    long s = idx / per, r = idx - s * per;
    int i = (int)(r / nz), j = (int)(r - (long)i * nz);
    int ip = min(i + 1, nx - 1), im = max(i - 1, 0), jp = min(j + 1, nz - 1), jm = max(j - 1, 0);
    const double *us = u + s * per, *ws = w + s * per;
    double dx = __dsub_rn(x1[ip], x1[im]);
    double dz = __dsub_rn(z1[jp], z1[jm]);
    double dw = __dsub_rn(ws[(long)ip * nz + j], ws[(long)im * nz + j]);
    double du = __dsub_rn(us[(long)i * nz + jp], us[(long)i * nz + jm]);
    ome[idx] = __dsub_rn(__ddiv_rn(dw, dx), __ddiv_rn(du, dz));
}

// ---------------------------------------------------------------------------------------------------
// host-side planning and launch
// ---------------------------------------------------------------------------------------------------
static int ilog2_ceil(long v)
{
    int l = 0;
    while ((1L << l) < v) l++;
    return l;
}

// Source chunking of the fast tiled kernels.  It depends on the number of SOURCES only, so a row's sum does not depend
// on how the target rows are sharded over GPUs (G-rank results are bitwise equal to 1-rank results).
// 16 chunks from 131072 sources (measured at 2^20 sources: 866 ms against 875 ms with 8 for 2^20 rows, and 109.4 ms
// against 115.5 ms for the 131072 rows of one rank of eight, where 8 chunks force R = 2; 32 chunks: no further gain;
// profiles/r01e_chunks_probe.txt).  LUDVM_FAST_CHUNKS overrides (experiments; same value on every rank).
static void fast_chunking(int n, int *chunk_len, long *chunks)
{
    const char *ce = getenv("LUDVM_FAST_CHUNKS");
    const long cap = ce ? std::max(1L, atol(ce)) : (n >= 131072 ? 16L : 8L);
    long ch = std::max(1L, std::min(cap, (long)n / (FT_TILE * 2)));
    *chunk_len = (int)(((n + ch - 1) / ch + FT_TILE - 1) / FT_TILE * FT_TILE);
    *chunks = ((long)n + *chunk_len - 1) / *chunk_len;
}

static void set_plan(ludvm_ctx *ctx, int kernel, int R, int fold, int tma, int cluster, int variant = 0)
{
    ctx->plan[0] = kernel; ctx->plan[1] = R; ctx->plan[2] = fold; ctx->plan[3] = tma; ctx->plan[4] = cluster;
    ctx->plan[5] = variant; ctx->plan[6] = ctx->plan[7] = 0;
}

// Targets as (up to) two coordinate arrays for the range scan.
static void tgt_arrays(const TgtArray &T, long nrows, const double **a, int *na, const double **b, int *nb)
{
    *a = T.x; *na = (int)nrows; *b = T.z; *nb = (int)nrows;
}
static void tgt_arrays(const TgtGrid &T, long nrows, const double **a, int *na, const double **b, int *nb)
{
    *a = T.x1 + T.row0; *na = (int)((nrows + T.nz - 1) / T.nz); *b = T.z1; *nb = T.nz;
}

// Launch the range scan for an exact-mode evaluation; *flag = device flag for the kernels, or nullptr when the proof
// does not apply (per-source cores, vc^4 outside its window -- e.g. viscous=False --, LUDVM_EXACT_FLAGS=1).
template <class Tgt>
static int prove_exact_ranges(ludvm_ctx *ctx, const SrcView &S, const Tgt &T, long nrows, const int **flag)
{
    *flag = nullptr;
    if (S.vc4 != nullptr || !vc4_in_safe_window(S.vc4s) || S.n0 != S.n || getenv("LUDVM_EXACT_FLAGS")) return LUDVM_OK;
    void *p;
    int rc = scratch_reserve(ctx, 7, 256, &p);
    if (rc) return rc;
    int *bad = (int *)((char *)p + 128);
    CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
    const double *ta, *tb;
    int na, nb;
    tgt_arrays(T, nrows, &ta, &na, &tb, &nb);
    const long tot = 2L * S.n + na + nb;
    const int blocks = (int)std::min<long>((tot + 1023) / 1024, (long)ctx->sm_count * 8);
    k_range_scan<<<blocks, 256, 0, ctx->stream>>>(S.x, S.n, S.z, S.n, ta, na, tb, nb, bad);
    ctx->launches++;
    ctx->range_flag = bad;
    *flag = bad;
    return LUDVM_OK;
}

template <int R, int UNROLL, int WARPS, int CL, int V, class Tgt>
static int launch_fused_inst(ludvm_ctx *ctx, const SrcView &S, const Tgt &T, int nrows, int chunk_len, int nchunks,
                             const FusedOut &O)
{
    auto kern = k_fast_fused<R, UNROLL, WARPS, CL, V, Tgt>;
    static bool configured[16] = {};                       // per device; a function attribute is per device
    if (!configured[ctx->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwSmem<WARPS>)));
        configured[ctx->device & 15] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ceil_div(nrows, 32 * R), CL, 1);
    cfg.blockDim = dim3(32 * WARPS, 1, 1);
    cfg.dynamicSmemBytes = sizeof(FwSmem<WARPS>);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = CL;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, S, T, nrows, chunk_len, nchunks, O));
    ctx->launches++;
    set_plan(ctx, LUDVM_K_FAST_FUSED, R, nchunks, 1, CL, UNROLL);
    ctx->plan[6] = WARPS;
    ctx->plan[7] = V;
    return LUDVM_OK;
}

// The fused single-launch path: fast f64, one contiguous 16-byte aligned source segment with a scalar core, 8 or 16
// source chunks (>= 8192 sources) and enough rows for >= ~6 waves of 2 CTAs/SM.  *launched = false: not eligible, the
// caller takes the partial-sum path.
template <class Tgt>
static int try_launch_fused(ludvm_ctx *ctx, int mode, const SrcView &S, const Tgt &T, long nrows, const FusedOut &O,
                            bool *launched)
{
    *launched = false;
    if ((mode != LUDVM_FAST_F64 && mode != LUDVM_FAST12_F64) || getenv("LUDVM_NO_FUSED") || getenv("LUDVM_NO_TMA"))
        return LUDVM_OK;
    if (S.vc4 != nullptr || S.gstride != 1 || S.n0 != S.n ||
        (((uintptr_t)S.x | (uintptr_t)S.z | (uintptr_t)S.g) & 15) != 0)
        return LUDVM_OK;
    int chunk_len;
    long chunks;
    fast_chunking(S.n, &chunk_len, &chunks);
    if (chunks != 8 && chunks != 16) return LUDVM_OK;
    // 16 chunks: one 16-warp CTA per SM (LUDVM_FUSED_WARPS=16) or a cluster of two 8-warp CTAs (default)
    const char *we = getenv("LUDVM_FUSED_WARPS");
    const int W = (chunks == 16 && we && atoi(we) == 16) ? 16 : 8;
    const int CL = (int)chunks / W;
    const long want = (long)ctx->sm_count * 2 * 6;
    int R = 4;
    while (R >= 1 && (long)ceil_div(nrows, 32 * R) * (chunks / 8) < want) R >>= 1;
    if (R < 1) return LUDVM_OK;
    const char *ue = getenv("LUDVM_FUSED_UNROLL");
    const int U = ue ? atoi(ue) : 4;
    int rc;
#define FUSED_CASE(RR, UU, VV)                                                                                      \
    rc = W == 16 ? launch_fused_inst<RR, UU, 16, 1, VV>(ctx, S, T, (int)nrows, chunk_len, (int)chunks, O)           \
       : CL == 2 ? launch_fused_inst<RR, UU, 8, 2, VV>(ctx, S, T, (int)nrows, chunk_len, (int)chunks, O)            \
                 : launch_fused_inst<RR, UU, 8, 1, VV>(ctx, S, T, (int)nrows, chunk_len, (int)chunks, O)
    if (mode == LUDVM_FAST12_F64) {   // opt-in 12-slot pair arithmetic
        if (R == 4 && U == 4) FUSED_CASE(4, 4, 1);
        else if (R == 4) FUSED_CASE(4, 8, 1);
        else if (R == 2) FUSED_CASE(2, 4, 1);
        else FUSED_CASE(1, 8, 1);
    } else if (R == 4 && U == 1) FUSED_CASE(4, 1, 0);
    else if (R == 4 && U == 2) FUSED_CASE(4, 2, 0);
    else if (R == 4 && U == 8) FUSED_CASE(4, 8, 0);
    else if (R == 4) FUSED_CASE(4, 4, 0);
    else if (R == 2) FUSED_CASE(2, 4, 0);
    else FUSED_CASE(1, 8, 0);
#undef FUSED_CASE
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    *launched = true;
    return LUDVM_OK;
}

// Evaluate partial row sums for `nrows` targets against S; on return *nfold and the partial buffers (scratch
// slots slot/slot+1) describe what k_combine must fold.
template <class Tgt>
static int launch_partials(ludvm_ctx *ctx, int mode, const SrcView &S, const Tgt &T, long nrows, int slot,
                           double **pu, double **pw_, int *nfold)
{
    const int sm = ctx->sm_count;
    if (mode == LUDVM_FAST12_F64) mode = LUDVM_FAST_F64;   // outside the fused kernel: the (more accurate) 13-slot pair
    if (mode == LUDVM_EXACT_F64) {
        const int *bad;
        {
            int rc = prove_exact_ranges(ctx, S, T, nrows, &bad);
            if (rc) return rc;
        }
        if (nrows >= 4096 && (S.n >= 1024 || nrows >= (long)sm * ET_THREADS * 3)) {   // (few sources: only if the rows alone fill the GPU)
            // Many rows: one thread per row, sources staged through shared memory; the tree is cut at depth d so
            // that >= ~6 waves of 4 CTAs/SM are in flight.
            long rblocks = (nrows + ET_THREADS - 1) / ET_THREADS;
            int want = ilog2_ceil(std::max(1L, (long)sm * 24 / rblocks));
            int d = std::min(pw_max_depth(S.n), want);
            size_t bytes = sizeof(double) * (size_t)nrows * ((size_t)1 << d);
            void *a, *b;
            int rc;
            if ((rc = scratch_reserve(ctx, slot, bytes, &a))) return rc;
            if ((rc = scratch_reserve(ctx, slot + 1, bytes, &b))) return rc;
            k_exact_tiled<<<dim3((unsigned)rblocks, 1u << d), ET_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, d, (double *)a,
                                                                                            (double *)b, bad);
            ctx->launches++;
            set_plan(ctx, LUDVM_K_EXACT_TILED, 1, d, 0, 1, bad ? 1 : 0);
            *pu = (double *)a; *pw_ = (double *)b; *nfold = d;
            return LUDVM_OK;
        }
        // Few rows (or few rows and few sources): 8 lanes per row, tree nodes spread over warps.
        constexpr int R = 1;
        long nquads = (nrows + 4 * R - 1) / (4 * R);
        int want = ilog2_ceil(std::max(1L, (long)sm * 64 / std::max(1L, nquads)));
        int d = std::min(pw_max_depth(S.n), want);
        size_t bytes = sizeof(double) * (size_t)nrows * ((size_t)1 << d);
        void *a, *b;
        int rc;
        if ((rc = scratch_reserve(ctx, slot, bytes, &a))) return rc;
        if ((rc = scratch_reserve(ctx, slot + 1, bytes, &b))) return rc;
        long ntasks = nquads << d;
        int blocks = (int)std::min((ntasks + 7) / 8, (long)sm * 16);
        k_exact_rows<R><<<blocks, 256, 0, ctx->stream>>>(S, T, (int)nrows, d, (double *)a, (double *)b, bad);
        ctx->launches++;
        set_plan(ctx, LUDVM_K_EXACT_ROWS, R, d, 0, 1, bad ? 1 : 0);
        *pu = (double *)a; *pw_ = (double *)b; *nfold = d;
        return LUDVM_OK;
    }
    const bool f32 = (mode == LUDVM_FAST_F32);
    if (f32 || nrows >= 8192) {
        // source chunking from the number of sources only (fast_chunking); the rows-per-thread factor R adapts to the
        // number of rows to keep >= ~6 waves of 2 CTAs/SM in flight
        int chunk_len;
        long chunks;
        fast_chunking(S.n, &chunk_len, &chunks);
        // fp32: the packed fp32x2 kernel (8 rows per thread) when the core radius is a scalar and there are enough rows
        const bool f32x2 = f32 && S.vc4 == nullptr && (long)ceil_div(nrows, FT_THREADS * 8) * chunks >= (long)sm * 2 &&
                           !getenv("LUDVM_NO_F32X2");
        int R = f32x2 ? 8 : 4;
        if (!f32) {
            while (R > 1 && (long)ceil_div(nrows, FT_THREADS * R) * chunks < (long)sm * 2 * 6) R >>= 1;
        }
        int row_blocks = ceil_div(nrows, FT_THREADS * R);
        size_t bytes = sizeof(double) * (size_t)nrows * (size_t)chunks;
        void *a, *b;
        int rc;
        if ((rc = scratch_reserve(ctx, slot, bytes, &a))) return rc;
        if ((rc = scratch_reserve(ctx, slot + 1, bytes, &b))) return rc;
        dim3 grid(row_blocks, (unsigned)chunks);
        // one contiguous, 16-byte aligned source segment with a scalar core: TMA-staged tiles
        const bool tma = !f32 && S.vc4 == nullptr && S.gstride == 1 && S.n0 == S.n &&
                         (((uintptr_t)S.x | (uintptr_t)S.z | (uintptr_t)S.g) & 15) == 0 && !getenv("LUDVM_NO_TMA");
        if (f32x2) k_fast32x2_tiled<<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        else if (f32) k_fast32_tiled<4><<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        else if (tma && R == 4) k_fast_tiled_tma<4><<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        else if (tma && R == 2) k_fast_tiled_tma<2><<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        else if (tma) k_fast_tiled_tma<1><<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        else if (R == 4) k_fast_tiled<4><<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        else if (R == 2) k_fast_tiled<2><<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        else k_fast_tiled<1><<<grid, FT_THREADS, 0, ctx->stream>>>(S, T, (int)nrows, chunk_len, (double *)a, (double *)b);
        ctx->launches++;
        set_plan(ctx, f32x2 ? LUDVM_K_FAST32X2_TILED : f32 ? LUDVM_K_FAST32_TILED : tma ? LUDVM_K_FAST_TILED_TMA : LUDVM_K_FAST_TILED,
                 R, (int)chunks, tma ? 1 : 0, 1);
        *pu = (double *)a; *pw_ = (double *)b; *nfold = (int)chunks;
        return LUDVM_OK;
    }
    long nquads = (nrows + 3) / 4;
    long chunks = std::max(1L, std::min((long)sm * 64 / std::max(1L, nquads), ((long)S.n + 63) / 64));
    size_t bytes = sizeof(double) * (size_t)nrows * (size_t)chunks;
    void *a, *b;
    int rc;
    if ((rc = scratch_reserve(ctx, slot, bytes, &a))) return rc;
    if ((rc = scratch_reserve(ctx, slot + 1, bytes, &b))) return rc;
    long ntasks = nquads * chunks;
    int blocks = (int)std::min((ntasks + 7) / 8, (long)sm * 16);
    k_fast_rows<<<blocks, 256, 0, ctx->stream>>>(S, T, (int)nrows, (int)chunks, (double *)a, (double *)b);
    ctx->launches++;
    set_plan(ctx, LUDVM_K_FAST_ROWS, 1, (int)chunks, 0, 1);
    *pu = (double *)a; *pw_ = (double *)b; *nfold = (int)chunks;
    return LUDVM_OK;
}

static int launch_combine(ludvm_ctx *ctx, const CombineArgs &a)
{
    k_combine<<<ceil_div(a.nrows, 256), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return LUDVM_OK;
}

// Copy a host array to device scratch (slot) on the context's stream.
static int stage_in(ludvm_ctx *ctx, int slot, const double *host, size_t n, double **dev)
{
    void *p;
    int rc = scratch_reserve(ctx, slot, std::max<size_t>(n, 1) * sizeof(double), &p);
    if (rc) return rc;
    if (n) CUDA_TRY(cudaMemcpyAsync(p, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    *dev = (double *)p;
    return LUDVM_OK;
}

static int check_mode(int mode)
{
    if (mode != LUDVM_EXACT_F64 && mode != LUDVM_FAST_F64 && mode != LUDVM_FAST_F32 && mode != LUDVM_FAST12_F64)
        return set_error(LUDVM_E_ARG, "unknown arithmetic mode %d", mode);
    return LUDVM_OK;
}

}  // namespace ludvm

using namespace ludvm;

// Scratch slot map for the generic entry points: 0,1 partials A; 2,3 partials B; 4 inputs; 5 outputs; 6 grid axes.
LUDVM_API int ludvm_induced_velocity(ludvm_ctx *ctx, int mode, const double *gamma, long ngamma, const double *xw,
                                     const double *zw, const double *vc4_per_source, double vc4, long nw,
                                     const double *xp, const double *zp, long np_, double *u, double *w, int ptr_kind)
{
    ARG_CHECK(ctx != nullptr);
    int rc = check_mode(mode);
    if (rc) return rc;
    ARG_CHECK(nw >= 0 && np_ >= 0 && nw < (1L << 30) && np_ < (1L << 30));
    ARG_CHECK(ngamma == nw || ngamma == 1);
    ARG_CHECK(ptr_kind == LUDVM_PTR_HOST || ptr_kind == LUDVM_PTR_DEVICE);
    if (np_ == 0) return LUDVM_OK;
    ARG_CHECK(u && w && xp && zp);
    ARG_CHECK(nw == 0 || (gamma && xw && zw));
    DeviceGuard g(ctx->device);

    const double *dg = gamma, *dxw = xw, *dzw = zw, *dvc = vc4_per_source, *dxp = xp, *dzp = zp;
    double *du = u, *dw = w;
    if (ptr_kind == LUDVM_PTR_HOST) {
        // one staging buffer: [g | xw | zw | vc | xp | zp], outputs in slot 5
        size_t nsrc = (size_t)nw, ntg = (size_t)np_;
        size_t total = (size_t)ngamma + 2 * nsrc + (vc4_per_source ? nsrc : 0) + 2 * ntg;
        void *p;
        if ((rc = scratch_reserve(ctx, 4, (total + 8) * sizeof(double), &p))) return rc;
        double *b = (double *)p;
        auto put = [&](const double *h, size_t n, const double **d) -> int {
            if (n) CUDA_TRY(cudaMemcpyAsync(b, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            *d = b;
            b += n;
            return LUDVM_OK;
        };
        if (nw) {
            if ((rc = put(gamma, (size_t)ngamma, &dg))) return rc;
            if ((rc = put(xw, nsrc, &dxw))) return rc;
            if ((rc = put(zw, nsrc, &dzw))) return rc;
            if (vc4_per_source && (rc = put(vc4_per_source, nsrc, &dvc))) return rc;
        }
        if ((rc = put(xp, ntg, &dxp))) return rc;
        if ((rc = put(zp, ntg, &dzp))) return rc;
        void *o;
        if ((rc = scratch_reserve(ctx, 5, 2 * ntg * sizeof(double), &o))) return rc;
        du = (double *)o;
        dw = du + ntg;
    }
    if (nw == 0) {  // np.sum over an empty axis
        CUDA_TRY(cudaMemsetAsync(du, 0, (size_t)np_ * sizeof(double), ctx->stream));
        CUDA_TRY(cudaMemsetAsync(dw, 0, (size_t)np_ * sizeof(double), ctx->stream));
    } else {
        SrcView S = make_src(dg, ngamma == nw ? 1 : 0, dxw, dzw, dvc, vc4, (int)nw);
        TgtArray T{dxp, dzp};
        FusedOut fo{};
        fo.u = du;
        fo.w = dw;
        bool fused;
        if ((rc = try_launch_fused(ctx, mode, S, T, np_, fo, &fused))) return rc;
        if (!fused) {
            CombineArgs a{};
            if ((rc = launch_partials(ctx, mode, S, T, np_, 0, (double **)&a.pu, (double **)&a.pw, &a.nfold))) return rc;
            CUDA_TRY(cudaGetLastError());
            a.exact = (mode == LUDVM_EXACT_F64);
            a.nrows = (int)np_;
            a.u = du;
            a.w = dw;
            if ((rc = launch_combine(ctx, a))) return rc;
        }
    }
    if (ptr_kind == LUDVM_PTR_HOST) {
        CUDA_TRY(cudaMemcpyAsync(u, du, (size_t)np_ * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(w, dw, (size_t)np_ * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return LUDVM_OK;
}

LUDVM_API int ludvm_selfconv_step(ludvm_ctx *ctx, int mode, const double *gamma, const double *x, const double *z,
                                  const double *vc4_per_source, double vc4, long n, long row0, long nrows, double dt,
                                  double *x_out, double *z_out, double *u_out, double *w_out)
{
    ARG_CHECK(ctx != nullptr);
    int rc = check_mode(mode);
    if (rc) return rc;
    ARG_CHECK(n > 0 && n < (1L << 30) && row0 >= 0 && nrows >= 0 && row0 + nrows <= n);
    ARG_CHECK(gamma && x && z && x_out && z_out);
    if (nrows == 0) return LUDVM_OK;
    DeviceGuard g(ctx->device);
    SrcView S = make_src(gamma, 1, x, z, vc4_per_source, vc4, (int)n);
    TgtArray T{x + row0, z + row0};
    {
        FusedOut fo{};
        fo.u = u_out;
        fo.w = w_out;
        fo.x = x + row0;
        fo.z = z + row0;
        fo.xo = x_out + row0;
        fo.zo = z_out + row0;
        fo.dt = dt;
        bool fused;
        if ((rc = try_launch_fused(ctx, mode, S, T, nrows, fo, &fused))) return rc;
        if (fused) return LUDVM_OK;
    }
    CombineArgs a{};
    if ((rc = launch_partials(ctx, mode, S, T, nrows, 0, (double **)&a.pu, (double **)&a.pw, &a.nfold))) return rc;
    CUDA_TRY(cudaGetLastError());
    a.exact = (mode == LUDVM_EXACT_F64);
    a.nrows = (int)nrows;
    a.u = u_out;
    a.w = w_out;
    a.x = x + row0;
    a.z = z + row0;
    a.xo = x_out + row0;
    a.zo = z_out + row0;
    a.dt = dt;
    return launch_combine(ctx, a);
}

LUDVM_API int ludvm_selfconv_step_p2p(ludvm_ctx *ctx, int mode, const double *gamma, const double *x, const double *z,
                                      const double *vc4_per_source, double vc4, long n, long row0, long nrows,
                                      double dt, int npeers, double *const *x_out_peers, double *const *z_out_peers)
{
    ARG_CHECK(ctx != nullptr);
    int rc = check_mode(mode);
    if (rc) return rc;
    ARG_CHECK(n > 0 && n < (1L << 30) && row0 >= 0 && nrows >= 0 && row0 + nrows <= n);
    ARG_CHECK(gamma && x && z && x_out_peers && z_out_peers && npeers >= 1 && npeers <= LUDVM_MAX_PEERS);
    if (nrows == 0) return LUDVM_OK;
    DeviceGuard g(ctx->device);
    SrcView S = make_src(gamma, 1, x, z, vc4_per_source, vc4, (int)n);
    TgtArray T{x + row0, z + row0};
    for (int p = 0; p < npeers; p++) ARG_CHECK(x_out_peers[p] && z_out_peers[p]);
    {
        FusedOut fo{};
        fo.x = x + row0;
        fo.z = z + row0;
        fo.dt = dt;
        fo.npeers = npeers;
        for (int p = 0; p < npeers; p++) {
            fo.xo_peer[p] = x_out_peers[p] + row0;
            fo.zo_peer[p] = z_out_peers[p] + row0;
        }
        bool fused;
        if ((rc = try_launch_fused(ctx, mode, S, T, nrows, fo, &fused))) return rc;
        if (fused) return LUDVM_OK;
    }
    CombineArgs a{};
    if ((rc = launch_partials(ctx, mode, S, T, nrows, 0, (double **)&a.pu, (double **)&a.pw, &a.nfold))) return rc;
    CUDA_TRY(cudaGetLastError());
    a.exact = (mode == LUDVM_EXACT_F64);
    a.nrows = (int)nrows;
    a.x = x + row0;
    a.z = z + row0;
    a.dt = dt;
    a.npeers = npeers;
    for (int p = 0; p < npeers; p++) {
        ARG_CHECK(x_out_peers[p] && z_out_peers[p]);
        a.xo_peer[p] = x_out_peers[p] + row0;
        a.zo_peer[p] = z_out_peers[p] + row0;
    }
    return launch_combine(ctx, a);
}

LUDVM_API int ludvm_flowfield_velocity(ludvm_ctx *ctx, int mode, const double *ga, const double *xa, const double *za,
                                       long na, const double *gb, const double *xb, const double *zb, long nb,
                                       double vc4, const double *x1, long nx, const double *z1, long nz, long row0,
                                       long nrows, double *u, double *w, int ptr_kind)
{
    ARG_CHECK(ctx != nullptr);
    int rc = check_mode(mode);
    if (rc) return rc;
    ARG_CHECK(na > 0 && nb >= 0 && nx > 0 && nz > 0 && row0 >= 0 && nrows >= 0 && row0 + nrows <= nx);
    ARG_CHECK(na < (1L << 30) && nb < (1L << 30) && nrows * nz < (1L << 31) - 1024);
    ARG_CHECK(ga && xa && za && x1 && z1 && u && w && (nb == 0 || (gb && xb && zb)));
    ARG_CHECK(ptr_kind == LUDVM_PTR_HOST || ptr_kind == LUDVM_PTR_DEVICE);
    if (nrows == 0) return LUDVM_OK;
    DeviceGuard g(ctx->device);
    long npts = nrows * nz;
    const double *dga = ga, *dxa = xa, *dza = za, *dgb = gb, *dxb = xb, *dzb = zb, *dx1 = x1, *dz1 = z1;
    double *du = u, *dw = w;
    if (ptr_kind == LUDVM_PTR_HOST) {
        size_t total = 3 * (size_t)na + 3 * (size_t)nb + (size_t)nx + (size_t)nz + 8;
        void *p;
        if ((rc = scratch_reserve(ctx, 4, total * sizeof(double), &p))) return rc;
        double *b = (double *)p;
        auto put = [&](const double *h, size_t n, const double **d) -> int {
            if (n) CUDA_TRY(cudaMemcpyAsync(b, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            *d = b;
            b += n;
            return LUDVM_OK;
        };
        if ((rc = put(ga, na, &dga)) || (rc = put(xa, na, &dxa)) || (rc = put(za, na, &dza))) return rc;
        if (nb && ((rc = put(gb, nb, &dgb)) || (rc = put(xb, nb, &dxb)) || (rc = put(zb, nb, &dzb)))) return rc;
        if ((rc = put(x1, nx, &dx1)) || (rc = put(z1, nz, &dz1))) return rc;
        void *o;
        if ((rc = scratch_reserve(ctx, 5, 2 * (size_t)npts * sizeof(double), &o))) return rc;
        du = (double *)o;
        dw = du + npts;
    }
    TgtGrid T{dx1, dz1, (int)nz, (int)row0};
    SrcView SA = make_src(dga, 1, dxa, dza, nullptr, vc4, (int)na);
    bool fused = false;
    if (nb == 0) {   // one source set: the whole sum in one launch when eligible
        FusedOut fo{};
        fo.u = du;
        fo.w = dw;
        if ((rc = try_launch_fused(ctx, mode, SA, T, npts, fo, &fused))) return rc;
    }
    if (!fused) {
        CombineArgs a{};
        if ((rc = launch_partials(ctx, mode, SA, T, npts, 0, (double **)&a.pu, (double **)&a.pw, &a.nfold))) return rc;
        if (nb) {
            SrcView SB = make_src(dgb, 1, dxb, dzb, nullptr, vc4, (int)nb);
            if ((rc = launch_partials(ctx, mode, SB, T, npts, 2, (double **)&a.qu, (double **)&a.qw, &a.qfold))) return rc;
        }
        CUDA_TRY(cudaGetLastError());
        a.exact = (mode == LUDVM_EXACT_F64);
        a.nrows = (int)npts;
        a.u = du;
        a.w = dw;
        if ((rc = launch_combine(ctx, a))) return rc;
    }
    if (ptr_kind == LUDVM_PTR_HOST) {
        CUDA_TRY(cudaMemcpyAsync(u, du, (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(w, dw, (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return LUDVM_OK;
}

// Velocity and vorticity of one snapshot in one call: with host buffers the fields make one trip (u, w, ome out) instead
// of u, w out, back in for the stencil, and ome out.
LUDVM_API int ludvm_flowfield(ludvm_ctx *ctx, int mode, const double *ga, const double *xa, const double *za, long na,
                              const double *gb, const double *xb, const double *zb, long nb, double vc4,
                              const double *x1, long nx, const double *z1, long nz, long row0, long nrows, double *u,
                              double *w, double *ome, int ptr_kind)
{
    ARG_CHECK(ctx && ome && u && w && x1 && z1);
    ARG_CHECK(ptr_kind == LUDVM_PTR_HOST || ptr_kind == LUDVM_PTR_DEVICE);
    ARG_CHECK(nrows >= 2 && nz >= 2 && row0 >= 0 && row0 + nrows <= nx);
    if (ptr_kind == LUDVM_PTR_DEVICE) {
        int rc = ludvm_flowfield_velocity(ctx, mode, ga, xa, za, na, gb, xb, zb, nb, vc4, x1, nx, z1, nz, row0, nrows, u, w,
                                          ptr_kind);
        if (rc) return rc;
        return ludvm_flowfield_vorticity(ctx, x1 + row0, nrows, z1, nz, u, w, 1, ome, ptr_kind);
    }
    DeviceGuard g(ctx->device);
    // device-resident u, w, ome and axes in scratch slot 5 / 6; the velocity call stages sources and axes itself (slot 4)
    const size_t npts = (size_t)nrows * (size_t)nz;
    void *p, *q;
    int rc;
    if ((rc = scratch_reserve(ctx, 5, 3 * npts * sizeof(double), &p))) return rc;
    if ((rc = scratch_reserve(ctx, 6, ((size_t)nx + nz + 8) * sizeof(double), &q))) return rc;
    double *du = (double *)p, *dw = du + npts, *dome = dw + npts, *dx1 = (double *)q, *dz1 = dx1 + nx;
    CUDA_TRY(cudaMemcpyAsync(dx1, x1, (size_t)nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(dz1, z1, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // sources: staged by a host-pointer call would also download u, w; stage them here and run on device pointers
    const size_t nsrc = 3 * (size_t)na + 3 * (size_t)nb;
    void *sp;
    if ((rc = scratch_reserve(ctx, 4, (nsrc + 8) * sizeof(double), &sp))) return rc;
    double *b = (double *)sp;
    const double *hsrc[6] = {ga, xa, za, gb, xb, zb};
    const double *dsrc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 6; k++) {
        const size_t n = k < 3 ? (size_t)na : (size_t)nb;
        if (n == 0) continue;
        ARG_CHECK(hsrc[k] != nullptr);
        CUDA_TRY(cudaMemcpyAsync(b, hsrc[k], n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dsrc[k] = b;
        b += n;
    }
    if ((rc = ludvm_flowfield_velocity(ctx, mode, dsrc[0], dsrc[1], dsrc[2], na, dsrc[3], dsrc[4], dsrc[5], nb, vc4, dx1, nx,
                                       dz1, nz, row0, nrows, du, dw, LUDVM_PTR_DEVICE)))
        return rc;
    if ((rc = ludvm_flowfield_vorticity(ctx, dx1 + row0, nrows, dz1, nz, du, dw, 1, dome, LUDVM_PTR_DEVICE))) return rc;
    CUDA_TRY(cudaMemcpyAsync(u, du, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(w, dw, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(ome, dome, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return LUDVM_OK;
}

LUDVM_API int ludvm_flowfield_vorticity(ludvm_ctx *ctx, const double *x1, long nx, const double *z1, long nz,
                                        const double *u, const double *w, long ns, double *ome, int ptr_kind)
{
    ARG_CHECK(ctx && x1 && z1 && u && w && ome);
    ARG_CHECK(nx >= 2 && nz >= 2 && ns >= 1);
    ARG_CHECK(ptr_kind == LUDVM_PTR_HOST || ptr_kind == LUDVM_PTR_DEVICE);
    DeviceGuard g(ctx->device);
    size_t tot = (size_t)ns * nx * nz;
    const double *dx1 = x1, *dz1 = z1, *du = u, *dw = w;
    double *dome = ome;
    int rc;
    if (ptr_kind == LUDVM_PTR_HOST) {
        double *a, *b, *c;
        if ((rc = stage_in(ctx, 6, x1, nx, &a))) return rc;
        dx1 = a;
        void *p;
        if ((rc = scratch_reserve(ctx, 4, (2 * tot + nz + 8) * sizeof(double), &p))) return rc;
        b = (double *)p;
        CUDA_TRY(cudaMemcpyAsync(b, z1, nz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dz1 = b;
        b += nz;
        CUDA_TRY(cudaMemcpyAsync(b, u, tot * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        du = b;
        b += tot;
        CUDA_TRY(cudaMemcpyAsync(b, w, tot * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dw = b;
        if ((rc = scratch_reserve(ctx, 5, tot * sizeof(double), &p))) return rc;
        c = (double *)p;
        dome = c;
    }
    k_vorticity<<<ceil_div((long)tot, 256), 256, 0, ctx->stream>>>(dx1, (int)nx, dz1, (int)nz, du, dw, (int)ns, dome);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    if (ptr_kind == LUDVM_PTR_HOST) {
        CUDA_TRY(cudaMemcpyAsync(ome, dome, tot * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return LUDVM_OK;
}
