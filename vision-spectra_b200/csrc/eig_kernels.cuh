// eig_kernels.cuh -- CTA wrappers around tridiag.cuh / bisect_metrics.cuh.
//
//   sbr_band_kernel +     n <= kSmemMaxN : (sbr_band.cuh, band_tridiag.cuh) Gram (packed lower) held in shared
//   band_tridiag_kernel                    memory, blocked Householder reduction to bandwidth 4 on the FP64
//                                          tensor cores, then pipelined bulge chasing to (d, e).
//   tridiag_global_kernel larger n       : Gram (full, symmetric) reduced in place in
//                         global memory / L2 with coalesced column walks.
//   bisect_metrics_kernel all n          : d/e -> sorted eigenvalues -> singular values,
//                         the four metrics and the integer outputs, one record per matrix.
#pragma once

#include "bisect_metrics.cuh"
#include "refine_bidiag.cuh"
#include "tridiag.cuh"
#include "sbr_band.cuh"
#include "sbr8.cuh"
#include "chase8.cuh"

namespace vsp {

constexpr int kSmemMaxN = 768;  // blocked band reduction (packed Gram): one warp per 32-index block up to 256, three
                                // blocks per warp with the matrix in the global workspace up to 768

__host__ __device__ inline size_t tridiag_global_smem_bytes(int npad) {
    return sizeof(double) * (5 * (size_t)npad + CtaCtx::kScratchDoubles);
}
__host__ __device__ inline size_t bisect_smem_bytes(int npad) {
    // + the eigenvalue work counter + the coarse grid (bisect_threads(npad) >= the launch's thread count)
    return sizeof(double) * (5 * (size_t)npad + CtaCtx::kScratchDoubles + 2 + 3 * (2 * (size_t)(((npad + 2) / 3 + 31) & ~31) + 2 + 64));
}
__host__ __device__ inline int bisect_threads(int n) {  // two brackets per thread, ~1.5 eigenvalues per bracket
    int t = ((n + 2) / 3 + 31) & ~31;
    return t > 1024 ? 1024 : (t < 32 ? 32 : t);
}

__global__ void tridiag_global_kernel(const ItemDesc* __restrict__ items, int item_base, double* __restrict__ ws,
                                      int npad, RefineGate gate) {
    extern __shared__ __align__(16) double smem[];
    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n;
    double* red = smem;
    double* v = red + CtaCtx::kScratchDoubles;
    double* p = v + npad;
    double* d = p + npad;
    double* e = d + npad;
    double* part = e + npad;
    CtaCtx ctx(red);

    double* G = ws + it.gram_off;
    double md = 0.0;
    int bad = 0;
    for (int i = ctx.tid; i < n; i += ctx.nthreads) {
        const double g = G[(int64_t)i * n + i];
        if (!isfinite(g)) bad = 1;
        md = fmax(md, g);
    }
    int flags = 0;
    const double scale = gram_scale(ctx, md, bad, &flags);
    double* out = ws + it.de_off;
    if (flags) {
        for (int i = ctx.tid; i < 2 * n; i += ctx.nthreads) out[i] = 0.0;
        if (ctx.tid == 0) {
            out[2 * n + MISC_SCALE] = 1.0;
            out[2 * n + MISC_FLAGS] = (double)flags;
        }
        return;
    }
    const int64_t total = (int64_t)n * n;
    for (int64_t i = ctx.tid; i < total; i += ctx.nthreads) G[i] *= scale;
    __threadfence_block();
    ctx.sync();
    tridiagonalize(ctx, FullSym{G, n}, n, npad, 1, d, e, v, p, part);
    for (int i = ctx.tid; i < n; i += ctx.nthreads) {
        out[i] = d[i];
        out[n + i] = e[i];
    }
    if (ctx.tid == 0) {
        int oflags = 0, slot = -1;
        const bool rounded = gate.inexact != nullptr && gate.inexact[item_base + blockIdx.x] != 0;
        if (gate.counter != nullptr && has_tiny_eigenvalue(d, e, n, rounded ? kRefineRatioInexact : kRefineRatio)) {
            slot = atomicAdd(gate.counter, 1);  // list entry (the list holds every item of the class)
            oflags = VSP_ST_ILLCOND;
            gate.slot_items[slot] = item_base + blockIdx.x;
        }
        out[2 * n + MISC_SCALE] = scale;
        out[2 * n + MISC_FLAGS] = (double)oflags;
        out[2 * n + MISC_SLOT] = (double)slot;
    }
}

// MAXT / MINB: launch bounds.  The Sturm recurrences are two dependent FP64 chains per thread, i.e. latency-bound:
// the small-order instantiation caps the registers so that 8 CTAs (24+ warps) share an SM.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bisect_metrics_kernel(const ItemDesc* __restrict__ items, int item_base,
                                      const double* __restrict__ ws, int npad, vsp_opts opts,
                                      double* __restrict__ sv_out, vsp_record* __restrict__ records, double* __restrict__ dist_out) {
    extern __shared__ __align__(16) double smem[];
    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n;
    double* red = smem;
    double* d = red + CtaCtx::kScratchDoubles;
    double* e = d + npad;
    double* lam = e + npad;
    DE* de = reinterpret_cast<DE*>(lam + npad);  // (d_i, max(e_{i-1}^2, floor)) interleaved for one 16-byte load per row
    CtaCtx ctx(red);

    const double* __restrict__ in = ws + it.de_off;
    for (int i = ctx.tid; i < n; i += ctx.nthreads) {
        d[i] = in[i];
        e[i] = in[n + i];
    }
    const double scale = in[2 * n + MISC_SCALE];
    const int flags = (int)in[2 * n + MISC_FLAGS];
    if (flags & VSP_ST_ILLCOND) return;  // claimed by refine_kernel, which runs concurrently and writes the record
    ctx.sync();
    int iters = 0;
    if (!flags) {
        const TriInfo t = tri_bounds(ctx, d, e, n);
        for (int i = ctx.tid; i < n; i += ctx.nthreads) {
            de[i].d = d[i];
            de[i].e2 = (i > 0) ? fmax(e[i - 1] * e[i - 1], kE2Floor) : 0.0;
        }
        int* next_k = reinterpret_cast<int*>(lam + 3 * npad);  // after (d_i, e2_i): the eigenvalue work counter
        if (ctx.tid == 0) *next_k = 0;
        ctx.sync();
        iters = bisect_all(ctx, de, n, t, lam, next_k, lam + 3 * npad + 2);
    } else {
        for (int i = ctx.tid; i < n; i += ctx.nthreads) lam[i] = 0.0;
    }
    iters = ctx.max_i(iters);  // barrier: lam[] complete
    double* sv = (opts.want_sv != 0 && sv_out != nullptr) ? sv_out + it.sv_off : nullptr;
    double* aux = dist_out != nullptr ? dist_out + (int64_t)it.item * VSP_AUX_STRIDE(opts.dist_k, opts.clauset) : nullptr;
    const MetricOut mo = spectral_metrics(ctx, lam, n, scale, flags, opts.fit_start, opts.fit_end, opts.hill_k, sv, aux, opts.dist_k,
                                          (aux != nullptr && opts.clauset) ? aux + VSP_AUX_STRIDE(opts.dist_k, 0) : nullptr);
    if (ctx.tid == 0) {
        vsp_record r;
        r.item = it.item;
        r.status = mo.status;
        r.m = mo.m;
        r.start = mo.start;
        r.end = mo.end;
        r.k = mo.k;
        r.n = n;
        r.iters = iters;
        r.metrics[0] = mo.metrics[0];
        r.metrics[1] = mo.metrics[1];
        r.metrics[2] = mo.metrics[2];
        r.metrics[3] = mo.metrics[3];
        records[it.item] = r;
    }
}

// ------------------------------------------------------------------------------------------
// Re-solve of the matrices the tridiagonalisation flagged VSP_ST_ILLCOND (refine_bidiag.cuh).
// One CTA per pool buffer: CTA b serves the list entries b, b + slots, b + 2 slots, ... with its own FP64
// buffer, so any number of flagged items is re-solved (in rounds when there are more than `slots`).
struct RefinePool {
    double* base;        // slots x slot_doubles
    int slots;
    int64_t slot_doubles;
};
constexpr int kRefineMaxN = 2048;
__host__ __device__ inline size_t refine_smem_fixed_bytes(int npad) {
    return sizeof(double) * (8 * (size_t)npad + 1024 + CtaCtx::kScratchDoubles + 2);
}

template <typename TIn>
__global__ void __launch_bounds__(1024)
    refine_kernel(const ItemDesc* __restrict__ items, RefineGate gate, RefinePool pool, int npad, int xs_doubles,
                  vsp_opts opts, double* __restrict__ sv_out, vsp_record* __restrict__ records, double* __restrict__ dist_out) {
    extern __shared__ __align__(16) double smem[];
    const int slot = blockIdx.x;
    const int nflagged = *gate.counter;
    for (int entry = slot; entry < nflagged; entry += pool.slots) {
    const ItemDesc it = items[gate.slot_items[entry]];
    const int n = it.n, K = it.kdim;
    double* red = smem;
    double* dq = red + CtaCtx::kScratchDoubles;
    double* eq = dq + npad;
    double* u = eq + npad;
    double* lam = u + npad;
    double* part = lam + npad;                                  // [1024]
    DE* de = reinterpret_cast<DE*>(part + 1024);                // [2 npad]
    int* slot_s = reinterpret_cast<int*>(de + 2 * (size_t)npad);
    CtaCtx ctx(red);
    double* X = pool.base + (int64_t)slot * pool.slot_doubles;
    const TIn* __restrict__ W = reinterpret_cast<const TIn*>(it.ptr);

    // scale by a power of two so that max |x| is in [0.5, 1)
    double mx = 0.0;
    for (int64_t e = ctx.tid; e < (int64_t)it.rows * it.cols; e += ctx.nthreads) {
        const int r = (int)(e / it.cols), c = (int)(e % it.cols);
        mx = fmax(mx, fabs((double)W[(int64_t)r * it.ld + c]));
    }
    mx = ctx.max(mx);
    int ex = 0;
    (void)frexp(mx, &ex);
    const double sc = ldexp(1.0, -ex);
    // X: column-major K x n with the Gram index on the columns
    if (!it.trans) {  // n = rows: column i of X is row i of W
        for (int64_t e = ctx.tid; e < (int64_t)n * K; e += ctx.nthreads) {
            const int i = (int)(e / K), k = (int)(e % K);
            X[e] = (double)W[(int64_t)i * it.ld + k] * sc;
        }
    } else {  // n = cols: column i of X is column i of W
        for (int64_t e = ctx.tid; e < (int64_t)n * K; e += ctx.nthreads) {
            const int k = (int)(e / n), i = (int)(e % n);  // read W coalesced
            X[(int64_t)i * K + k] = (double)W[(int64_t)k * it.ld + i] * sc;
        }
    }
    __threadfence_block();
    ctx.sync();
    // first steps on the global (L2) copy until the trailing block fits in shared memory
    double* Xs = reinterpret_cast<double*>(slot_s + 2);
    int j0 = 0;
    while (j0 < n && (int64_t)(K - j0) * (n - j0) > (int64_t)xs_doubles) ++j0;
    if (j0 > 0) bidiag_steps(ctx, X, K, K, n, 0, j0, dq, eq, u, part);
    if (j0 < n) {
        const int lds = K - j0, nc = n - j0;
        for (int64_t e = ctx.tid; e < (int64_t)lds * nc; e += ctx.nthreads) {
            const int c = (int)(e / lds), r = (int)(e % lds);
            Xs[e] = X[(int64_t)(c + j0) * K + (r + j0)];
        }
        ctx.sync();
        // same indexing X[c*ld + r] with c, r >= j0 on the shifted base
        bidiag_steps(ctx, Xs - ((int64_t)j0 * lds + j0), lds, K, n, j0, n, dq, eq, u, part);
    }
    int iters = gk_singular_values(ctx, dq, eq, n, de, lam);
    iters = ctx.max_i(iters);  // barrier: lam[] complete
    double* sv = (opts.want_sv != 0 && sv_out != nullptr) ? sv_out + it.sv_off : nullptr;
    double* aux = dist_out != nullptr ? dist_out + (int64_t)it.item * VSP_AUX_STRIDE(opts.dist_k, opts.clauset) : nullptr;
    const MetricOut mo = spectral_metrics(ctx, lam, n, sc * sc, 0, opts.fit_start, opts.fit_end, opts.hill_k, sv, aux, opts.dist_k,
                                          (aux != nullptr && opts.clauset) ? aux + VSP_AUX_STRIDE(opts.dist_k, 0) : nullptr);
    if (ctx.tid == 0) {
        vsp_record r;
        r.item = it.item;
        r.status = mo.status | VSP_ST_ILLCOND | VSP_ST_REFINED;
        r.m = mo.m;
        r.start = mo.start;
        r.end = mo.end;
        r.k = mo.k;
        r.n = n;
        r.iters = iters;
        r.metrics[0] = mo.metrics[0];
        r.metrics[1] = mo.metrics[1];
        r.metrics[2] = mo.metrics[2];
        r.metrics[3] = mo.metrics[3];
        records[it.item] = r;
    }
    ctx.sync();  // the next entry reuses the shared scratch and the pool buffer
    }
}

}  // namespace vsp

#include "refine_cluster.cuh"  // needs RefinePool
