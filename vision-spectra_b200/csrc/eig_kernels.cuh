// eig_kernels.cuh -- CTA wrappers around tridiag.cuh / bisect_metrics.cuh.
//
//   tridiag_fused_kernel  n <= kSmemMaxN : (tridiag_fused.cuh) Gram (packed lower) copied
//                         into shared memory once, reduced there, only d/e go back.
//   tridiag_global_kernel larger n       : Gram (full, symmetric) reduced in place in
//                         global memory / L2 with coalesced column walks.
//   bisect_metrics_kernel all n          : d/e -> sorted eigenvalues -> singular values,
//                         the four metrics and the integer outputs, one record per matrix.
#pragma once

#include "bisect_metrics.cuh"
#include "tridiag.cuh"
#include "tridiag_fused.cuh"

namespace vsp {

constexpr int kSmemMaxN = 256;  // fused path: 4 chunk pairs of 64 columns; rows beyond 113 KB stay in L2

__host__ __device__ inline size_t tridiag_global_smem_bytes(int npad) {
    return sizeof(double) * (5 * (size_t)npad + CtaCtx::kScratchDoubles);
}
__host__ __device__ inline size_t bisect_smem_bytes(int npad) {
    return sizeof(double) * (5 * (size_t)npad + CtaCtx::kScratchDoubles);
}
__host__ __device__ inline int bisect_threads(int n) {  // two eigenvalues per thread
    int t = (((n + 1) >> 1) + 31) & ~31;
    return t > 1024 ? 1024 : (t < 32 ? 32 : t);
}

__global__ void tridiag_global_kernel(const ItemDesc* __restrict__ items, int item_base, double* __restrict__ ws,
                                      int npad) {
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
        out[2 * n + MISC_SCALE] = scale;
        out[2 * n + MISC_FLAGS] = 0.0;
    }
}

__global__ void bisect_metrics_kernel(const ItemDesc* __restrict__ items, int item_base,
                                      const double* __restrict__ ws, int npad, vsp_opts opts,
                                      double* __restrict__ sv_out, vsp_record* __restrict__ records) {
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
    ctx.sync();
    int iters = 0;
    if (!flags) {
        const TriInfo t = tri_bounds(ctx, d, e, n);
        for (int i = ctx.tid; i < n; i += ctx.nthreads) {
            de[i].d = d[i];
            de[i].e2 = (i > 0) ? fmax(e[i - 1] * e[i - 1], kE2Floor) : 0.0;
        }
        ctx.sync();
        iters = bisect_all(ctx, de, n, t, lam);
    } else {
        for (int i = ctx.tid; i < n; i += ctx.nthreads) lam[i] = 0.0;
    }
    iters = ctx.max_i(iters);  // barrier: lam[] complete
    double* sv = (opts.want_sv != 0 && sv_out != nullptr) ? sv_out + it.sv_off : nullptr;
    const MetricOut mo = spectral_metrics(ctx, lam, n, scale, flags, opts.fit_start, opts.fit_end, opts.hill_k, sv);
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

}  // namespace vsp
