// tridiag.cuh -- stage 2a: Householder reduction of a symmetric (Gram) matrix to
// tridiagonal form, one matrix per CTA, FP64.
//
// Replaces (together with bisect_metrics.cuh) the LAPACK dgesdd call behind
// scipy.linalg.svd at vision_spectra/metrics/spectral.py:91,158,239,339 and
// experiments/run_spectral_analysis.py:333.  Eigenvalues only are needed, so the
// reduction costs (4/3) n^3 flops instead of ~60 n^3 for Jacobi sweeps.
//
// Storage policies
//   PackedLower : row-major lower triangle, element (r,c), c<=r at tri(r)+c.  Used
//                 when the matrix is resident in shared memory (n <= ~230).  With
//                 lanes on consecutive rows and a common column the 8-byte accesses
//                 tri(r)+c hit 16 distinct banks per half-warp (triangular numbers
//                 are a permutation mod 16), so both the row walk and the column
//                 walk of the symmetric mat-vec are conflict free.
//   FullSym     : full n*n array in global memory (n too large for shared memory);
//                 element (r,c) is read at A[c*ld + r] so that lanes on consecutive
//                 rows r coalesce.
//
// Elimination order: last row first (row i = n-1 .. 1).  The reflector comes from
// the part of row i left of the diagonal, which is contiguous in both layouts, and
// the trailing block is always the leading i x i block, so the active work items
// stay packed at low indices.
#pragma once

#include "common.cuh"

namespace vsp {

struct PackedLower {
    double* a;
    int n;
    VSP_HD double* row(int r) const { return a + tri(r); }        // row r, columns 0..r
    VSP_HD double& sym(int r, int c) const { return (c <= r) ? a[tri(r) + c] : a[tri(c) + r]; }
    VSP_HD int owned_cols(int r, int m) const { (void)m; return r + 1; }  // columns this row stores
    VSP_HD double& own(int r, int c) const { return a[tri(r) + c]; }
};

struct FullSym {
    double* a;
    int n;
    VSP_HD double* row(int r) const { return a + (int64_t)r * n; }  // row r == column r
    VSP_HD double& sym(int r, int c) const { return a[(int64_t)c * n + r]; }
    VSP_HD int owned_cols(int r, int m) const { (void)r; return m; }  // both triangles are kept
    VSP_HD double& own(int r, int c) const { return a[(int64_t)c * n + r]; }
};

// Work-item decomposition: `split` cooperating items per row (item = s*npad + r
// handles columns c = s, s+split, ...).  On the device nthreads == split*npad.
//
// Scratch (doubles): v[n], p[n], part[split*npad].
// Outputs: d[n] diagonal, e[n] (e[i-1] couples i-1 and i; e[n-1] = 0).
template <class Ctx, class Store>
VSP_DEV void tridiagonalize(Ctx& ctx, Store A, int n, int npad, int split, double* d, double* e,
                            double* v, double* p, double* part) {
    const int nitems = split * npad;
    if (n == 1) {
        if (ctx.tid == 0) {
            d[0] = A.own(0, 0);
            e[0] = 0.0;
        }
        ctx.sync();
        return;
    }
    for (int i = n - 1; i >= 1; --i) {
        const int m = i;  // order of the leading block; row i holds x[0..m-1] | diag
        double* x = A.row(i);
        // ---- reflector: H x = (0,..,0,beta)
        double ss = 0.0;
        for (int j = ctx.tid; j < m - 1; j += ctx.nthreads) ss += x[j] * x[j];
        const double xnorm2 = ctx.sum(ss);
        const double alpha = x[m - 1];
        double beta = alpha, tau = 0.0, vscale = 0.0;
        if (xnorm2 > 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tau = (beta - alpha) / beta;
            vscale = 1.0 / (alpha - beta);
        }
        if (ctx.tid == 0) {
            e[i - 1] = beta;
            d[i] = x[m];
        }
        if (tau == 0.0) continue;  // uniform: every thread holds the same xnorm2
        for (int j = ctx.tid; j < m; j += ctx.nthreads) v[j] = (j == m - 1) ? 1.0 : x[j] * vscale;
        ctx.sync();
        // ---- p = tau * A v on the leading m x m block (each stored element visited twice)
        for (int it = ctx.tid; it < nitems; it += ctx.nthreads) {
            const int s = it / npad, r = it - s * npad;
            if (r < m) {
                double acc = 0.0;
                for (int c = s; c < m; c += split) acc += A.sym(r, c) * v[c];
                part[it] = acc;
            }
        }
        ctx.sync();
        double dot = 0.0;
        for (int r = ctx.tid; r < m; r += ctx.nthreads) {
            double acc = part[r];
            for (int s = 1; s < split; ++s) acc += part[s * npad + r];
            acc *= tau;
            p[r] = acc;
            dot += acc * v[r];
        }
        const double pv = ctx.sum(dot);  // barrier inside: p[] complete
        // ---- w = p - (tau/2)(p.v) v, kept in p
        const double a2 = -0.5 * tau * pv;
        for (int r = ctx.tid; r < m; r += ctx.nthreads) p[r] += a2 * v[r];
        ctx.sync();
        // ---- A <- A - v w^T - w v^T on the stored part of the leading block
        for (int it = ctx.tid; it < nitems; it += ctx.nthreads) {
            const int s = it / npad, r = it - s * npad;
            if (r < m) {
                const double vr = v[r], wr = p[r];
                const int lim = A.owned_cols(r, m);
                for (int c = s; c < lim; c += split) A.own(r, c) -= vr * p[c] + wr * v[c];
            }
        }
        ctx.sync();
    }
    if (ctx.tid == 0) {
        d[0] = A.own(0, 0);
        e[n - 1] = 0.0;
    }
    ctx.sync();
}

// Load-time conditioning shared by both storage policies: find the largest
// diagonal entry, flag non-finite input (a NaN/Inf anywhere in W reaches the Gram
// diagonal), and pick a power-of-four scale so that every |G_ij| <= 1 afterwards
// (|G_ij| <= sqrt(G_ii G_jj)).  sqrt(1/scale) is then an exact power of two, so
// un-scaling the singular values is exact.  Mirrors the scale robustness the
// reference gets from f64 LAPACK (tests/test_metrics.py:295-315).
template <class Ctx>
VSP_DEV double gram_scale(Ctx& ctx, double maxdiag_local, int nonfinite_local, int* flags_out) {
    const double maxdiag = ctx.max(maxdiag_local);
    const int nonfinite = ctx.max_i(nonfinite_local);
    int flags = 0;
    double scale = 1.0;
    if (nonfinite) {
        flags = VSP_ST_NONFINITE;
    } else if (!(maxdiag > 0.0)) {
        flags = VSP_ST_ZERO;
    } else {
        int ex;
        (void)frexp(maxdiag, &ex);  // maxdiag = f * 2^ex, f in [0.5,1)
        if (ex & 1) ex += 1;        // even exponent -> power of four
        scale = ldexp(1.0, -ex);
    }
    *flags_out = flags;
    return scale;
}

}  // namespace vsp
