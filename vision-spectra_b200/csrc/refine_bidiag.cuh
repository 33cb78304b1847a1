// refine_bidiag.cuh -- re-solve for ill-conditioned matrices.
//
// The Gram route squares the condition number: singular values below ~1e-5 sigma_max come
// out of an FP64 Gram matrix with fewer than five correct digits.  Matrices whose computed
// lambda_min / lambda_max falls under kRefineRatio are therefore solved again from W itself,
// still on the GPU: Golub-Kahan Householder bidiagonalisation of the FP64 copy of W
// (backward stable: every sigma is perturbed by O(eps sigma_max), i.e. LAPACK's accuracy,
// the reference's dgesdd takes the same route) followed by Sturm bisection on the
// Golub-Kahan tridiagonal form [0 B^T; B 0], which resolves tiny singular values of the
// bidiagonal to high relative accuracy.
//
// One CTA per flagged matrix, matrix in global memory (L2), column-major K x n with the Gram
// index on the columns.  Written against the cooperative context of common.cuh so the same
// source is checked on the host (tests/emul/host_emul.cpp).
#pragma once

#include "bisect_metrics.cuh"

namespace vsp {

// X: column-major K x n (K >= n, leading dimension K).  Outputs dq[n] (diagonal of B) and
// eq[n] (super-diagonal, eq[n-1] = 0).  Scratch: u[n], part[nthreads] (shared memory on the
// device).  Left reflectors are applied one warp per trailing column (lanes on rows: coalesced,
// the dot product needs warp shuffles only); right reflectors with a rows x column-groups
// decomposition whose partial row sums meet in `part`.
//
// Steps j_begin <= j < j_end only, on storage with leading dimension ld: the device kernel runs
// the first steps on the global copy and moves the (much smaller) trailing block into shared
// memory for the rest, where a step costs shared-memory instead of L2 latency.
template <class Ctx>
VSP_DEV void bidiag_steps(Ctx& ctx, double* X, int ld, int K, int n, int j_begin, int j_end, double* dq, double* eq,
                          double* u, double* part) {
    for (int j = j_begin; j < j_end; ++j) {
        double* col = X + j * ld;
        // ---- left reflector: annihilate X[j+1:K, j]
        double ss = 0.0;
        for (int r = j + 1 + ctx.tid; r < K; r += ctx.nthreads) ss += col[r] * col[r];
        const double xn2 = ctx.sum(ss);
        const double alpha = col[j];
        double beta = alpha, tau = 0.0, vs = 0.0;
        if (xn2 > 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
            tau = (beta - alpha) / beta;
            vs = 1.0 / (alpha - beta);
        }
        if (ctx.tid == 0) dq[j] = beta;
        if (tau != 0.0) {
            for (int r = j + 1 + ctx.tid; r < K; r += ctx.nthreads) col[r] *= vs;  // v (v_j = 1 implicit)
            ctx.sync();
        }
        // one warp per trailing column; the pass also collects |X[j, j+2:]|^2 of the updated row j for the
        // right reflector (one partial per warp), so that row needs no pass and no reduction of its own
        double s2 = 0.0;
        for (int c = j + 1 + ctx.warp; c < n; c += ctx.nwarps) {
            double* cc = X + c * ld;
            const double cj = cc[j];
            double w = 0.0;
            if (tau != 0.0) {
                double dot = 0.0;
                for (int r = j + 1 + ctx.lane; r < K; r += ctx.wsize) dot += col[r] * cc[r];
                w = tau * (ctx.warp_sum(dot) + cj);
                for (int r = j + 1 + ctx.lane; r < K; r += ctx.wsize) cc[r] -= w * col[r];
                if (ctx.lane == 0) cc[j] = cj - w;
            }
            if (c >= j + 2) s2 += (cj - w) * (cj - w);
        }
        if (ctx.lane == 0) part[ctx.warp] = s2;
        ctx.sync();
        // ---- right reflector: annihilate X[j, j+2:n]
        if (j + 1 >= n) {
            if (ctx.tid == 0) eq[j] = 0.0;
            break;
        }
        double yn2 = 0.0;
        for (int q = 0; q < ctx.nwarps; ++q) yn2 += part[q];
        const double a2 = X[(j + 1) * ld + j];
        double b2 = a2, tau2 = 0.0, us = 0.0;
        if (yn2 > 0.0) {
            b2 = -copysign(sqrt(a2 * a2 + yn2), a2);
            tau2 = (b2 - a2) / b2;
            us = 1.0 / (a2 - b2);
        }
        if (ctx.tid == 0) eq[j] = b2;
        if (tau2 != 0.0) {
            for (int c = j + 1 + ctx.tid; c < n; c += ctx.nthreads)
                u[c] = (c == j + 1) ? 1.0 : X[c * ld + j] * us;
            ctx.sync();  // u complete; every thread has read the warp partials
            // rows j+1..K-1, split into G column groups so that every thread has work
            const int nrows = K - (j + 1), ncols = n - (j + 1);
            int G = ctx.nthreads / (nrows > 0 ? nrows : 1);
            if (G < 1) G = 1;
            if (G > ncols) G = ncols;
            const int per = (ncols + G - 1) / G;
            if (G == 1) {
                for (int r = j + 1 + ctx.tid; r < K; r += ctx.nthreads) {
                    double y = 0.0;
                    for (int c = j + 1; c < n; ++c) y += X[c * ld + r] * u[c];
                    y *= tau2;
                    for (int c = j + 1; c < n; ++c) X[c * ld + r] -= y * u[c];
                }
            } else {  // device only (the host context has one thread): G * nrows <= nthreads
                const int g = ctx.tid / nrows, r = j + 1 + (ctx.tid - g * nrows);
                const bool active = g < G;
                const int c_lo = j + 1 + g * per, c_hi = (c_lo + per < n) ? c_lo + per : n;
                if (active) {
                    double y = 0.0;
                    for (int c = c_lo; c < c_hi; ++c) y += X[c * ld + r] * u[c];
                    part[ctx.tid] = y;
                }
                ctx.sync();
                if (active) {
                    double y = 0.0;
                    for (int q = 0; q < G; ++q) y += part[q * nrows + (r - (j + 1))];
                    y *= tau2;
                    for (int c = c_lo; c < c_hi; ++c) X[c * ld + r] -= y * u[c];
                }
            }
        }
        ctx.sync();
    }
    ctx.sync();
}

template <class Ctx>
VSP_DEV void bidiagonalize(Ctx& ctx, double* X, int K, int n, double* dq, double* eq, double* u, double* part) {
    bidiag_steps(ctx, X, K, K, n, 0, n, dq, eq, u, part);
}

// Singular values of the bidiagonal (dq, eq) as the n largest eigenvalues of the 2n x 2n
// Golub-Kahan tridiagonal (zero diagonal, off-diagonals dq0, eq0, dq1, eq1, ..., dq_{n-1}).
// de: scratch of 2n DE entries; lam_out[k] = sigma_k^2 ascending, k = 0..n-1.
template <class Ctx>
VSP_DEV int gk_singular_values(Ctx& ctx, const double* dq, const double* eq, int n, DE* de, double* lam_out) {
    const int n2 = 2 * n;
    double mx = 0.0;
    for (int i = ctx.tid; i < n2; i += ctx.nthreads) {
        double g = 0.0;  // off-diagonal coupling i-1 and i
        if (i > 0) g = ((i - 1) & 1) ? eq[(i - 1) >> 1] : dq[(i - 1) >> 1];
        de[i].d = 0.0;
        de[i].e2 = (i > 0) ? fmax(g * g, kE2Floor) : 0.0;
        mx = fmax(mx, fabs(g));
    }
    mx = ctx.max(mx);
    const double bound = 2.0 * mx * (1.0 + 2.220446049250313e-16 * n2) + 4.4501477170144028e-308;
    const double floor_w = bound * 8.470329472543003e-22;  // 2^-70 sigma_max: far below LAPACK's own noise
    ctx.sync();
    int maxit = 0;
    const int half = (n + 1) >> 1;
    for (int k = ctx.tid; k < half; k += ctx.nthreads) {
        // two singular values per work item (k-th and (k+half)-th smallest); index n + k in the
        // ascending 2n spectrum; only the non-negative half [0, bound] is searched, with the same
        // geometric / counting / interpolating brackets as the Gram route (bisect_metrics.cuh)
        const int kb = k + half;
        const bool has_b = kb < n;
        Bracket A, B;
        A.init(0.0, bound, n2, true);
        B.init(0.0, bound, n2, has_b);
        A.clo = B.clo = n;  // the Golub-Kahan spectrum is symmetric: n eigenvalues below zero
        int it = 0;
        for (; it < 400; ++it) {
            const double xa = A.next(floor_w), xb = B.next(floor_w);
            if (A.done && B.done) break;
            int na, nb, ea, eb;
            double fa, fb;
            sturm_count2(de, n2, xa, xb, na, nb, fa, ea, fb, eb);
            A.update(xa, na, fa, ea, n + k);
            B.update(xb, nb, fb, eb, n + kb);
        }
        const double sa = 0.5 * (A.lo + A.hi);
        lam_out[k] = sa * sa;
        if (has_b) {
            const double sb = 0.5 * (B.lo + B.hi);
            lam_out[kb] = sb * sb;
        }
        maxit = it > maxit ? it : maxit;
    }
    return maxit;
}

}  // namespace vsp
