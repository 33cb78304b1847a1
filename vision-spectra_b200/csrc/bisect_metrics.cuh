// bisect_metrics.cuh -- stage 2b + stage 3: eigenvalues of the symmetric
// tridiagonal matrix by Sturm-count bisection (one eigenvalue per work item, so the
// spectrum comes out already sorted), then the reference's four metrics fused on
// the eigenvalues while they are still on chip.
//
// Reference semantics restated (file:line in vision_spectra/metrics/spectral.py):
//   singular values  s_i = sqrt(lambda_i(W^T W)), descending        :91 (scipy svd)
//   spectral_entropy : keep s > 0, p = s^2 / sum s^2, H = -sum p ln p   :96-109
//   stable_rank      : sum s^2 / max(s)^2                                :162-173
//   alpha_exponent   : m = #positive; m < 8 -> NaN; start = max(1, int(.10 m)),
//                      end = min(max(start+6, int(.60 m)), m); OLS slope of ln s_i on
//                      ln(rank), rank = index+1, over [start,end); return -slope  :243-273
//   power_law_alpha_hill : n = #(s^2 > 0); n < 8 -> NaN; k = min(max(5,int(.1 n)),
//                      max(5,n-1)); H = mean ln(lambda_(i)/lambda_(k)), i < k;
//                      H <= 0 -> NaN; 1 + 1/H                            :344-368
// All four are invariant under W -> cW, so they are evaluated on the scaled
// eigenvalues; only the singular values are un-scaled.
#pragma once

#include "common.cuh"

namespace vsp {

struct TriInfo {
    double gl, gu;   // Gershgorin interval, widened
    double atol;     // absolute width at which bisection stops
};

// sign / zero tests on the bit pattern: integer pipe on the device, so the FP64 pipe
// only sees the three arithmetic instructions of the recurrence.
VSP_DEV bool dbl_neg(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x) < 0;
#else
    return std::signbit(x);
#endif
}
VSP_DEV bool dbl_zero(double x) {
#if defined(__CUDA_ARCH__)
    return ((__double2hiint(x) & 0x7fffffff) | __double2loint(x)) == 0;
#else
    return x == 0.0;
#endif
}
// 2^(1023 - biased_exponent(max(|a|,|b|))): multiplying by it brings the pair back to O(1).
VSP_DEV double rescale_factor(double a, double b) {
#if defined(__CUDA_ARCH__)
    int ea = (__double2hiint(a) >> 20) & 0x7ff, eb = (__double2hiint(b) >> 20) & 0x7ff;
    int ex = ea > eb ? ea : eb;
    if (ex == 0 || ex == 0x7ff) return 1.0;
    return __hiloint2double((2046 - ex) << 20, 0);
#else
    int ea, eb;
    (void)frexp(a, &ea);
    (void)frexp(b, &eb);
    if (a == 0.0 && b == 0.0) return 1.0;
    int ex = (a == 0.0) ? eb : (b == 0.0 ? ea : (ea > eb ? ea : eb));
    return ldexp(1.0, 1 - ex);
#endif
}

// Number of eigenvalues of T (diag d, squared off-diagonals e2) below x, as the number
// of sign changes in the Sturm sequence p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}
// (division free: three FP64 instructions per row).  A zero p_i takes the sign opposite
// to p_{i-1}, which is LAPACK dlaebz's "pivot <= 0 counts, |pivot| < pivmin -> -pivmin".
// The pair is renormalised every 8 rows; T is pre-scaled to ||T|| <= O(n), so |p| cannot
// overflow or vanish between renormalisations.
VSP_DEV int sturm_count(const double* d, const double* e2, int n, double x) {
    double p0 = 1.0;
    double p1 = d[0] - x;
    if (dbl_zero(p1)) p1 = -2.4e-181;
    int c = dbl_neg(p1) ? 1 : 0;
    for (int i = 1; i < n; ++i) {
        const double t = d[i] - x;
        double p2 = fma(t, p1, -(e2[i - 1] * p0));
        if (dbl_zero(p2)) p2 = -p1 * 2.4e-181;
        c += (dbl_neg(p2) != dbl_neg(p1)) ? 1 : 0;
        p0 = p1;
        p1 = p2;
        if ((i & 7) == 7) {
            const double s = rescale_factor(p0, p1);
            p0 *= s;
            p1 *= s;
        }
    }
    return c;
}

template <class Ctx>
VSP_DEV TriInfo tri_bounds(Ctx& ctx, const double* d, const double* e, int n) {
    double lo = 1e300, hi = -1e300;
    for (int i = ctx.tid; i < n; i += ctx.nthreads) {
        const double el = (i > 0) ? fabs(e[i - 1]) : 0.0;
        const double er = (i < n - 1) ? fabs(e[i]) : 0.0;
        lo = fmin(lo, d[i] - el - er);
        hi = fmax(hi, d[i] + el + er);
    }
    lo = ctx.min(lo);
    hi = ctx.max(hi);
    TriInfo t;
    const double bnorm = fmax(fabs(lo), fabs(hi));
    const double widen = 2.0 * bnorm * 2.220446049250313e-16 * n + 4.4501477170144028e-308;
    t.gl = lo - widen;
    t.gu = hi + widen;
    // Absolute floor of the interval width.  Householder reduction of a graded Gram matrix
    // keeps small eigenvalues far better than its eps ||T|| worst case (power-law spectra,
    // tests/test_metrics.py:113-135), so resolve well below eps ||T||: 2^-66 ||T||.
    t.atol = bnorm * 1.3552527156068805e-20;
    return t;
}

// lam[k], k = 0..n-1 ascending.  e2[] must hold squared off-diagonals.  Returns the
// largest iteration count used by this work item.
template <class Ctx>
VSP_DEV int bisect_all(Ctx& ctx, const double* d, const double* e2, int n, const TriInfo& t, double* lam) {
    int maxit = 0;
    for (int k = ctx.tid; k < n; k += ctx.nthreads) {
        double lo = t.gl, hi = t.gu;
        int it = 0;
        for (; it < 128; ++it) {
            const double mid = 0.5 * (lo + hi);
            const double tol = fmax(t.atol, 4.440892098500626e-16 * fmax(fabs(lo), fabs(hi)));
            if (hi - lo <= tol || mid <= lo || mid >= hi) break;
            if (sturm_count(d, e2, n, mid) >= k + 1)
                hi = mid;
            else
                lo = mid;
        }
        lam[k] = 0.5 * (lo + hi);
        maxit = it > maxit ? it : maxit;
    }
    return maxit;
}

struct MetricOut {
    double metrics[4];
    int m, start, end, k, status;
};

// lam[] ascending, in scaled units (scale = power of four applied to the Gram
// matrix).  Writes sv[0..n) descending (if sv != nullptr) and returns the record
// fields.  `flags` carries VSP_ST_NONFINITE / VSP_ST_ZERO from the Gram stage.
template <class Ctx>
VSP_DEV MetricOut spectral_metrics(Ctx& ctx, double* lam, int n, double scale, int flags,
                                   int fit_start, int fit_end, int hill_k, double* sv) {
    MetricOut out;
    const double nan = NAN;
    out.metrics[0] = out.metrics[1] = out.metrics[2] = out.metrics[3] = nan;
    out.m = 0;
    out.start = out.end = out.k = -1;
    out.status = flags;
    if (flags & VSP_ST_NONFINITE) {
        if (sv)
            for (int i = ctx.tid; i < n; i += ctx.nthreads) sv[i] = nan;
        return out;
    }
    const double lmax = lam[n - 1];
    if ((flags & VSP_ST_ZERO) || !(lmax > 0.0)) {
        out.status |= VSP_ST_ZERO;
        if (sv)
            for (int i = ctx.tid; i < n; i += ctx.nthreads) sv[i] = 0.0;
        return out;
    }
    // Eigenvalues of a Gram matrix are >= 0; rounding can push the smallest ones to
    // <= 0.  LAPACK on W itself reports them as tiny positive numbers, so
    // floor at (eps * sigma_max)^2 instead of dropping them: m stays min(rows, cols).
    const double lfloor = lmax * 4.930380657631324e-32;  // 2^-104
    ctx.sync();
    for (int i = ctx.tid; i < n; i += ctx.nthreads)
        if (!(lam[i] > lfloor)) lam[i] = lfloor;
    ctx.sync();
    const int m = n;  // every floored eigenvalue is positive and finite
    out.m = m;

    // singular values, descending; sqrt(1/scale) is an exact power of two
    if (sv) {
        const double unscale = sqrt(1.0 / scale);
        for (int i = ctx.tid; i < n; i += ctx.nthreads) sv[i] = sqrt(lam[n - 1 - i]) * unscale;
    }

    // ---- entropy and stable rank
    double s2 = 0.0;
    for (int i = ctx.tid; i < n; i += ctx.nthreads) s2 += lam[i];
    const double total = ctx.sum(s2);
    double h = 0.0;
    for (int i = ctx.tid; i < n; i += ctx.nthreads) {
        const double p = lam[i] / total;
        if (p > 0.0) h -= p * log(p);
    }
    out.metrics[0] = ctx.sum(h);
    out.metrics[1] = total / lmax;

    // ---- alpha: OLS slope of ln sigma on ln rank over [start, end)
    int start = -1, end = -1;
    if (fit_start < 0 && fit_end < 0) {
        if (m >= 8) {
            start = (int)(0.10 * (double)m);
            if (start < 1) start = 1;
            end = (int)(0.60 * (double)m);
            if (end < start + 6) end = start + 6;
            if (end > m) end = m;
            if (end - start < 2) start = end = -1;
        } else {
            out.status |= VSP_ST_FEW_SV;
        }
    } else if (fit_start >= 0 && fit_end <= m && fit_end - fit_start >= 2) {
        start = fit_start;
        end = fit_end;
    }
    out.start = start;
    out.end = end;
    if (start >= 0) {
        const double cnt = (double)(end - start);
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = start + ctx.tid; i < end; i += ctx.nthreads) {
            acc[0] += log((double)(i + 1));
            acc[1] += 0.5 * log(lam[n - 1 - i]);
        }
        ctx.sum4(acc);
        const double xbar = acc[0] / cnt, ybar = acc[1] / cnt;
        double acc2[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = start + ctx.tid; i < end; i += ctx.nthreads) {
            const double dx = log((double)(i + 1)) - xbar;
            const double dy = 0.5 * log(lam[n - 1 - i]) - ybar;
            acc2[0] += dx * dy;
            acc2[1] += dx * dx;
        }
        ctx.sum4(acc2);
        const double slope = acc2[0] / acc2[1];
        if (isfinite(slope))
            out.metrics[2] = -slope;
        else
            out.status |= VSP_ST_ALPHA_NAN;
    } else if (!(out.status & VSP_ST_FEW_SV)) {
        out.status |= VSP_ST_ALPHA_NAN;
    }

    // ---- Hill estimator on the k largest eigenvalues
    if (m >= 8) {
        int k = hill_k;
        if (k < 0) {
            k = (int)(0.10 * (double)m);
            if (k < 5) k = 5;
            const int cap = (m - 1 > 5) ? m - 1 : 5;
            if (k > cap) k = cap;
        }
        out.k = k;
        const int keff = k < m ? k : m;  // numpy slicing [:k] clips at n
        if (keff >= 1) {
            const double xmin = lam[n - keff];
            double acc = 0.0;
            for (int i = ctx.tid; i < keff; i += ctx.nthreads) acc += log(lam[n - 1 - i] / xmin);
            const double hm = ctx.sum(acc) / (double)keff;
            if (hm > 0.0 && isfinite(hm))
                out.metrics[3] = 1.0 + 1.0 / hm;
            else
                out.status |= VSP_ST_HILL_NAN;
        } else {
            out.status |= VSP_ST_HILL_NAN;
        }
    } else {
        out.status |= VSP_ST_FEW_SV;
    }
    return out;
}

}  // namespace vsp
