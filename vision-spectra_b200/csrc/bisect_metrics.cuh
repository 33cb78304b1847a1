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

// 2^(1023 - biased_exponent(max(|a|,|b|))): multiplying by it brings the pair back to O(1).
// `e2` accumulates the binary exponent of the factors, so that the caller can recover the true
// magnitude of the Sturm terms: p_true = p_scaled * 2^(-e2).
VSP_DEV double rescale_factor(double a, double b, int& e2) {
#if defined(__CUDA_ARCH__)
    int ea = (__double2hiint(a) >> 20) & 0x7ff, eb = (__double2hiint(b) >> 20) & 0x7ff;
    int ex = ea > eb ? ea : eb;
    if (ex == 0 || ex == 0x7ff) return 1.0;
    e2 += 1023 - ex;
    return __hiloint2double((2046 - ex) << 20, 0);
#else
    int ea, eb;
    (void)frexp(a, &ea);
    (void)frexp(b, &eb);
    if (a == 0.0 && b == 0.0) return 1.0;
    int ex = (a == 0.0) ? eb : (b == 0.0 ? ea : (ea > eb ? ea : eb));
    e2 += 1 - ex;
    return ldexp(1.0, 1 - ex);
#endif
}
VSP_DEV double rescale_factor(double a, double b) {
    int unused = 0;
    return rescale_factor(a, b, unused);
}

// Sign-change bookkeeping on the integer pipe: the high words of consecutive Sturm terms
// are XORed and the resulting sign bit is funnel-shifted into a 32-bit history that is
// pop-counted every 32 rows.  The FP64 pipe only sees the three arithmetic instructions.
struct SignCounter {
    unsigned hist = 0;
    int count = 0;
    VSP_DEV void push(double prev, double cur) {
#if defined(__CUDA_ARCH__)
        hist = __funnelshift_l((unsigned)(__double2hiint(prev) ^ __double2hiint(cur)), hist, 1);
#else
        hist = (hist << 1) | (std::signbit(prev) != std::signbit(cur) ? 1u : 0u);
#endif
    }
    VSP_DEV void flush() {
#if defined(__CUDA_ARCH__)
        count += __popc(hist);
#else
        count += __builtin_popcount(hist);
#endif
        hist = 0;
    }
};

// Number of eigenvalues of T below xa and below xb (two shifts per work item share the
// loads and give the FP64 pipe two independent chains).  de[i] = (d_i, e2_{i-1}) with
// e2 = max(e^2, kE2Floor) and e2_{-1} = 0.  Sturm sequence in product form,
//     p_i = (d_i - x) p_{i-1} - e2_{i-1} p_{i-2},
// three FP64 instructions per row and shift; the count is the number of sign changes.
// Because e2 > 0, a term that is exactly zero is followed by one of the sign opposite to its
// predecessor, which is LAPACK dlaebz's convention (a zero pivot counts as negative), so no
// zero test is needed.  The pairs are renormalised every 8 rows; T is pre-scaled to
// ||T|| <= O(n), so |p| can neither overflow nor vanish in between.
constexpr double kE2Floor = 1e-200;

struct DE {
    double d, e2;
};

VSP_DEV void sturm_step(const DE r, double xa, double xb, double& a0, double& a1, double& b0, double& b1,
                        SignCounter& ca, SignCounter& cb) {
    const double a2 = fma(r.d - xa, a1, -(r.e2 * a0));
    const double b2 = fma(r.d - xb, b1, -(r.e2 * b0));
    ca.push(a1, a2);
    cb.push(b1, b2);
    a0 = a1;
    a1 = a2;
    b0 = b1;
    b1 = b2;
}

// Also returns the last Sturm term p_n(x) = det(T - x I) of each shift as (mantissa f, exponent e):
// p_n = f * 2^(-e).  Its sign is (-1)^count; bisect_all interpolates on it once a bracket is isolating.
VSP_DEV void sturm_count2(const DE* de, int n, double xa, double xb, int& na, int& nb, double& fa, int& ea,
                          double& fb, int& eb) {
    double a0 = 1.0, a1 = de[0].d - xa;
    double b0 = 1.0, b1 = de[0].d - xb;
    SignCounter ca, cb;
    ca.push(1.0, a1);
    cb.push(1.0, b1);
    int xa2 = 0, xb2 = 0;
    int i = 1;
    for (int blk = 0; i + 8 <= n; i += 8, ++blk) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sturm_step(de[i + k], xa, xb, a0, a1, b0, b1, ca, cb);
        const double sa = rescale_factor(a0, a1, xa2), sb = rescale_factor(b0, b1, xb2);
        a0 *= sa;
        a1 *= sa;
        b0 *= sb;
        b1 *= sb;
        if ((blk & 1) == 1) {  // 1 + 16 rows pushed at most 17..32 bits: flush before overflow
            ca.flush();
            cb.flush();
        }
    }
    ca.flush();
    cb.flush();
    for (; i < n; ++i) sturm_step(de[i], xa, xb, a0, a1, b0, b1, ca, cb);
    ca.flush();
    cb.flush();
    na = ca.count;
    nb = cb.count;
    fa = a1;
    ea = xa2;
    fb = b1;
    eb = xb2;
}
VSP_DEV void sturm_count2(const DE* de, int n, double xa, double xb, int& na, int& nb) {
    double fa, fb;
    int ea, eb;
    sturm_count2(de, n, xa, xb, na, nb, fa, ea, fb, eb);
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

// Bracket of one eigenvalue (index k, ascending): count(lo) <= k < count(hi).  Bisection on the counts until
// the bracket holds exactly one eigenvalue; from then on det(T - x I) changes sign exactly once inside it and
// the next shift comes from false position with the Illinois weights (superlinear), clamped `tol` away from
// the ends so that the bracket itself collapses to the tolerance (the stopping rule is the same as for plain
// bisection), with a bisection step after three interpolated shifts in a row that failed to halve the width.
// The counts alone decide which end moves: the interpolation only proposes shifts.
struct Bracket {
    double lo, hi, flo, fhi;
    int elo, ehi, clo, chi, side, slow;
    bool have_lo, have_hi, done, interp;

    VSP_DEV void init(double gl, double gu, int n, bool active) {
        lo = gl;
        hi = gu;
        flo = fhi = 0.0;
        elo = ehi = 0;
        clo = 0;
        chi = n;
        side = 0;
        have_lo = have_hi = false;
        done = !active;
        slow = 0;
        interp = false;
    }
    VSP_DEV double tolerance(double atol) const { return fmax(atol, 4.440892098500626e-16 * fmax(fabs(lo), fabs(hi))); }
    // next shift; sets `done` when the bracket has collapsed
    VSP_DEV double next(double atol) {
        const double mid = 0.5 * (lo + hi);
        interp = false;
        if (done) return mid;
        const double tol = tolerance(atol), width = hi - lo;
        if (width <= tol || mid <= lo || mid >= hi) {
            done = true;
            return mid;
        }
        // The spectrum of a Gram matrix is non-negative and spans many orders of magnitude.  While the bracket
        // is wide on the logarithmic scale (hi > 2 lo), search the magnitude geometrically -- 3 bits per step
        // from above while it still reaches (almost) down to zero, then geometric means -- instead of one bit
        // per arithmetic halving; det(T - x I) is far from linear across such a bracket, so no interpolation yet.
        if (hi > 0.0 && !(hi <= 2.0 * lo)) {
            const double x = (lo > 0.015625 * hi) ? sqrt(lo * hi) : ((lo > 0.0) ? fmax(sqrt(lo * hi), 0.125 * hi) : 0.125 * hi);
            if (x > lo && x < hi && hi - x > tol && x - lo > tol) return x;
            return mid;
        }
        if (chi - clo != 1 || !have_lo || !have_hi || (flo < 0.0) == (fhi < 0.0) || slow >= 3 || width <= 4.0 * tol) {
            slow = 0;
            return mid;
        }
        int de = ehi - elo;
        de = de > 1000 ? 1000 : (de < -1000 ? -1000 : de);
        const double r = ldexp(flo / fhi, de);  // f(lo) / f(hi) < 0
        const double tt = r / (r - 1.0);
        if (!(tt > 0.0 && tt < 1.0)) return mid;
        double x = fma(tt, width, lo);
        x = fmax(x, lo + tol);
        x = fmin(x, hi - tol);
        if (!(x > lo && x < hi)) return mid;
        interp = true;
        return x;
    }
    VSP_DEV void update(double x, int cnt, double f, int e, int k) {
        if (done) return;
        const double wold = hi - lo;
        if (cnt >= k + 1) {
            hi = x;
            chi = cnt;
            fhi = f;
            ehi = e;
            have_hi = true;
            if (side > 0) flo *= 0.5;  // the low end survived twice: Illinois
            side = 1;
        } else {
            lo = x;
            clo = cnt;
            flo = f;
            elo = e;
            have_lo = true;
            if (side < 0) fhi *= 0.5;
            side = -1;
        }
        // an interpolated shift that failed to halve the bracket counts as slow; three in a row force a bisection
        slow = (interp && hi - lo > 0.5 * wold) ? slow + 1 : 0;
    }
};

// Coarse grid shared by all brackets of one matrix: G - 1 shifts spaced geometrically over the 44 binary orders
// of magnitude below the Gershgorin upper bound, evaluated once (two per work item, one Sturm pass) before the
// brackets start.  An eigenvalue's bracket then starts as one grid cell -- already narrow on the log scale, with
// counts and det(T - xI) known at both ends -- instead of the whole Gershgorin interval: ~10 Sturm evaluations per
// eigenvalue instead of ~20.  Layout of `buf` (3 (G + 1) doubles): x[G+1] | f[G+1] | {count, exponent}[G+1].
struct CoarseGrid {
    double* x;
    double* f;
    int* ce;  // 2 ints per point: count below x, binary exponent of f
    int G;
    VSP_DEV static int points(int nthreads) { return 2 * nthreads + 1; }  // G
    VSP_DEV static int doubles(int nthreads) { return 3 * (points(nthreads) + 1); }
    VSP_DEV void bind(double* buf, int nthreads) {
        G = points(nthreads);
        x = buf;
        f = buf + (G + 1);
        ce = reinterpret_cast<int*>(buf + 2 * (G + 1));
    }
};

// lam[k], k = 0..n-1 ascending.  Every work item runs two brackets at a time (two independent FP64 chains);
// a bracket that has converged takes the next eigenvalue index from the shared counter `*next_k` (zeroed by the
// caller before the call), so that lanes whose eigenvalues converge early keep working instead of idling until
// the slowest lane of their warp is done.  `gridbuf`: CoarseGrid::doubles(nthreads) doubles of shared scratch.
// Returns the number of Sturm evaluations of this work item.
template <class Ctx>
VSP_DEV int bisect_all(Ctx& ctx, const DE* de, int n, const TriInfo& t, double* lam, int* next_k, double* gridbuf) {
    CoarseGrid grid;
    grid.bind(gridbuf, ctx.nthreads);
    const int G = grid.G;
    {   // grid points 1..G-1: x_i = gu * 2^(-44 (G - i) / (G - 1)); 0 and G are the Gershgorin ends
        const double step = -44.0 / (double)(G - 1);
        const int ia = 1 + 2 * ctx.tid, ib = 2 + 2 * ctx.tid;
        const double xa = (t.gu > 0.0) ? t.gu * exp2(step * (double)(G - ia)) : t.gu;
        const double xb = (t.gu > 0.0 && ib < G) ? t.gu * exp2(step * (double)(G - ib)) : t.gu;
        int na, nb, ea, eb;
        double fa, fb;
        sturm_count2(de, n, xa, xb, na, nb, fa, ea, fb, eb);
        if (ia < G) {
            grid.x[ia] = xa;
            grid.f[ia] = fa;
            grid.ce[2 * ia] = na;
            grid.ce[2 * ia + 1] = ea;
        }
        if (ib < G) {
            grid.x[ib] = xb;
            grid.f[ib] = fb;
            grid.ce[2 * ib] = nb;
            grid.ce[2 * ib + 1] = eb;
        }
        if (ctx.tid == 0) {
            grid.x[0] = t.gl;
            grid.f[0] = 0.0;
            grid.ce[0] = 0;
            grid.ce[1] = 0;
            grid.x[G] = t.gu;
            grid.f[G] = 0.0;
            grid.ce[2 * G] = n;
            grid.ce[2 * G + 1] = 0;
        }
    }
    ctx.sync();
    // bracket of eigenvalue k = the grid cell [i, i+1] with count(x_i) <= k < count(x_{i+1})
    auto start = [&](Bracket& br, int k) {
        br.init(t.gl, t.gu, n, k < n);
        if (k >= n || !(t.gu > 0.0)) return;
        int lo = 0, hi = G;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (grid.ce[2 * mid] >= k + 1) hi = mid; else lo = mid;
        }
        if (!(grid.x[lo] < grid.x[hi])) return;  // degenerate cell: keep the whole interval
        br.lo = grid.x[lo];
        br.hi = grid.x[hi];
        br.clo = grid.ce[2 * lo];
        br.chi = grid.ce[2 * hi];
        br.flo = grid.f[lo];
        br.fhi = grid.f[hi];
        br.elo = grid.ce[2 * lo + 1];
        br.ehi = grid.ce[2 * hi + 1];
        br.have_lo = lo >= 1;
        br.have_hi = hi <= G - 1;
        if (br.lo < t.gl) {  // the cell reaches below the Gershgorin interval (clustered spectra): clip it
            br.lo = t.gl;
            br.clo = 0;
            br.have_lo = false;
        }
    };
    int ka = ctx.fetch_add(next_k), kb = ctx.fetch_add(next_k);
    Bracket A, B;
    start(A, ka);
    start(B, kb);
    int it = 1;
    for (; it < 100000; ++it) {
        double xa = A.next(t.atol), xb = B.next(t.atol);
        if (A.done && ka < n) {
            lam[ka] = xa;  // next() returned the midpoint of the collapsed bracket
            ka = ctx.fetch_add(next_k);
            start(A, ka);
            xa = A.next(t.atol);
        }
        if (B.done && kb < n) {
            lam[kb] = xb;
            kb = ctx.fetch_add(next_k);
            start(B, kb);
            xb = B.next(t.atol);
        }
        if (ka >= n && kb >= n) break;
        int na, nb, ea, eb;
        double fa, fb;
        sturm_count2(de, n, xa, xb, na, nb, fa, ea, fb, eb);
        A.update(xa, na, fa, ea, ka);
        B.update(xb, nb, fb, eb, kb);
    }
    return it;
}

// Gram route: lambda_min <= ratio * lambda_max (kappa >~ 3e4) marks the matrix for the re-solve
// from W itself (refine_bidiag.cuh); below that ratio eps * kappa^2 eats into the 1e-5 gate.
constexpr double kRefineRatio = 1e-9;

// Early test used by the tridiagonalisation kernels (one thread): is there an eigenvalue of
// T (d, e) below kRefineRatio times the Gershgorin upper bound?  Same product-form recurrence
// and e^2 floor as sturm_count2.
// Matrices whose int8 digit planes had to round an element (an entry more than 2^17 below the largest one of its Gram
// row: gram_i8.cuh) carry a Gram error of up to K 2^-42 ||G|| instead of 2^-46: they are re-solved from W already
// when lambda_min / lambda_max < kRefineRatioInexact, which keeps delta sigma / sigma <= K 2^-43 / ratio ~ 2e-6.
constexpr double kRefineRatioInexact = 1e-5;
VSP_DEV bool has_tiny_eigenvalue(const double* d, const double* e, int n, double ratio = kRefineRatio) {
    double gu = 0.0;
    for (int i = 0; i < n; ++i) {
        const double el = (i > 0) ? fabs(e[i - 1]) : 0.0, er = (i < n - 1) ? fabs(e[i]) : 0.0;
        gu = fmax(gu, d[i] + el + er);
    }
    const double x = ratio * gu;
    double p0 = 1.0, p1 = d[0] - x;
    int changes = (p1 < 0.0) ? 1 : 0;
    for (int i = 1; i < n && changes == 0; ++i) {
        const double e2 = fmax(e[i - 1] * e[i - 1], kE2Floor);
        const double p2 = (d[i] - x) * p1 - e2 * p0;
        if ((p2 < 0.0) != (p1 < 0.0)) changes = 1;
        p0 = p1;
        p1 = p2;
        if ((i & 7) == 7) {
            const double s = rescale_factor(p0, p1);
            p0 *= s;
            p1 *= s;
        }
    }
    return changes != 0;
}

struct MetricOut {
    double metrics[4];
    int m, start, end, k, status;
};

// lam[] ascending, in scaled units (scale = power of four applied to the Gram
// matrix).  Writes sv[0..n) descending (if sv != nullptr) and returns the record
// fields.  `flags` carries VSP_ST_NONFINITE / VSP_ST_ZERO from the Gram stage.
// dist != nullptr: also the first dist_k entries of the distribution arrays of get_spectral_distribution
// (spectral.py:545-557), rows sv | sv^2 | sv / sv_0 | cumsum(sv^2) / sum(sv^2) of a [4][dist_k] block; NaN beyond n.
template <class Ctx>
VSP_DEV MetricOut spectral_metrics(Ctx& ctx, double* lam, int n, double scale, int flags,
                                   int fit_start, int fit_end, int hill_k, double* sv, double* dist = nullptr, int dist_k = 0,
                                   double* clauset = nullptr) {
    MetricOut out;
    const double nan = NAN;
    if (dist_k <= 0) dist = nullptr;
    if (dist)  // NaN everywhere first: the failure paths below and the entries beyond n keep it
        for (int i = ctx.tid; i < 4 * dist_k; i += ctx.nthreads) dist[i] = nan;
    if (clauset && ctx.tid < 8) clauset[ctx.tid] = (ctx.tid == 3 || ctx.tid == 4) ? -1.0 : ((ctx.tid < 3) ? nan : 0.0);
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
        if (dist) {  // all-zero matrix: s = 0, s_max -> 1, total variance 0 -> zeros (spectral.py:548-555)
            ctx.sync();
            for (int i = ctx.tid; i < 4 * dist_k; i += ctx.nthreads)
                if (i % dist_k < n) dist[i] = 0.0;
        }
        return out;
    }
    // Eigenvalues of a Gram matrix are >= 0; rounding can push the smallest ones to
    // <= 0.  LAPACK on W itself reports them as tiny positive numbers, so
    // floor at (eps * sigma_max)^2 instead of dropping them: m stays min(rows, cols).
    const double lfloor = lmax * 4.930380657631324e-32;  // 2^-104
    if (lam[0] <= kRefineRatio * lmax) out.status |= VSP_ST_ILLCOND;
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
    if (dist) {
        // entry i: sequential partial sum of the i + 1 largest eigenvalues, as np.cumsum forms it (dist_k is small:
        // SpectralTracker's max_singular_values; the full-length case costs n additions per thread once)
        const double unscale = sqrt(1.0 / scale), s0 = sqrt(lmax) * unscale;
        const int kk = dist_k < n ? dist_k : n;
        for (int i = ctx.tid; i < kk; i += ctx.nthreads) {
            double c = 0.0;
            for (int j = 0; j <= i; ++j) c += lam[n - 1 - j];
            const double s = sqrt(lam[n - 1 - i]) * unscale;
            dist[i] = s;
            dist[dist_k + i] = s * s;
            dist[2 * dist_k + i] = s / s0;
            dist[3 * dist_k + i] = c / total;
        }
    }

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

    // ---- Clauset-Shalizi-Newman x_min scan over the eigenvalue spectrum (opt-in; no counterpart in the reference).
    // Candidate cutoff k (0-based from the SMALLEST eigenvalue, k <= m - 2): tail = lam[k..m), t = m - k,
    // alpha_k = 1 + t / sum_{i >= k} ln(lam_i / lam_k), D_k = max_j max(|j/t - P(x_j)|, |(j-1)/t - P(x_j)|) over the
    // tail in ascending order, P(x) = 1 - (x / lam_k)^(1 - alpha_k).  All candidates in parallel (one per thread, cyclic),
    // the smallest D wins (ties: the smaller cutoff).  lam[] is overwritten by ln lam (nothing reads it afterwards).
    if (clauset && m >= 8) {
        ctx.sync();
        for (int i = ctx.tid; i < n; i += ctx.nthreads) lam[i] = log(lam[i]);
        ctx.sync();
        double bestD = 1e300, bestA = nan;
        int bestK = -1;
        for (int k = ctx.tid; k <= m - 2; k += ctx.nthreads) {
            const double lk = lam[k];
            const int t = m - k;
            double sl = 0.0;
            for (int i = k; i < m; ++i) sl += lam[i] - lk;
            if (!(sl > 0.0)) continue;  // a flat tail has no power-law fit
            const double a = 1.0 + (double)t / sl;
            const double e = 1.0 - a, it = 1.0 / (double)t;
            double D = 0.0;
            for (int j = 0; j < t; ++j) {
                const double P = 1.0 - exp(e * (lam[k + j] - lk));
                D = fmax(D, fmax(fabs((double)(j + 1) * it - P), fabs((double)j * it - P)));
            }
            if (D < bestD) {
                bestD = D;
                bestA = a;
                bestK = k;
            }
        }
        // arg-min over the threads: smallest D, then smallest cutoff
        const double gD = ctx.min(bestD);
        const int cand = (bestD == gD && bestK >= 0) ? bestK : 0x7fffffff;
        const int gK = -ctx.max_i(-cand);
        if (bestK == gK && gK != 0x7fffffff) {
            clauset[0] = bestA;
            clauset[1] = exp(lam[gK]) / scale;     // x_min in the units of sigma^2
            clauset[2] = gD;
            clauset[3] = (double)(m - 1 - gK);     // index in DESCENDING order (0 = largest eigenvalue)
            clauset[4] = (double)(m - gK);         // tail count
        }
    }
    return out;
}

}  // namespace vsp
