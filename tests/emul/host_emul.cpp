// host_emul.cpp -- TEST INFRASTRUCTURE.  Compiles the per-matrix device algorithms
// (tridiag.cuh, bisect_metrics.cuh) with a host compiler, where the cooperative
// context degenerates to one thread, so tests/test_host_emul.py can check the
// product's indexing and numerics against the oracle without a GPU.
// Build: g++ -O2 -shared -fPIC -o _host_emul.so host_emul.cpp
#include <cstdlib>
#include <vector>

#include "../../vision-spectra_b200/csrc/bisect_metrics.cuh"
#include "../../vision-spectra_b200/csrc/refine_bidiag.cuh"
#include "../../vision-spectra_b200/csrc/tridiag.cuh"

using namespace vsp;

extern "C" int vsp_emul_eig_metrics(const double* gram, int n, int use_full, int split, int fit_start,
                                    int fit_end, int hill_k, double* sv, double* metrics4, int* ints6) {
    HostCtx ctx;
    const int npad = (n + 31) / 32 * 32;
    std::vector<double> a(use_full ? (size_t)n * n : (size_t)tri(n));
    double maxdiag = 0.0;
    int nonfinite = 0;
    for (int i = 0; i < n; ++i) {
        const double g = gram[(size_t)i * n + i];
        if (!std::isfinite(g)) nonfinite = 1;
        if (g > maxdiag) maxdiag = g;
    }
    int flags = 0;
    const double scale = gram_scale(ctx, maxdiag, nonfinite, &flags);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            const double g = flags ? 0.0 : gram[(size_t)r * n + c] * scale;
            if (use_full)
                a[(size_t)c * n + r] = g;
            else if (c <= r)
                a[tri(r) + c] = g;
        }
    std::vector<double> d(n), e(n), v(n), p(n), part((size_t)split * npad), lam(n);
    if (use_full)
        tridiagonalize(ctx, FullSym{a.data(), n}, n, npad, split, d.data(), e.data(), v.data(), p.data(), part.data());
    else
        tridiagonalize(ctx, PackedLower{a.data(), n}, n, npad, split, d.data(), e.data(), v.data(), p.data(), part.data());
    TriInfo t = tri_bounds(ctx, d.data(), e.data(), n);
    std::vector<DE> de(n);
    for (int i = 0; i < n; ++i) {
        de[i].d = d[i];
        de[i].e2 = i > 0 ? std::fmax(e[i - 1] * e[i - 1], kE2Floor) : 0.0;
    }
    int next_k = 0;
    std::vector<double> gridbuf(CoarseGrid::doubles(1));
    const int iters = bisect_all(ctx, de.data(), n, t, lam.data(), &next_k, gridbuf.data());
    MetricOut out = spectral_metrics(ctx, lam.data(), n, scale, flags, fit_start, fit_end, hill_k, sv);
    for (int q = 0; q < 4; ++q) metrics4[q] = out.metrics[q];
    ints6[0] = out.m;
    ints6[1] = out.start;
    ints6[2] = out.end;
    ints6[3] = out.k;
    ints6[4] = out.status;
    ints6[5] = iters;
    return 0;
}

// W: row-major rows x cols (f64).  Bidiagonalise the K x n column-major copy (Gram index on the
// columns) and return the singular values (descending) and the metrics computed from them.
extern "C" int vsp_emul_refine(const double* w, int rows, int cols, double* sv, double* metrics4, int* ints6) {
    HostCtx ctx;
    const int n = rows < cols ? rows : cols, K = rows < cols ? cols : rows;
    std::vector<double> X((size_t)K * n);
    double mx = 0.0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) mx = std::fmax(mx, std::fabs(w[(size_t)r * cols + c]));
    int ex = 0;
    if (mx > 0.0) (void)std::frexp(mx, &ex);
    const double sc = std::ldexp(1.0, -ex);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const double x = w[(size_t)r * cols + c] * sc;
            if (rows <= cols) X[(size_t)r * K + c] = x;   // column r of X = row r of W
            else X[(size_t)c * K + r] = x;                // column c of X = column c of W
        }
    std::vector<double> dq(n), eq(n), u(n), lam(n), part(4);
    std::vector<DE> de(2 * n);
    bidiagonalize(ctx, X.data(), K, n, dq.data(), eq.data(), u.data(), part.data());
    const int iters = gk_singular_values(ctx, dq.data(), eq.data(), n, de.data(), lam.data());
    // scale of the Gram-route convention: lam are eigenvalues of (sc W)^T (sc W) -> scale = sc^2
    MetricOut out = spectral_metrics(ctx, lam.data(), n, sc * sc, 0, -1, -1, -1, sv);
    for (int q = 0; q < 4; ++q) metrics4[q] = out.metrics[q];
    ints6[0] = out.m; ints6[1] = out.start; ints6[2] = out.end; ints6[3] = out.k; ints6[4] = out.status; ints6[5] = iters;
    return 0;
}
