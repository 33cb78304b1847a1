// host_emul.cpp -- TEST INFRASTRUCTURE.  Compiles the per-matrix device algorithms
// (tridiag.cuh, bisect_metrics.cuh) with a host compiler, where the cooperative
// context degenerates to one thread, so tests/test_host_emul.py can check the
// product's indexing and numerics against the oracle without a GPU.
// Build: g++ -O2 -shared -fPIC -o _host_emul.so host_emul.cpp
#include <cstdlib>
#include <vector>

#include "../../vision-spectra_b200/csrc/bisect_metrics.cuh"
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
    const int iters = bisect_all(ctx, de.data(), n, t, lam.data());
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
