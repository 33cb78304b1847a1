// band_tridiag.cuh -- stage 2a-2: symmetric band (bandwidth kSbrB = 4, produced in place by
// sbr_band.cuh) -> tridiagonal (d, e) by Householder bulge chasing.
//
// One warp per matrix.  The working band (bandwidth grows to 2b-1 while bulges are in flight)
// lives in shared memory as L[r][jj] = B[r][r - jj], jj = 0..7, with 3b zero rows of padding so
// that the blocks at the bottom of the matrix need no special cases (a reflector built from
// zeros is the identity).
//
// Sweep k (k = 0..n-3) annihilates column k below the sub-diagonal; its step j works on the rows
// R = [r0, r0+3], r0 = k + 1 + 4j: it builds the reflector from the first column of the bulge
// (column k itself for j = 0) and applies it to
//     (a) the 4x4 block left of the diagonal block (rows R, columns r0-4..r0-1)     - from the left
//     (b) the 4x4 symmetric diagonal block (rows/columns R)                          - two-sided
//     (c) the 4x4 block below it (rows r0+4..r0+7, columns R): the next bulge        - from the right
// Step (k, j) only depends on steps (k, j-1) and (k-1, <= j+3), so the sweeps are pipelined:
// the warp is split into 8 groups of 4 lanes, group g runs the sweeps g, g+8, g+16, ... and stays
// at least 4 steps behind the group that runs the previous sweep.  All groups advance in
// lock-step "ticks" (one __syncwarp per tick); within a step the four lanes of a group work
// without communication: every lane rebuilds the reflector from the same four numbers, lane q
// then owns column q of (a), row q of (b) and row q of (c).
// n = 192: 808 ticks instead of 4 560 sequential steps.
//
// Work: 6 n^2 b flops per matrix (0.9 MF at n = 192) -- latency-bound, not on the FP64 roofline.
#pragma once

#include "bisect_metrics.cuh"
#include "common.cuh"

namespace vsp {

constexpr int kSbrB = 4;        // bandwidth after stage 2a-1
constexpr int kChaseW = 9;      // row stride: 2b stored diagonals + 1 pad (odd stride spreads the rows over the banks)
constexpr int kChasePadRows = 3 * kSbrB;
constexpr int kChaseWarps = 3;  // matrices per CTA

__host__ __device__ inline size_t band_tridiag_smem_bytes(int n) {
    return sizeof(double) * (size_t)kChaseWarps * ((size_t)(n + kChasePadRows) * kChaseW + 32);
}

#if defined(__CUDACC__)

__global__ void __launch_bounds__(32 * kChaseWarps)
    band_tridiag_kernel(const ItemDesc* __restrict__ items, int item_base, int count, double* __restrict__ ws,
                        RefineGate gate) {
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * kChaseWarps + warp;
    if (idx >= count) return;
    const ItemDesc it = items[item_base + idx];
    const int n = it.n;
    const int rows = n + kChasePadRows;
    double* L = smem + (size_t)warp * ((size_t)rows * kChaseW + 32);
    int* prog = reinterpret_cast<int*>(L + (size_t)rows * kChaseW);  // [8 groups][k, j, active, pad]
    double* out = ws + it.de_off;
    const double* __restrict__ G = ws + it.gram_off;

    if (out[2 * n + MISC_FLAGS] != 0.0) return;  // non-finite / all-zero: stage 2a-1 wrote d = e = 0

    for (int i = lane; i < rows * kChaseW; i += 32) {
        const int r = i / kChaseW, jj = i - r * kChaseW;
        double val = 0.0;
        if (r < n && jj <= kSbrB && r - jj >= 0) val = G[poff(r) + r - jj];
        L[i] = val;
    }
    const int g = lane >> 2, q = lane & 3;
    int k = g, j = 0;
    bool active = g <= n - 3;
    if (q == 0) {
        prog[4 * g + 0] = k;
        prog[4 * g + 1] = j;
        prog[4 * g + 2] = active ? 1 : 0;
    }
    __syncwarp();

    while (__any_sync(0xffffffffu, active)) {
        bool ready = false;
        if (active) {
            const int pg = (g + 7) & 7;
            const int kp = prog[4 * pg + 0], jp = prog[4 * pg + 1], ap = prog[4 * pg + 2];
            ready = (k == 0) || (ap == 0) || (kp > k - 1) || (kp == k - 1 && jp >= j + 4);
        }
        const unsigned rmask = __ballot_sync(0xffffffffu, ready);  // also: everyone has read the progress table
        if (ready) {
            const int r0 = k + 1 + 4 * j;
            const int xj = (j == 0) ? 1 : 4;  // jj of x_0: column r0-1 (first step) or r0-4
            double* Lr = L + (size_t)r0 * kChaseW;
            const double x0 = Lr[xj], x1 = Lr[kChaseW + xj + 1], x2 = Lr[2 * kChaseW + xj + 2],
                         x3 = Lr[3 * kChaseW + xj + 3];
            const double xn2 = fma(x1, x1, fma(x2, x2, x3 * x3));
            // ---- loads of the three blocks: (b) whole diagonal block (every lane), (a) column q of the
            //      left block, (c) row q of the lower block
            const double d00 = Lr[0];
            const double d10 = Lr[kChaseW + 1], d11 = Lr[kChaseW];
            const double d20 = Lr[2 * kChaseW + 2], d21 = Lr[2 * kChaseW + 1], d22 = Lr[2 * kChaseW];
            const double d30 = Lr[3 * kChaseW + 3], d31 = Lr[3 * kChaseW + 2], d32 = Lr[3 * kChaseW + 1],
                         d33 = Lr[3 * kChaseW];
            const int ca = r0 - 4 + q;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (ca >= 0) {
                a0 = Lr[4 - q];
                a1 = Lr[kChaseW + 5 - q];
                a2 = Lr[2 * kChaseW + 6 - q];
                a3 = Lr[3 * kChaseW + 7 - q];
            }
            double* Lc = Lr + (size_t)(4 + q) * kChaseW;  // row r0 + 4 + q
            const double c0 = Lc[4 + q], c1 = Lc[3 + q], c2 = Lc[2 + q], c3 = Lc[1 + q];
            __syncwarp(rmask);  // all loads of this tick precede all of its stores
            if (xn2 > 0.0) {
                double beta, tau, vs;
                const double s2 = fma(x0, x0, xn2);
                if (s2 > 1e-280) {
                    const double rs = fast_rsqrt(s2);
                    const double nrm = s2 * rs;
                    beta = -copysign(nrm, x0);
                    tau = fma(fabs(x0), rs, 1.0);
                    vs = copysign(fast_rcp(fabs(x0) + nrm), x0);
                } else {
                    beta = -copysign(sqrt(s2), x0);
                    tau = (beta - x0) / beta;
                    vs = 1.0 / (x0 - beta);
                }
                const double v1 = x1 * vs, v2 = x2 * vs, v3 = x3 * vs;
                const double p0 = tau * fma(d30, v3, fma(d20, v2, fma(d10, v1, d00)));
                const double p1 = tau * fma(d31, v3, fma(d21, v2, fma(d11, v1, d10)));
                const double p2 = tau * fma(d32, v3, fma(d22, v2, fma(d21, v1, d20)));
                const double p3 = tau * fma(d33, v3, fma(d32, v2, fma(d31, v1, d30)));
                const double K = 0.5 * tau * fma(v3, p3, fma(v2, p2, fma(v1, p1, p0)));
                const double w0 = p0 - K, w1 = fma(-K, v1, p1), w2 = fma(-K, v2, p2), w3 = fma(-K, v3, p3);
                if (q == 0) {
                    Lr[0] = d00 - 2.0 * w0;
                } else if (q == 1) {
                    Lr[kChaseW + 1] = d10 - fma(v1, w0, w1);
                    Lr[kChaseW] = d11 - 2.0 * v1 * w1;
                } else if (q == 2) {
                    Lr[2 * kChaseW + 2] = d20 - fma(v2, w0, w2);
                    Lr[2 * kChaseW + 1] = d21 - fma(v2, w1, w2 * v1);
                    Lr[2 * kChaseW] = d22 - 2.0 * v2 * w2;
                } else {
                    Lr[3 * kChaseW + 3] = d30 - fma(v3, w0, w3);
                    Lr[3 * kChaseW + 2] = d31 - fma(v3, w1, w3 * v1);
                    Lr[3 * kChaseW + 1] = d32 - fma(v3, w2, w3 * v2);
                    Lr[3 * kChaseW] = d33 - 2.0 * v3 * w3;
                }
                // (a)
                if (ca >= 0) {
                    if (4 - q == xj) {  // the column the reflector was built from
                        a0 = beta;
                        a1 = a2 = a3 = 0.0;
                    } else {
                        const double ts = tau * fma(a3, v3, fma(a2, v2, fma(a1, v1, a0)));
                        a0 -= ts;
                        a1 = fma(-ts, v1, a1);
                        a2 = fma(-ts, v2, a2);
                        a3 = fma(-ts, v3, a3);
                    }
                    Lr[4 - q] = a0;
                    Lr[kChaseW + 5 - q] = a1;
                    Lr[2 * kChaseW + 6 - q] = a2;
                    Lr[3 * kChaseW + 7 - q] = a3;
                }
                // (c)
                {
                    const double ts = tau * fma(c3, v3, fma(c2, v2, fma(c1, v1, c0)));
                    Lc[4 + q] = c0 - ts;
                    Lc[3 + q] = fma(-ts, v1, c1);
                    Lc[2 + q] = fma(-ts, v2, c2);
                    Lc[1 + q] = fma(-ts, v3, c3);
                }
            }
            ++j;
            if (k + 1 + 4 * j > n - 2) {
                k += 8;
                j = 0;
                if (k > n - 3) active = false;
            }
            if (q == 0) {
                prog[4 * g + 0] = k;
                prog[4 * g + 1] = j;
                prog[4 * g + 2] = active ? 1 : 0;
            }
        }
        __syncwarp();
    }

    // d, e -> workspace (same slots the single-stage kernels fill) and, compacted, to the head of
    // the band buffer for the ill-conditioning gate (one sequential Sturm count by lane 0)
    // (in chunks of 8 x 32 entries: the compacted copy overwrites band rows that have already been read)
    for (int i0 = 0; i0 < n; i0 += 256) {
        double dv[8], ev[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = i0 + lane + 32 * t;
            dv[t] = (i < n) ? L[(size_t)i * kChaseW] : 0.0;
            ev[t] = (i < n - 1) ? L[(size_t)(i + 1) * kChaseW + 1] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = i0 + lane + 32 * t;
            if (i < n) {
                out[i] = dv[t];
                out[n + i] = ev[t];
            }
        }
    }
    __syncwarp();
    // compact d | e at the head of the (now dead) band buffer, re-read from `out`
    for (int i = lane; i < 2 * n; i += 32) L[i] = out[i];
    __syncwarp();
    if (lane == 0) {
        int oflags = 0, slot = -1;
        const bool rounded = gate.inexact != nullptr && gate.inexact[item_base + idx] != 0;
        if (gate.counter != nullptr && has_tiny_eigenvalue(L, L + n, n, rounded ? kRefineRatioInexact : kRefineRatio)) {  // re-solve from W
            slot = atomicAdd(gate.counter, 1);  // list entry (the list holds every item of the class)
            oflags = VSP_ST_ILLCOND;
            gate.slot_items[slot] = item_base + idx;
        }
        out[2 * n + MISC_FLAGS] = (double)oflags;
        out[2 * n + MISC_SLOT] = (double)slot;
    }
}

#endif  // __CUDACC__

}  // namespace vsp
