// chase8.cuh -- stage 2a-2 for the orders sbr8.cuh handles: symmetric band (bandwidth 8, compact rows
// Bd[r * 9 + j] = A[r][r - j] in frame indices) -> tridiagonal (d, e) by Householder bulge chasing.
//
// One CTA of kChase8Groups / 4 warps per matrix.  The working band (bandwidth grows to 15 while bulges are in flight) lives in shared memory
// as L[r][jj] = B[r][r - jj], jj = 0..15, row stride 17, with 14 zero rows of padding so that the blocks at the bottom
// of the matrix need no special cases (a reflector built from zeros is the identity).
//
// Sweep k (k = 0..n-3) annihilates column k below the sub-diagonal; its step j works on the rows R = [r0, r0 + 7],
// r0 = k + 1 + 8 j: it builds the reflector from the first column of the bulge (column k itself for j = 0) and applies
// it to
//     (a) the 8 x 8 block left of the diagonal block (rows R, columns r0-8 .. r0-1)   - from the left
//     (b) the 8 x 8 symmetric diagonal block (rows / columns R)                       - two-sided
//     (c) the 8 x 8 block below it (rows r0+8 .. r0+15, columns R): the next bulge    - from the right
// Step (k, j) touches rows r0 .. r0 + 15 only, and needs the steps (k, j-1) and (k-1, <= j+2): the CTA is split into
// kChase8Groups = 8 groups of eight lanes (a sweep of n = 192 has 24 steps, so eight sweeps fit behind one another),
// group g runs the sweeps g, g+8, g+16, ... and stays at least kChase8Lag = 3 steps behind the group that runs the
// previous sweep (3 steps = 24 rows: the row ranges of concurrent steps are disjoint, so the only ordering inside a
// tick is "a group's loads before its stores").  All groups advance in lock-step ticks (one CTA barrier per tick).
// Inside a step lane q owns column q of (a), row q of (b) and row q of (c); the only communication is the
// eight-lane exchange of p = tau D v.  n = 192: 546 ticks for 2 304 steps.
//
// Work: 6 n^2 b flops per matrix (1.8 MF at n = 192) -- latency-bound, not on the FP64 roofline.
#pragma once

#include "bisect_metrics.cuh"
#include "common.cuh"
#include "sbr8.cuh"

namespace vsp {

constexpr int kChase8W = 17;        // row stride: 16 stored diagonals + 1 pad
constexpr int kChase8PadRows = 14;  // a step reads the rows r0 .. r0 + 15 with r0 <= n - 2; 14 (not 3 b = 24) makes n = 192 fit EIGHT CTAs per SM
constexpr int kChase8Groups = 8;    // eight-lane groups per matrix (two warps)
constexpr int kChase8Threads = 8 * kChase8Groups;
constexpr int kChase8Lag = 3;

__host__ __device__ inline size_t chase8_smem_bytes(int n) {
    return sizeof(double) * ((size_t)(n + kChase8PadRows) * kChase8W + kChase8Groups);  // + progress table [2][groups] ints
}

// dot product of two 8-vectors as two chains of four (half the dependent latency of one chain, one more instruction)
__host__ __device__ inline double dot8(const double (&a)[8], const double (&b)[8]) {
    double s0 = a[0] * b[0], s1 = a[1] * b[1];
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        s0 = fma(a[i], b[i], s0);
        s1 = fma(a[i + 1], b[i + 1], s1);
    }
    return s0 + s1;
}

#if defined(__CUDACC__)

__global__ void __launch_bounds__(kChase8Threads)
    chase8_kernel(const ItemDesc* __restrict__ items, int item_base, int count, double* __restrict__ ws, RefineGate gate) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int idx = blockIdx.x;
    const ItemDesc it = items[item_base + idx];
    const int n = it.n;
    const int off = sbr8_order(n) - n;
    const int rows = n + kChase8PadRows;
    double* L = smem;
    // progress table, double-buffered by tick parity: [2][groups] words k | j << 16 | active << 24 -- a group reads its predecessor's
    // entry of the previous tick while everybody writes this tick's, so one barrier per tick is enough
    int* prog = reinterpret_cast<int*>(L + (size_t)rows * kChase8W);
    double* out = ws + it.de_off;
    const double* __restrict__ Bd = ws + it.gram_off + sbr8_band_off(n);

    if (out[2 * n + MISC_FLAGS] != 0.0) return;  // non-finite / all-zero: stage 2a-1 wrote d = e = 0 (uniform over the CTA)

    for (int i = tid; i < rows * kChase8W; i += kChase8Threads) {
        const int r = i / kChase8W, jj = i - r * kChase8W;
        double val = 0.0;
        if (r < n && jj <= 8 && r - jj >= 0) val = Bd[(off + r) * kBandW + jj];
        L[i] = val;
    }
    const int g = tid >> 3, q = tid & 7;
    const unsigned gmask = 0xffu << (lane & 24);
    // Sweeps shorter than 12 steps (k >= S) need only FOUR groups to keep the one-sweep-per-three-ticks rate (4 x lag = 12
    // ticks per round): they run on the groups of warp 0, and warp 1 sits the second half of the ticks out (23 % fewer
    // instructions; the kernel is throughput-bound on issue slots and the shared-memory pipe together).
    const int S = ((n > 97 ? n - 97 : 0) + 7) / 8 * 8;
    int k = g, j = 0;
    bool active = g <= n - 3 && (g < 4 || g < S);
    if (q == 0) prog[g] = k | (j << 16) | (active ? 1 << 24 : 0);
    int par = 0;
    while (__syncthreads_or(active)) {  // the barrier also orders last tick's stores and progress entries
        bool ready = false;
        if (active) {
            const int pg = (k > S) ? ((g + 3) & 3) : ((g + kChase8Groups - 1) % kChase8Groups);  // the group that runs sweep k - 1
            const int pv = prog[kChase8Groups * par + pg];
            const int kp = pv & 0xffff, jp = (pv >> 16) & 0xff, ap = pv >> 24;
            ready = (k == 0) || (ap == 0) || (kp > k - 1) || (kp == k - 1 && jp >= j + kChase8Lag);
        }
        par ^= 1;
        // Every lane runs the step body, converged: groups that are not ready work on row 0 and store nothing.  (With
        // the body under `if (ready)` the eight-lane exchange became a sub-warp shuffle in divergent code, which
        // the compiler lowers to a WARPSYNC.COLLECTIVE loop.)
        if (__any_sync(0xffffffffu, active)) {  // warp-uniform
            const int r0 = ready ? k + 1 + 8 * j : 0;
            const int xj = (j == 0) ? 1 : 8;  // jj of x_0: column r0-1 (first step) or r0-8
            double* Lr = L + (size_t)r0 * kChase8W;
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = Lr[i * kChase8W + xj + i];
            // (b) row q of the symmetric diagonal block
            double dq[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) dq[c] = (c <= q) ? Lr[q * kChase8W + q - c] : Lr[c * kChase8W + c - q];
            // (a) column q of the left block, (c) row q of the lower block
            // (columns left of the matrix, ca < 0, read band slots beyond the row start: zero and never written)
            const int ca = r0 - 8 + q;
            double a[8], cc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = Lr[i * kChase8W + 8 + i - q];
            double* Lc = Lr + (size_t)(8 + q) * kChase8W;  // row r0 + 8 + q
#pragma unroll
            for (int c = 0; c < 8; ++c) cc[c] = Lc[8 + q - c];
            const double xq = Lr[q * kChase8W + xj + q];  // x_q: this lane's entry of the reflector
            __syncwarp();  // a group's loads precede its stores (other groups work on other rows)
            double xn2;
            {
                double s0 = x[1] * x[1], s1 = x[2] * x[2];
                s0 = fma(x[3], x[3], s0);
                s1 = fma(x[4], x[4], s1);
                s0 = fma(x[5], x[5], s0);
                s1 = fma(x[6], x[6], s1);
                xn2 = fma(x[7], x[7], s0) + s1;
            }
            const bool doit = ready && xn2 > 0.0;
            double beta, tau, vs;
            const double s2 = fma(x[0], x[0], xn2);
            {
                const double rs = fast_rsqrt(s2);
                const double nrm = s2 * rs;
                beta = -copysign(nrm, x[0]);
                tau = fma(fabs(x[0]), rs, 1.0);
                vs = copysign(fast_rcp(fabs(x[0]) + nrm), x[0]);
            }
            if (doit && !(s2 > 1e-280)) {  // sub-normal range: exact division (never taken after the power-of-four scaling)
                beta = -copysign(sqrt(s2), x[0]);
                tau = (beta - x[0]) / beta;
                vs = 1.0 / (x[0] - beta);
            }
            double v[8];
            v[0] = 1.0;
#pragma unroll
            for (int i = 1; i < 8; ++i) v[i] = x[i] * vs;
            // (b): p = tau D v (lane q: entry q), exchanged inside the group
            const double pq = tau * dot8(dq, v);
            double p[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) p[c] = __shfl_sync(0xffffffffu, pq, c, 8);
            const double K = 0.5 * tau * dot8(v, p);
            const double vq = (q == 0) ? 1.0 : xq * vs;  // v_q, w_q of this lane without register indexing
            const double wq = fma(-K, vq, pq);
#pragma unroll
            for (int c = 0; c < 8; ++c) p[c] = fma(-K, v[c], p[c]);  // w
            if (doit) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c <= q) Lr[q * kChase8W + q - c] = dq[c] - fma(vq, p[c], wq * v[c]);
            }
            // (a)
            {
                const double ts = tau * dot8(a, v);
                const bool src = (8 - q == xj);  // the column the reflector was built from
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = src ? (i == 0 ? beta : 0.0) : fma(-ts, v[i], a[i]);
                if (doit && ca >= 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) Lr[i * kChase8W + 8 + i - q] = a[i];
                }
            }
            // (c)
            {
                const double ts = tau * dot8(cc, v);
                if (doit) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) Lc[8 + q - c] = fma(-ts, v[c], cc[c]);
                }
            }
        }
        if (ready) {
            ++j;
            if (k + 1 + 8 * j > n - 2) {
                k += (k >= S) ? 4 : kChase8Groups;  // (k < S: k + 8 = S + g when it crosses S)
                j = 0;
                if (k > n - 3 || (k >= S && g >= 4)) active = false;
            }
        }
        if (q == 0) {  // every tick, ready or not: the other buffer is one tick old
            prog[kChase8Groups * par + g] = k | (j << 16) | (active ? 1 << 24 : 0);
        }
    }
    if (tid >= 32) return;  // warp 0 publishes

    // d, e -> workspace (same slots the other reduction kernels fill) and, compacted, to the head of the band buffer
    // for the ill-conditioning gate (one sequential Sturm count by lane 0); in chunks of 8 x 32 entries: the
    // compacted copy overwrites band rows that have already been read
    for (int i0 = 0; i0 < n; i0 += 256) {
        double dv[8], ev[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int i = i0 + lane + 32 * t;
            dv[t] = (i < n) ? L[(size_t)i * kChase8W] : 0.0;
            ev[t] = (i < n - 1) ? L[(size_t)(i + 1) * kChase8W + 1] : 0.0;
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
