// tridiag_fused.cuh -- stage 2a, shared-memory path (n <= kSmemMaxN): Householder
// tridiagonalisation with the rank-2 update of step k fused into the symmetric
// mat-vec of step k+1, so every stored element of the lower triangle is read once and
// written once per step (the unfused form in tridiag.cuh reads it three times and
// writes it once).
//
// Storage: lower triangle, row r at poff(r) = tri(r) + (r+1)/2, i.e. rows padded to an
// even length so that every row starts 16-byte aligned and lanes can move two columns
// per LDS.128/STS.128.  The Gram kernel writes this layout directly.
//
// Work decomposition (one CTA per matrix, NW warps, two CTAs per SM):
//   * lane l owns the column pairs {2l, 2l+1} + 64 q, q < NP; the reflector entries
//     v_c, w_c, u_c and the column sums for those columns live in its registers.
//   * warp w owns the row pairs (2p, 2p+1), p = w, w+NW, ... (cyclic: the triangular row
//     lengths balance).  One row pair x one column pair is a 2x2 register block:
//         a -= v_r w_c + w_r v_c        (rank-2 update of the previous step)
//         rowacc_r += a u_c ;  colacc_c += a u_r  (c < r)      (this step's A u)
//     = 16 DFMA for 2 LDS.128 + 2 STS.128; the per-row bookkeeping is shared by two rows.
//     A row pair that spans QC chunks is straight-line code (pair_pass<NP,QC>), selected
//     by a warp-uniform switch.
//   * row sums of 8 rows are reduced with a folding butterfly (9 shuffle pairs per 8
//     rows); column sums are combined across warps through a [NW][n] scratch.
//   * the serial part of a step (p.u, w, next pivot row, its norm, the new reflector) is
//     done by warp 0 alone while the co-resident CTA keeps the SM busy: three block
//     barriers per step.
// FP64 work: 4 DFMA per stored element per step = (4/3) n^3 flops in total, the
// algorithmic count of the reduction.
//
// Same reflector convention and elimination order as tridiag.cuh (which remains the
// global-memory path for large n and the host-emulated statement of the algorithm).
#pragma once

#include <cstdio>

#include "bisect_metrics.cuh"
#include "common.cuh"

namespace vsp {

// offset of row r in the padded-even packed lower triangle
VSP_HD int poff(int r) { return ((r * (r + 1)) >> 1) + ((r + 1) >> 1); }

__host__ __device__ inline int fused_warps(int n, int rows_per_warp = 24) {
    int nw = (n + rows_per_warp - 1) / rows_per_warp;
    if (nw < 2) nw = 2;
    if (nw > 8) nw = 8;
    return nw;
}
constexpr int kFusedPad = 64;  // doubles after the triangle: masked lanes may read past a row pair
// Two CTAs must fit on one SM so that one matrix's per-step dependency chain (p.u -> w ->
// pivot row -> norm -> reflector) overlaps the other's pass: <= 113 KB of shared memory each.
constexpr size_t kFusedSmemBudget = 113 * 1024;
__host__ __device__ inline size_t tridiag_fused_smem_bytes(int rows_smem, int npad, int nw) {
    // v w u p prow diag d e (8 npad) | pcol [nw][pstride] (doubles as the fold buffer) | scalars (4) |
    // A (rows < rows_smem) + pad
    const size_t pstride = npad > 8 * 34 ? npad : 8 * 34;
    return sizeof(double) * ((size_t)poff(rows_smem) + kFusedPad + (size_t)8 * npad + (size_t)nw * pstride + 4);
}
// Rows [0, rows_smem) of the triangle live in shared memory; the rest (the rows eliminated
// first) stay in the global workspace and are updated in place through L2.  Even, so that a
// row pair never straddles the two address spaces.
__host__ __device__ inline int fused_rows_in_smem(int n, int npad, int nw) {
    int r = (n + 1) & ~1;
    while (r > 0 && tridiag_fused_smem_bytes(r, npad, nw) > kFusedSmemBudget) r -= 2;
    return r < n ? r : n;
}

#if defined(__CUDACC__)

__device__ __forceinline__ double shfl_xor_d(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// Reduce 8 per-lane row partials through a per-warp shared-memory transpose: every lane
// stores its 8 partials (row-slot major, stride 34 keeps both the stores and the transposed
// loads bank-conflict free), then lane l sums 8 of the 32 partials of row slot k = l & 7 and
// two butterfly steps finish the job.  ~30 instructions per 8 rows; the register-only folding
// butterfly needs ~95 because every step has to select which half to send.
// On return the lanes with the same (lane & 7) hold the full sum of row slot k = lane & 7.
constexpr int kFoldStride = 34;
constexpr int kFoldDoubles = 8 * kFoldStride;
__device__ __forceinline__ double fold8(const double (&rs)[8], int lane, double* __restrict__ buf) {
#pragma unroll
    for (int k = 0; k < 8; ++k) buf[k * kFoldStride + lane] = rs[k];
    __syncwarp();
    const double* src = buf + (lane & 7) * kFoldStride + (lane >> 3);
    double s = 0.0, t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        s += src[4 * i];
        t += src[4 * (i + 1)];
    }
    s += t;
    __syncwarp();  // the buffer is reused by the next group of rows
    s += shfl_xor_d(s, 8);
    s += shfl_xor_d(s, 16);
    return s;
}

// 1/x and 1/sqrt(x) for normal positive doubles: hardware seed + two Newton steps
// (|rel err| ~ 1e-16; the Householder scalars do not need correctly rounded division).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = fma(fma(-hx * r, r, 0.5), r, r);
    r = fma(fma(-hx * r, r, 0.5), r, r);
    return r;
}

// One row pair (r0 even, r1 = r0 + 1) of the fused pass.  The pair spans chunk pairs
// 0..QC-1; the last one holds both diagonals.  row1 == row0 + (r0 + 2).
template <int NP, int QC>
__device__ __forceinline__ void pair_pass(double* __restrict__ row0, int r0, int lane, double vr0, double wr0,
                                          double ur0, double vr1, double wr1, double ur1,
                                          const double2 (&vq)[NP], const double2 (&wq)[NP],
                                          const double2 (&uq)[NP], double2 (&colacc)[NP], double& rs0,
                                          double& rs1, double* __restrict__ diag) {
    double* __restrict__ row1 = row0 + (r0 + 2);
    const int r1 = r0 + 1;
    double2 a0[QC], a1[QC];
#pragma unroll
    for (int q = 0; q < QC; ++q) {
        a0[q] = *reinterpret_cast<const double2*>(row0 + 2 * lane + 64 * q);
        a1[q] = *reinterpret_cast<const double2*>(row1 + 2 * lane + 64 * q);
    }
#pragma unroll
    for (int q = 0; q < QC; ++q) {
        a0[q].x = fma(-wr0, vq[q].x, fma(-vr0, wq[q].x, a0[q].x));
        a0[q].y = fma(-wr0, vq[q].y, fma(-vr0, wq[q].y, a0[q].y));
        a1[q].x = fma(-wr1, vq[q].x, fma(-vr1, wq[q].x, a1[q].x));
        a1[q].y = fma(-wr1, vq[q].y, fma(-vr1, wq[q].y, a1[q].y));
    }
    const int c0 = 2 * lane + 64 * (QC - 1);  // first column of this lane in the last chunk
#pragma unroll
    for (int q = 0; q < QC - 1; ++q) {
        *reinterpret_cast<double2*>(row0 + 2 * lane + 64 * q) = a0[q];
        *reinterpret_cast<double2*>(row1 + 2 * lane + 64 * q) = a1[q];
    }
    // the slot after an even row's diagonal is padding, so a double2 store is always in-row
    if (c0 <= r0) *reinterpret_cast<double2*>(row0 + c0) = a0[QC - 1];
    if (c0 <= r1) *reinterpret_cast<double2*>(row1 + c0) = a1[QC - 1];

    double s0x = 0.0, s0y = 0.0, s1x = 0.0, s1y = 0.0;
#pragma unroll
    for (int q = 0; q < QC - 1; ++q) {
        s0x = fma(a0[q].x, uq[q].x, s0x);
        s0y = fma(a0[q].y, uq[q].y, s0y);
        s1x = fma(a1[q].x, uq[q].x, s1x);
        s1y = fma(a1[q].y, uq[q].y, s1y);
        colacc[q].x = fma(a1[q].x, ur1, fma(a0[q].x, ur0, colacc[q].x));
        colacc[q].y = fma(a1[q].y, ur1, fma(a0[q].y, ur0, colacc[q].y));
    }
    {
        constexpr int q = QC - 1;
        const int c1 = c0 + 1;
        // One mask per element (columns <= r).  The diagonal therefore enters both the row sum
        // and the column sum; the owner lane publishes it and phase (3) subtracts diag_c * u_c.
        // r0 is even and c0 is even, so the diagonals sit at (row0, c0 == r0).x and (row1, c0 == r0).y.
        const double r0x = (c0 <= r0) ? a0[q].x : 0.0, r0y = (c1 <= r0) ? a0[q].y : 0.0;
        const double r1x = (c0 <= r1) ? a1[q].x : 0.0, r1y = (c1 <= r1) ? a1[q].y : 0.0;
        if (c0 == r0) *reinterpret_cast<double2*>(diag + r0) = make_double2(a0[q].x, a1[q].y);
        s0x = fma(r0x, uq[q].x, s0x);
        s0y = fma(r0y, uq[q].y, s0y);
        s1x = fma(r1x, uq[q].x, s1x);
        s1y = fma(r1y, uq[q].y, s1y);
        colacc[q].x = fma(r1x, ur1, fma(r0x, ur0, colacc[q].x));
        colacc[q].y = fma(r1y, ur1, fma(r0y, ur0, colacc[q].y));
    }
    rs0 = s0x + s0y;
    rs1 = s1x + s1y;
}

template <int NP>
__device__ __forceinline__ void pair_dispatch(double* __restrict__ row0, int r0, int lane, double vr0, double wr0,
                                              double ur0, double vr1, double wr1, double ur1,
                                              const double2 (&vq)[NP], const double2 (&wq)[NP],
                                              const double2 (&uq)[NP], double2 (&colacc)[NP], double& rs0,
                                              double& rs1, double* __restrict__ diag) {
#define VSP_PAIR_ARGS row0, r0, lane, vr0, wr0, ur0, vr1, wr1, ur1, vq, wq, uq, colacc, rs0, rs1, diag
    switch ((r0 + 1) >> 6) {  // warp-uniform: chunk pair that holds the diagonals
        case 0: pair_pass<NP, 1>(VSP_PAIR_ARGS); return;
        case 1: if constexpr (NP >= 2) pair_pass<NP, 2>(VSP_PAIR_ARGS); return;
        case 2: if constexpr (NP >= 3) pair_pass<NP, 3>(VSP_PAIR_ARGS); return;
        case 3: if constexpr (NP >= 4) pair_pass<NP, 4>(VSP_PAIR_ARGS); return;
        default: return;
    }
#undef VSP_PAIR_ARGS
}

template <int NP>
__global__ void __launch_bounds__(256, 2)
    tridiag_fused_kernel(const ItemDesc* __restrict__ items, int item_base, double* __restrict__ ws, int npad,
                         int rows_smem, int debug_timing, RefineGate gate) {
    extern __shared__ __align__(16) double smem[];
    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, NW = nthreads >> 5;

    double* v = smem;           // previous reflector vector
    double* w = v + npad;       // previous w
    double* u = w + npad;       // current reflector vector
    double* p = u + npad;       // tau * A u
    double* prow = p + npad;    // lower-triangle row sums of A u
    double* diag = prow + npad;  // current diagonal, for the diagonal correction in (3)
    double* d = diag + npad;
    double* e = d + npad;
    // [NW][pstride] per-warp column sums, written at the end of the pass; during the pass the
    // same row is this warp's fold buffer
    const int pstride = npad > kFoldDoubles ? npad : kFoldDoubles;
    double* pcol = e + npad;
    double* scal = pcol + (size_t)NW * pstride;  // [0] tau of the current step
    double* A = scal + 4;                     // rows < rows_smem (+ kFusedPad)

    // ---- load + condition the Gram matrix: power-of-four scale so that |G_ij| <= 1 and the
    // singular values un-scale exactly; NaN/Inf anywhere in W shows on the Gram diagonal.
    double* __restrict__ G = ws + it.gram_off;
    double md = 0.0;
    int bad = 0;
    for (int c = lane; c < n; c += 32) {
        const double g = G[poff(c) + c];
        if (!isfinite(g)) bad = 1;
        md = fmax(md, g);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        md = fmax(md, shfl_xor_d(md, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    double* out = ws + it.de_off;
    int flags = 0;
    double scale = 1.0;
    if (bad) {
        flags = VSP_ST_NONFINITE;
    } else if (!(md > 0.0)) {
        flags = VSP_ST_ZERO;
    } else {
        int ex;
        (void)frexp(md, &ex);
        if (ex & 1) ex += 1;
        scale = ldexp(1.0, -ex);
    }
    if (flags) {  // uniform over the CTA: every warp derived the same flags
        for (int i = tid; i < 2 * n; i += nthreads) out[i] = 0.0;
        if (tid == 0) {
            out[2 * n + MISC_SCALE] = 1.0;
            out[2 * n + MISC_FLAGS] = (double)flags;
        }
        return;
    }
    const int total = poff(n), in_smem = poff(rows_smem);
    for (int i = tid; i < in_smem; i += nthreads) A[i] = G[i] * scale;
    for (int i = in_smem + tid; i < total; i += nthreads) G[i] *= scale;  // rows >= rows_smem: in place
    for (int i = tid; i < kFusedPad; i += nthreads) A[in_smem + i] = 0.0;
    for (int i = tid; i < npad; i += nthreads) {
        v[i] = 0.0;
        w[i] = 0.0;
        u[i] = 0.0;
        p[i] = 0.0;
    }
    __syncthreads();

    double2 vq[NP], wq[NP], uq[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) vq[q] = wq[q] = uq[q] = make_double2(0.0, 0.0);

#ifdef VSP_PHASE_TIMING  // per-phase cycle counters (development builds only)
    long long t_phase[4] = {0, 0, 0, 0};  // leader | pass | barrier+p | barrier   (this warp's view)
    long long t_mark = clock64();
#define VSP_LAP(k)                        \
    do {                                  \
        const long long now_ = clock64(); \
        t_phase[k] += now_ - t_mark;      \
        t_mark = now_;                    \
    } while (0)
#else
#define VSP_LAP(k) ((void)0)
#endif

    double tau_prev = 0.0;  // leader only
    for (int i = n - 1; i >= 1; --i) {
        const int m = i;  // leading block order; rows 0..m-1 remain after this step
        // ---- (L) warp 0: finish the previous step (w = p - (tau/2)(p.u) u, v = u), bring the
        //      pivot row i up to date and build the reflector that annihilates x[0..m-2]
        if (warp == 0) {
            if (i < n - 1) {
                double2 pq[NP];
                double dot = 0.0;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    pq[q] = make_double2(0.0, 0.0);
                    if (64 * q <= m) {  // columns < m + 1 (the previous block order)
                        pq[q] = *reinterpret_cast<const double2*>(p + 2 * lane + 64 * q);
                        dot = fma(pq[q].x, uq[q].x, fma(pq[q].y, uq[q].y, dot));
                    }
                }
                const double a2 = -0.5 * tau_prev * warp_sum(dot);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (64 * q <= m) {
                        wq[q].x = fma(a2, uq[q].x, pq[q].x);
                        wq[q].y = fma(a2, uq[q].y, pq[q].y);
                        vq[q] = uq[q];
                        *reinterpret_cast<double2*>(w + 2 * lane + 64 * q) = wq[q];
                        *reinterpret_cast<double2*>(v + 2 * lane + 64 * q) = vq[q];
                    }
                }
                __syncwarp();
            }
            const int offi = poff(i);
            const bool i_smem = i < rows_smem;
            const double vi = v[i], wi = w[i];
            double ss = 0.0;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                uq[q] = make_double2(0.0, 0.0);
                if (64 * q < m) {  // warp-uniform: later chunks are already eliminated
                    const int c0 = 2 * lane + 64 * q;
                    double2 x = i_smem ? *reinterpret_cast<const double2*>(A + offi + c0)
                                       : *reinterpret_cast<const double2*>(G + offi + c0);
                    x.x = fma(-wi, vq[q].x, fma(-vi, wq[q].x, x.x));
                    x.y = fma(-wi, vq[q].y, fma(-vi, wq[q].y, x.y));
                    x.x = (c0 < m) ? x.x : 0.0;
                    x.y = (c0 + 1 < m) ? x.y : 0.0;
                    uq[q] = x;
                    if (c0 < m - 1) ss = fma(x.x, x.x, ss);
                    if (c0 + 1 < m - 1) ss = fma(x.y, x.y, ss);
                }
            }
            const double xnorm2 = warp_sum(ss);
            const double rim1 = i_smem ? A[offi + m - 1] : G[offi + m - 1];
            const double rii = i_smem ? A[offi + i] : G[offi + i];
            const double alpha = fma(-wi, v[m - 1], fma(-vi, w[m - 1], rim1));
            double beta = alpha, tau = 0.0, vscale = 0.0;
            if (xnorm2 > 0.0) {
                // |x| <= n after the power-of-four scaling, so s2 is a normal number unless the
                // whole row is ~1e-140, in which case the IEEE path is taken.
                const double s2 = fma(alpha, alpha, xnorm2);
                if (s2 > 1e-280) {
                    const double rs = fast_rsqrt(s2);  // 1/||x||
                    const double nrm = s2 * rs;        // ||x||
                    beta = -copysign(nrm, alpha);
                    tau = fma(fabs(alpha), rs, 1.0);  // (beta - alpha)/beta = 1 + |alpha|/||x||
                    vscale = copysign(fast_rcp(fabs(alpha) + nrm), alpha);  // 1/(alpha - beta)
                } else {
                    beta = -copysign(sqrt(s2), alpha);
                    tau = (beta - alpha) / beta;
                    vscale = 1.0 / (alpha - beta);
                }
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (64 * q < m) {
                    const int c0 = 2 * lane + 64 * q;
                    uq[q].x = (tau == 0.0) ? 0.0 : ((c0 == m - 1) ? 1.0 : uq[q].x * vscale);
                    uq[q].y = (tau == 0.0) ? 0.0 : ((c0 + 1 == m - 1) ? 1.0 : uq[q].y * vscale);
                    *reinterpret_cast<double2*>(u + c0) = uq[q];  // zero for columns >= m
                }
            }
            if (lane == 0) {
                scal[0] = tau;
                e[i - 1] = beta;
                d[i] = fma(-2.0 * vi, wi, rii);
            }
            tau_prev = tau;
        }
        VSP_LAP(0);
        __syncthreads();
        const double tau = scal[0];
        if (warp != 0) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (64 * q <= m) {
                    vq[q] = *reinterpret_cast<const double2*>(v + 2 * lane + 64 * q);
                    wq[q] = *reinterpret_cast<const double2*>(w + 2 * lane + 64 * q);
                    uq[q] = *reinterpret_cast<const double2*>(u + 2 * lane + 64 * q);
                }
            }
        }

        // ---- (2) fused pass over the row pairs (2p, 2p+1) with 2p < m
        double2 colacc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) colacc[q] = make_double2(0.0, 0.0);
        for (int p0 = warp; 2 * p0 < m; p0 += NW * 4) {
            double rs[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r0 = 2 * (p0 + j * NW);
                rs[2 * j] = 0.0;
                rs[2 * j + 1] = 0.0;
                if (r0 < m) {  // warp-uniform
                    const bool has1 = r0 + 1 < m;  // the odd row may be the pivot row (already dead)
                    const double vr0 = v[r0], wr0 = w[r0], ur0 = u[r0];
                    const double vr1 = has1 ? v[r0 + 1] : 0.0, wr1 = has1 ? w[r0 + 1] : 0.0,
                                 ur1 = has1 ? u[r0 + 1] : 0.0;
                    const int off = poff(r0);
                    if (r0 < rows_smem)
                        pair_dispatch<NP>(A + off, r0, lane, vr0, wr0, ur0, vr1, wr1, ur1, vq, wq, uq, colacc,
                                          rs[2 * j], rs[2 * j + 1], diag);
                    else
                        pair_dispatch<NP>(G + off, r0, lane, vr0, wr0, ur0, vr1, wr1, ur1, vq, wq, uq, colacc,
                                          rs[2 * j], rs[2 * j + 1], diag);
                }
            }
            const double rsum = fold8(rs, lane, pcol + warp * pstride);
            if (lane < 8) {  // lane = row slot
                const int r = 2 * (p0 + (lane >> 1) * NW) + (lane & 1);
                if (r < m) prow[r] = rsum;
            }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (64 * q < m) *reinterpret_cast<double2*>(pcol + warp * pstride + 2 * lane + 64 * q) = colacc[q];
        VSP_LAP(1);
        __syncthreads();

        // ---- (3) p = tau (A u), one column per thread
        for (int c = tid; c < m; c += nthreads) {
            double s0 = prow[c], s1 = -diag[c] * u[c], s2 = 0.0, s3 = 0.0;  // the diagonal counted twice
            int k = 0;
            for (; k + 3 < NW; k += 4) {
                s0 += pcol[k * pstride + c];
                s1 += pcol[(k + 1) * pstride + c];
                s2 += pcol[(k + 2) * pstride + c];
                s3 += pcol[(k + 3) * pstride + c];
            }
            for (; k < NW; ++k) s0 += pcol[k * pstride + c];
            p[c] = tau * ((s0 + s1) + (s2 + s3));
        }
        VSP_LAP(2);
        __syncthreads();
        VSP_LAP(3);
    }
#ifdef VSP_PHASE_TIMING
    if (debug_timing && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == NW - 1))
        printf("[tridiag_fused n=%d warp %d] cycles: leader %lld  pass %lld  barrier+p %lld  barrier %lld\n", n, warp,
               t_phase[0], t_phase[1], t_phase[2], t_phase[3]);
#endif
    (void)debug_timing;
#undef VSP_LAP
    // The last pass (m = 1) applied the final update to element (0,0).
    if (tid == 0) {
        d[0] = (rows_smem > 0) ? A[0] : G[0];
        e[n - 1] = 0.0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthreads) {
        out[i] = d[i];
        out[n + i] = e[i];
    }
    if (tid == 0) {
        int oflags = 0, slot = -1;
        if (gate.counter != nullptr && has_tiny_eigenvalue(d, e, n)) {  // kappa >~ 3e4: re-solve from W
            slot = atomicAdd(gate.counter, 1);
            if (slot < gate.slots) {
                oflags = VSP_ST_ILLCOND;
                gate.slot_items[slot] = item_base + blockIdx.x;
            }
        }
        out[2 * n + MISC_SCALE] = scale;
        out[2 * n + MISC_FLAGS] = (double)oflags;
        out[2 * n + MISC_SLOT] = (double)slot;
    }
}

#endif  // __CUDACC__

}  // namespace vsp
