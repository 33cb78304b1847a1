// tridiag_fused.cuh -- stage 2a, shared-memory path (n <= kSmemMaxN): Householder
// tridiagonalisation with the rank-2 update of step k fused into the symmetric
// mat-vec of step k+1, so every stored element of the packed lower triangle is read
// once and written once per step (the unfused form in tridiag.cuh reads it three
// times and writes it once).
//
// Work decomposition (one CTA per matrix, NW warps):
//   * warp w owns rows r = w, w+NW, w+2NW, ... (cyclic, so the triangular row lengths
//     balance); it walks each of its rows with lanes on consecutive columns
//     c = lane + 32 q  -> 8-byte accesses to tri(r)+c are contiguous: conflict free.
//   * per element a(r,c):   a -= v_r w_c + w_r v_c        (update of the previous step)
//                           rowacc_r += a u_c             (lower-triangle part of A u)
//                           colacc_c += a u_r  (c < r)    (its transpose part)
//     v_c, w_c, u_c and colacc_c live in registers of the lane that owns column c;
//     v_r, w_r, u_r are warp-uniform shared-memory broadcasts.
//   * a row that spans QC 32-column chunks is handled by straight-line code
//     (row_pass<QC>): all loads first, then the FMAs, so the FP64 pipe sees QC
//     independent chains; the chunk count is a warp-uniform switch.
//   * row sums of 8 rows at a time are reduced with a folding butterfly (9 shuffle
//     pairs for 8 rows instead of 40); column sums are combined across warps through
//     a [NW][n] scratch.
//   * the reflector of the next step (norm, beta, tau, u) and w = p - (tau/2)(p.u)u
//     are computed by EVERY warp redundantly from its own lanes (each warp's lanes
//     cover all columns), so they need warp shuffles only: two block barriers per
//     step in total (after the pass, after p).
// FP64 work: 4 DFMA per stored element per step = (4/3) n^3 flops in total, the
// algorithmic count of the reduction.  Shared-memory traffic: 16 B per element per
// step.
//
// Same reflector convention and elimination order as tridiag.cuh (which remains the
// global-memory path for large n and the host-emulated statement of the algorithm).
#pragma once

#include "common.cuh"

namespace vsp {

__device__ __forceinline__ double shfl_xor_d(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// rs[0..7] hold one partial per row and lane.  On return rs[0] is the full sum of row
// j = 4*bit4(lane) + 2*bit3(lane) + bit2(lane), identical in the 4 lanes that share j.
__device__ __forceinline__ void fold8(double (&rs)[8], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double send = hi ? rs[k] : rs[k + 4];
            const double keep = hi ? rs[k + 4] : rs[k];
            rs[k] = keep + shfl_xor_d(send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double send = hi ? rs[k] : rs[k + 2];
            const double keep = hi ? rs[k + 2] : rs[k];
            rs[k] = keep + shfl_xor_d(send, 8);
        }
    }
    {
        const bool hi = lane & 4;
        const double send = hi ? rs[0] : rs[1];
        const double keep = hi ? rs[1] : rs[0];
        rs[0] = keep + shfl_xor_d(send, 4);
    }
    rs[0] += shfl_xor_d(rs[0], 2);
    rs[0] += shfl_xor_d(rs[0], 1);
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

__host__ __device__ inline int fused_warps(int n, int rows_per_warp = 24) {
    int nw = (n + rows_per_warp - 1) / rows_per_warp;
    if (nw < 2) nw = 2;
    if (nw > 8) nw = 8;
    return nw;
}
constexpr int kFusedPad = 32;  // doubles after the packed triangle: masked lanes may read past a row
// Two CTAs must fit on one SM so that one matrix's per-step dependency chain (norm ->
// reflector -> p -> w) overlaps the other's pass: <= 113 KB of shared memory each.
constexpr size_t kFusedSmemBudget = 113 * 1024;
__host__ __device__ inline size_t tridiag_fused_smem_bytes(int rows_smem, int npad, int nw) {
    // v w u p prow d e (7 npad) | pcol [nw][npad] | A packed (rows < rows_smem) + pad
    return sizeof(double) * ((size_t)tri(rows_smem) + kFusedPad + (size_t)(7 + nw) * npad);
}
// Rows [0, rows_smem) of the packed triangle live in shared memory; the rest (the rows
// eliminated first) stay in the global workspace and are updated in place through L2.
__host__ __device__ inline int fused_rows_in_smem(int n, int npad, int nw) {
    int r = n;
    while (r > 0 && tridiag_fused_smem_bytes(r, npad, nw) > kFusedSmemBudget) --r;
    return r;
}

// One row of the fused pass; the row spans chunks 0..QC-1, the last one holds the diagonal.
template <int NQ, int QC>
__device__ __forceinline__ double row_pass(double* __restrict__ row, int r, int lane, double vr, double wr,
                                           double ur, const double (&vq)[NQ], const double (&wq)[NQ],
                                           const double (&uq)[NQ], double (&colacc)[NQ]) {
    double a[QC];
#pragma unroll
    for (int q = 0; q < QC; ++q) a[q] = row[lane + 32 * q];
#pragma unroll
    for (int q = 0; q < QC; ++q) {
        a[q] = fma(-vr, wq[q], a[q]);
        a[q] = fma(-wr, vq[q], a[q]);
    }
#pragma unroll
    for (int q = 0; q < QC - 1; ++q) row[lane + 32 * q] = a[q];
    const int cl = lane + 32 * (QC - 1);
    if (cl <= r) row[cl] = a[QC - 1];
    double rs = 0.0;
#pragma unroll
    for (int q = 0; q < QC - 1; ++q) {
        rs = fma(a[q], uq[q], rs);
        colacc[q] = fma(a[q], ur, colacc[q]);
    }
    const double al = (cl <= r) ? a[QC - 1] : 0.0;
    const double ac = (cl < r) ? a[QC - 1] : 0.0;
    rs = fma(al, uq[QC - 1], rs);
    colacc[QC - 1] = fma(ac, ur, colacc[QC - 1]);
    return rs;
}

template <int NQ>
__device__ __forceinline__ double row_dispatch(double* __restrict__ row, int r, int lane, double vr, double wr,
                                               double ur, const double (&vq)[NQ], const double (&wq)[NQ],
                                               const double (&uq)[NQ], double (&colacc)[NQ]) {
    switch (r >> 5) {  // warp-uniform
        case 0: return row_pass<NQ, 1>(row, r, lane, vr, wr, ur, vq, wq, uq, colacc);
        case 1: if constexpr (NQ >= 2) return row_pass<NQ, 2>(row, r, lane, vr, wr, ur, vq, wq, uq, colacc); break;
        case 2: if constexpr (NQ >= 3) return row_pass<NQ, 3>(row, r, lane, vr, wr, ur, vq, wq, uq, colacc); break;
        case 3: if constexpr (NQ >= 4) return row_pass<NQ, 4>(row, r, lane, vr, wr, ur, vq, wq, uq, colacc); break;
        case 4: if constexpr (NQ >= 5) return row_pass<NQ, 5>(row, r, lane, vr, wr, ur, vq, wq, uq, colacc); break;
        case 5: if constexpr (NQ >= 6) return row_pass<NQ, 6>(row, r, lane, vr, wr, ur, vq, wq, uq, colacc); break;
        case 6: if constexpr (NQ >= 7) return row_pass<NQ, 7>(row, r, lane, vr, wr, ur, vq, wq, uq, colacc); break;
        default: break;
    }
    return 0.0;
}

template <int NQ>
__global__ void __launch_bounds__(256, 2)
    tridiag_fused_kernel(const ItemDesc* __restrict__ items, int item_base, double* __restrict__ ws, int npad,
                         int rows_smem) {
    extern __shared__ __align__(16) double smem[];
    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, NW = nthreads >> 5;

    double* v = smem;           // previous reflector vector          (identical writes by every warp)
    double* w = v + npad;       // previous w
    double* u = w + npad;       // current reflector vector
    double* p = u + npad;       // tau * A u
    double* prow = p + npad;    // lower-triangle row sums of A u
    double* d = prow + npad;
    double* e = d + npad;
    double* pcol = e + npad;               // [NW][npad] per-warp column sums
    double* A = pcol + (size_t)NW * npad;  // packed lower triangle (+ kFusedPad)

    // ---- load + condition the Gram matrix: power-of-four scale so that |G_ij| <= 1 and the
    // singular values un-scale exactly; NaN/Inf anywhere in W shows on the Gram diagonal.
    double* __restrict__ G = ws + it.gram_off;
    double md = 0.0;
    int bad = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int c = lane + 32 * q;
        if (c < n) {
            const double g = G[((c * (c + 1)) >> 1) + c];
            if (!isfinite(g)) bad = 1;
            md = fmax(md, g);
        }
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
    const int total = (int)tri(n), in_smem = (int)tri(rows_smem);
    for (int i = tid; i < in_smem; i += nthreads) A[i] = G[i] * scale;
    for (int i = in_smem + tid; i < total; i += nthreads) G[i] *= scale;  // rows >= rows_smem: in place
    for (int i = tid; i < kFusedPad; i += nthreads) A[in_smem + i] = 0.0;
    // row r of the evolving matrix
    auto rowptr = [&](int r) -> double* { return (r < rows_smem ? A : G) + ((r * (r + 1)) >> 1); };
    for (int i = tid; i < npad; i += nthreads) {
        v[i] = 0.0;
        w[i] = 0.0;
        u[i] = 0.0;
        p[i] = 0.0;
    }
    __syncthreads();

    double vq[NQ], wq[NQ], uq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) vq[q] = wq[q] = uq[q] = 0.0;

    for (int i = n - 1; i >= 1; --i) {
        const int m = i;  // leading block order; rows 0..m-1 remain after this step
        // ---- (1) every warp: bring row i up to date with (v,w) of the previous step and
        //          build the reflector that annihilates x[0..m-2]
        double tau, beta;
        {
            const double* rowi = rowptr(i);
            const double vi = v[i], wi = w[i];
            double ss = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                uq[q] = 0.0;
                if (32 * q < m) {  // warp-uniform: later chunks are already eliminated
                    const int c = lane + 32 * q;
                    double x = rowi[c];  // c > i reads past the row: masked below
                    x = fma(-vi, wq[q], x);
                    x = fma(-wi, vq[q], x);
                    x = (c < m) ? x : 0.0;
                    uq[q] = x;
                    if (c < m - 1) ss = fma(x, x, ss);
                }
            }
            const double xnorm2 = warp_sum(ss);
            const double alpha = fma(-wi, v[m - 1], fma(-vi, w[m - 1], rowi[m - 1]));
            beta = alpha;
            tau = 0.0;
            double vscale = 0.0;
            if (xnorm2 > 0.0) {
                // |x| <= n after the power-of-four scaling, so s is a normal number unless the
                // whole row is ~1e-154, in which case the IEEE path below is taken.
                const double s2 = fma(alpha, alpha, xnorm2);
                if (s2 > 1e-280) {
                    const double rs = fast_rsqrt(s2);          // 1/||x||
                    const double nrm = s2 * rs;                // ||x||
                    beta = -copysign(nrm, alpha);
                    tau = fma(fabs(alpha), rs, 1.0);           // (beta - alpha)/beta = 1 + |alpha|/||x||
                    vscale = copysign(fast_rcp(fabs(alpha) + nrm), alpha);  // 1/(alpha - beta)
                } else {
                    beta = -copysign(sqrt(s2), alpha);
                    tau = (beta - alpha) / beta;
                    vscale = 1.0 / (alpha - beta);
                }
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (32 * q < m) {
                    const int c = lane + 32 * q;
                    uq[q] = (tau == 0.0) ? 0.0 : ((c == m - 1) ? 1.0 : uq[q] * vscale);
                    if (c < m) u[c] = uq[q];
                }
            }
            if (tid == 0) {
                e[i - 1] = beta;
                d[i] = fma(-2.0 * vi, wi, rowi[i]);
            }
            __syncwarp();
        }

        // ---- (2) fused pass over rows r < m
        double colacc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) colacc[q] = 0.0;
        for (int r0 = warp; r0 < m; r0 += NW * 8) {
            double rs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = r0 + j * NW;
                rs[j] = 0.0;
                if (r < m)  // warp-uniform
                    rs[j] = row_dispatch<NQ>(rowptr(r), r, lane, v[r], w[r], u[r], vq, wq, uq, colacc);
            }
            fold8(rs, lane);
            if ((lane & 3) == 0) {
                const int j = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                const int r = r0 + j * NW;
                if (r < m) prow[r] = rs[0];
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = lane + 32 * q;
            if (32 * q < m && c < m) pcol[warp * npad + c] = colacc[q];
        }
        __syncthreads();

        // ---- (3a) p = tau (A u), one column per thread
        for (int c = tid; c < m; c += nthreads) {
            double s0 = prow[c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = 0;
            for (; k + 3 < NW; k += 4) {
                s0 += pcol[k * npad + c];
                s1 += pcol[(k + 1) * npad + c];
                s2 += pcol[(k + 2) * npad + c];
                s3 += pcol[(k + 3) * npad + c];
            }
            for (; k < NW; ++k) s0 += pcol[k * npad + c];
            p[c] = tau * ((s0 + s1) + (s2 + s3));
        }
        __syncthreads();
        // ---- (3b) every warp: w = p - (tau/2)(p.u) u, v = u  (registers + broadcast copies)
        {
            double pq[NQ];
            double dot = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                pq[q] = 0.0;
                if (32 * q < m) {
                    const int c = lane + 32 * q;
                    pq[q] = (c < m) ? p[c] : 0.0;
                    dot = fma(pq[q], uq[q], dot);
                }
            }
            const double a2 = -0.5 * tau * warp_sum(dot);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (32 * q < m) {
                    const int c = lane + 32 * q;
                    wq[q] = fma(a2, uq[q], pq[q]);
                    vq[q] = uq[q];
                    if (c < m) {
                        w[c] = wq[q];
                        v[c] = vq[q];
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid == 0) {
        d[0] = A[0];
        e[n - 1] = 0.0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthreads) {
        out[i] = d[i];
        out[n + i] = e[i];
    }
    if (tid == 0) {
        out[2 * n + MISC_SCALE] = scale;
        out[2 * n + MISC_FLAGS] = 0.0;
    }
}

}  // namespace vsp
