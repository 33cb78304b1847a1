// dgemm_dmma.cuh -- batched FP64 GEMM on the FP64 tensor cores (DMMA.8x8x4) for the singular-vector consumers
// (tail truncation, metrics/tail_truncation.py:63-152; rank-reducing gradient U V^T,
// metrics/gradient_alignment.py:48-70).  The hot path computes no singular vectors; both consumers are matrix
// FUNCTIONS of W -- W P_k with the spectral projector P_k = (I + sign(G - mu I)) / 2, and the polar factor
// W (W^T W)^(-1/2) -- which Newton-Schulz iterations evaluate with nothing but matrix products (lowrank.py).
//
//     C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b] + gamma * I          b = 0 .. batch-1, row-major, FP64
//
// One CTA per 64 x 64 tile of C, eight warps (4 x 2), warp tile 16 x 32 = 2 x 4 DMMA tiles, K in chunks of 16 staged
// through shared memory (A chunk as [m][k], B chunk as [k][n], padded: fragment loads hit distinct banks).  Operand
// loads are element-wise with bounds and transposition folded into the index -- this is an auxiliary op, not the hot path.
#pragma once

#include "sbr8.cuh"  // dmma8

namespace vsp {

constexpr int kGemmTile = 64;
constexpr int kGemmK = 16;

#if defined(__CUDACC__)

__global__ void __launch_bounds__(256)
    dgemm_dmma_kernel(int M, int N, int K, double alpha, const double* const* __restrict__ Aptr, int64_t lda, int transA,
                      const double* const* __restrict__ Bptr, int64_t ldb, int transB, double beta, double gamma,
                      double* const* __restrict__ Cptr, int64_t ldc) {
    __shared__ double As[kGemmTile][kGemmK + 4];  // [m][k]: stride 20 doubles -> rows g at 160 B, k-slices t contiguous
    __shared__ double Bs[kGemmK][kGemmTile + 8];  // [k][n]: stride 72 doubles
    const double* __restrict__ A = Aptr[blockIdx.z];
    const double* __restrict__ B = Bptr[blockIdx.z];
    double* __restrict__ C = Cptr[blockIdx.z];
    const int m0 = blockIdx.y * kGemmTile, n0 = blockIdx.x * kGemmTile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;  // warp tile: rows 16 wm .., columns 32 wn ..
    double2 acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_double2(0.0, 0.0);

    for (int k0 = 0; k0 < K; k0 += kGemmK) {
        for (int e = tid; e < kGemmTile * kGemmK; e += 256) {
            // A chunk: consecutive threads along the contiguous direction of the source
            const int mm = transA ? (e % kGemmTile) : (e / kGemmK), kk = transA ? (e / kGemmTile) : (e % kGemmK);
            const int m = m0 + mm, k = k0 + kk;
            double v = 0.0;
            if (m < M && k < K) v = transA ? A[(int64_t)k * lda + m] : A[(int64_t)m * lda + k];
            As[mm][kk] = v;
            const int nn = transB ? (e / kGemmK) : (e % kGemmTile), kb = transB ? (e % kGemmK) : (e / kGemmTile);
            const int n = n0 + nn, k2 = k0 + kb;
            double w = 0.0;
            if (n < N && k2 < K) w = transB ? B[(int64_t)n * ldb + k2] : B[(int64_t)k2 * ldb + n];
            Bs[kb][nn] = w;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < kGemmK; ks += 4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[16 * wm + 8 * i + g][ks + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[ks + t][32 * wn + 8 * j + g];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma8(acc[i][j].x, acc[i][j].y, a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + 16 * wm + 8 * i + g;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 32 * wn + 8 * j + 2 * t;
            double* c = C + (int64_t)m * ldc + n;
            if (n < N) {
                double v = alpha * acc[i][j].x;
                if (beta != 0.0) v = fma(beta, c[0], v);
                if (n == m) v += gamma;
                c[0] = v;
            }
            if (n + 1 < N) {
                double v = alpha * acc[i][j].y;
                if (beta != 0.0) v = fma(beta, c[1], v);
                if (n + 1 == m) v += gamma;
                c[1] = v;
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace vsp
