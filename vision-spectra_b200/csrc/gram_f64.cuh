// gram_f64.cuh -- stage 1 (exact path): Gram matrix of the smaller side,
//   G = W W^T (rows <= cols)   or   G = W^T W (rows > cols),
// accumulated in FP64 from fp32/fp64 inputs.  Products of two fp32 values are exact
// in FP64 (48 <= 53 bits), so G is the correctly-rounded-sum Gram matrix that the
// eigensolve needs to keep sigma within 1e-5 on square matrices.
//
// One CTA computes one TILE x TILE tile of the lower triangle (tile row ti >= tile
// column tj) of one matrix; q/k/v arrive as row-block views of the fused qkv
// buffer (pointer + ld), no copies (reference split: metrics/extraction.py:59-62).
#pragma once

#include "common.cuh"  // gram_store()

namespace vsp {

template <typename TIn, int TILE, int KC>
__global__ void __launch_bounds__(256) gram_f64_kernel(const ItemDesc* __restrict__ items, int item_base,
                                                       double* __restrict__ ws) {
    constexpr int TPB = 256;
    constexpr int RT = TILE / 16;  // outputs per thread per dimension
    __shared__ double As[KC][TILE + 2];
    __shared__ double Bs[KC][TILE + 2];

    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n, K = it.kdim;
    // decode lower-triangular tile index
    int ti = 0, t = blockIdx.y;
    while (t >= ti + 1) {
        t -= ti + 1;
        ++ti;
    }
    const int tj = t;
    const int i0 = ti * TILE, j0 = tj * TILE;
    if (i0 >= n) return;

    const TIn* __restrict__ W = reinterpret_cast<const TIn*>(it.ptr);
    const int64_t ld = it.ld;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    double acc[RT][RT];
#pragma unroll
    for (int a = 0; a < RT; ++a)
#pragma unroll
        for (int b = 0; b < RT; ++b) acc[a][b] = 0.0;

    for (int k0 = 0; k0 < K; k0 += KC) {
        // ---- stage the two operand tiles as FP64
        if (it.trans) {  // G = W^T W: element (i,k) is W[k*ld + i]; i is contiguous
            for (int e = tid; e < TILE * KC; e += TPB) {
                const int ii = e % TILE, kk = e / TILE;
                const int k = k0 + kk;
                double a = 0.0, b = 0.0;
                if (k < K) {
                    if (i0 + ii < n) a = (double)W[(int64_t)k * ld + i0 + ii];
                    if (j0 + ii < n) b = (double)W[(int64_t)k * ld + j0 + ii];
                }
                As[kk][ii] = a;
                Bs[kk][ii] = b;
            }
        } else {  // G = W W^T: element (i,k) is W[i*ld + k]; k is contiguous
            for (int e = tid; e < TILE * KC; e += TPB) {
                const int kk = e % KC, ii = e / KC;
                const int k = k0 + kk;
                double a = 0.0, b = 0.0;
                if (k < K) {
                    if (i0 + ii < n) a = (double)W[(int64_t)(i0 + ii) * ld + k];
                    if (j0 + ii < n) b = (double)W[(int64_t)(j0 + ii) * ld + k];
                }
                As[kk][ii] = a;
                Bs[kk][ii] = b;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            double a[RT], b[RT];
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                a[r] = As[kk][ty + 16 * r];
                b[r] = Bs[kk][tx + 16 * r];
            }
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int c = 0; c < RT; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }

    double* G = ws + it.gram_off;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const int i = i0 + ty + 16 * r;
#pragma unroll
        for (int c = 0; c < RT; ++c) {
            const int j = j0 + tx + 16 * c;
            if (i < n && j <= i) gram_store(G, it.full, n, i, j, acc[r][c]);
        }
    }
}

}  // namespace vsp
