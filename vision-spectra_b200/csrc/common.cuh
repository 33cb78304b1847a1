// common.cuh -- shared definitions for the weight-spectrum kernels.
//
// The per-matrix algorithms (tridiagonalisation, bisection, metrics) are written
// against a small "cooperative context" (tid / nthreads / sync / all-reduce).  On
// the device the context is one CTA; compiled by a host compiler it degenerates to
// a single thread, which lets tests/test_host_emul.py run the very same source on
// the CPU to check indexing and numerics before any GPU time is spent.  The host
// build is test infrastructure only; nothing in the product calls it.
#pragma once

#include <cmath>
#include <cstdint>

#include "../../include/vspectra.h"

#if defined(__CUDACC__)
#define VSP_DEV __device__ __forceinline__
#define VSP_HD __host__ __device__ __forceinline__
#else
#define VSP_DEV inline
#define VSP_HD inline
#endif

namespace vsp {

// One matrix of a batch, as the kernels see it.  Built on the host by the plan,
// uploaded once; only `ptr` changes between executions.
struct ItemDesc {
    const void* ptr;   // device pointer to the row-major matrix
    int64_t ld;        // row stride in elements
    int64_t gram_off;  // offset (doubles) of this item's Gram matrix in the workspace
    int64_t de_off;    // offset (doubles) of d[n], e[n], misc[4]
    int64_t sv_off;    // offset (doubles) into the packed SV output
    int32_t rows, cols;
    int32_t n;         // min(rows, cols): order of the Gram matrix
    int32_t kdim;      // max(rows, cols): contraction length
    int32_t trans;     // 1: Gram = W^T W (rows > cols), 0: Gram = W W^T
    int32_t item;      // index in the caller's batch
    int32_t full;      // Gram layout: kGramPacked (row-packed lower), kGramFull (n*n), kGramTiled (8x8 tiles, sbr8.cuh)
    int32_t pad;
};

VSP_HD int64_t tri(int64_t i) { return (i * (i + 1)) >> 1; }

enum { kGramPacked = 0, kGramFull = 1, kGramTiled = 2 };

// kGramTiled: the matrix sits at the bottom-right of an N x N frame, N = round_up(n, 8) (off = N - n zero rows / columns
// at the top-left); lower triangle of 8 x 8 tiles, tile (I, J), J <= I, at (I (I + 1) / 2 + J) * 64 doubles, row-major
// inside the tile (= the DMMA C-fragment order); diagonal tiles hold both triangles.
VSP_HD int tile_off(int I, int J) { return (((I * (I + 1)) >> 1) + J) << 6; }

// offset of row r in the padded-even packed lower triangle: row r starts at tri(r) + (r+1)/2, i.e. rows are
// padded to an even length so that every row starts 16-byte aligned (the Gram kernels write this layout)
VSP_HD int poff(int r) { return ((r * (r + 1)) >> 1) + ((r + 1) >> 1); }

#if defined(__CUDACC__)
// store G[i][j] (j <= i) of an n x n Gram matrix in the item's layout (the Gram kernels' epilogues)
__device__ __forceinline__ void gram_store(double* __restrict__ G, int layout, int n, int i, int j, double g) {
    if (layout == kGramFull) {
        G[(int64_t)i * n + j] = g;
        G[(int64_t)j * n + i] = g;
    } else if (layout == kGramTiled) {
        const int off = ((n + 7) & ~7) - n;
        const int fi = i + off, fj = j + off;
        const int base = tile_off(fi >> 3, fj >> 3);
        G[base + ((fi & 7) << 3) + (fj & 7)] = g;
        if ((fi >> 3) == (fj >> 3) && i != j) G[base + ((fj & 7) << 3) + (fi & 7)] = g;  // diagonal tile: mirror
        if (off > 0 && j == 0) {  // zero the frame padding once: columns < off of this row, and the rows < off (by row 0)
            for (int c = 0; c < off; ++c) G[tile_off(fi >> 3, 0) + ((fi & 7) << 3) + c] = 0.0;
            if (i == 0)
                for (int e = 0; e < 8 * off; ++e) G[e] = 0.0;
        }
    } else {
        G[poff(i) + j] = g;
    }
}

__device__ __forceinline__ double shfl_xor_d(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

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
#endif

// misc[] slots written by the tridiagonalisation stage
enum { MISC_SCALE = 0, MISC_FLAGS = 1, MISC_SLOT = 2, MISC_UNUSED1 = 3, MISC_COUNT = 4 };

// Work list of the ill-conditioned re-solve: the tridiagonalisation kernels test the small end of
// the spectrum with one Sturm count and, if it is populated, append the item to the list.  The list
// has one entry per item of the shape class, so it cannot overflow; the FP64 pool only bounds how
// many entries are re-solved at the same time (refine_kernel: CTA b serves entries b, b + slots, ...).
struct RefineGate {
    int* counter;     // zeroed by the host before every execution; nullptr = re-solve disabled
    int* slot_items;  // list entry -> position of the flagged item in the plan's item table
    int slots;        // pool buffers = CTAs of the re-solve launch
    const int* inexact;  // per plan item (sorted position): stage 1 rounded an element of this matrix; may be nullptr
};

#if !defined(__CUDACC__)
// ----------------------------------------------------------------- host emulation
using std::copysign;
using std::fabs;
using std::fmax;
using std::fmin;
using std::frexp;
using std::isfinite;
using std::ldexp;
using std::log;
using std::sqrt;

struct HostCtx {
    int tid = 0;
    int nthreads = 1;
    int lane = 0, warp = 0, nwarps = 1, wsize = 1;  // a "warp" of one lane
    double warp_sum(double v) { return v; }
    void sync() {}
    double sum(double v) { return v; }
    double max(double v) { return v; }
    double min(double v) { return v; }
    void sum4(double (&v)[4]) { (void)v; }
    int sum_i(int v) { return v; }
    int max_i(int v) { return v; }
    int fetch_add(int* p) { return (*p)++; }
};
#else
// --------------------------------------------------------------------- device CTA
// Block all-reduce with one barrier per call: two scratch rows are used alternately,
// so row A is not rewritten before every thread has passed the barrier of the
// following reduce (which used row B) and therefore finished reading A.
struct CtaCtx {
    int tid;
    int nthreads;
    int lane, warp, nwarps, wsize;
    double* red;  // shared scratch: 2 rows x 32 warps x 4 values
    int flip;

    __device__ CtaCtx(double* scratch)
        : tid(threadIdx.x), nthreads(blockDim.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5),
          nwarps(blockDim.x >> 5), wsize(32), red(scratch), flip(0) {}
    __device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    static constexpr int kScratchDoubles = 2 * 32 * 4;

    __device__ __forceinline__ void sync() { __syncthreads(); }

    template <class Op>
    __device__ __forceinline__ double reduce(double v, Op op) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
        double* row = red + flip * 128;
        flip ^= 1;
        const int nw = (nthreads + 31) >> 5;
        if ((tid & 31) == 0) row[tid >> 5] = v;
        __syncthreads();
        if (nw <= 8) {
            double r = row[0];
            for (int w = 1; w < nw; ++w) r = op(r, row[w]);
            return r;
        }
        // many warps: one partial per lane and a second (masked) butterfly instead of nw serial loads
        bool have = lane < nw;
        double r = have ? row[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, r, o);
            const bool ohave = __shfl_xor_sync(0xffffffffu, (int)have, o) != 0;
            if (have && ohave)
                r = op(r, other);
            else if (ohave) {
                r = other;
                have = true;
            }
        }
        return r;
    }
    __device__ __forceinline__ double sum(double v) {
        return reduce(v, [](double a, double b) { return a + b; });
    }
    __device__ __forceinline__ double max(double v) {
        return reduce(v, [](double a, double b) { return fmax(a, b); });
    }
    __device__ __forceinline__ double min(double v) {
        return reduce(v, [](double a, double b) { return fmin(a, b); });
    }
    __device__ __forceinline__ void sum4(double (&v)[4]) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        }
        double* row = red + flip * 128;
        flip ^= 1;
        const int nw = (nthreads + 31) >> 5;
        if ((tid & 31) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) row[(tid >> 5) * 4 + q] = v[q];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double r = row[q];
            for (int w = 1; w < nw; ++w) r += row[w * 4 + q];
            v[q] = r;
        }
    }
    __device__ __forceinline__ int sum_i(int v) { return (int)(sum((double)v) + 0.5); }
    __device__ __forceinline__ int max_i(int v) { return (int)max((double)v); }
    __device__ __forceinline__ int fetch_add(int* p) { return atomicAdd(p, 1); }  // p in shared memory
};
#endif

}  // namespace vsp
