// gram_i8.cuh -- stage 1 for fp32 inputs on the 5th-generation tensor cores:
// an exact ("FP32-emulated split", Ozaki-style) Gram matrix with tcgen05 / TMEM / TMA.
//
//   x_ik = q_ik * 2^(E_i - 167),  q_ik a 42-bit signed integer (E_i = largest biased exponent
//   of Gram row i, so q keeps every mantissa bit of elements within 2^-17 of the row maximum
//   and rounds smaller ones at 2^-41 of it),
//   q = sum_{t=0..5} s_t 128^(5-t),  s_t in [-64, 64]   (balanced base-128 digits, int8),
//   sum_k q_ik q_jk = sum_{s} 128^(10-s) P_s[i][j],   P_s = sum_{t+t'=s} S_t S_t'^T.
//
// Every P_s is an int8 x int8 -> int32 tensor-core product (tcgen05.mma kind::i8): exact, no
// rounding anywhere (|P_s| <= 6 K 64^2: < 2^27 for K <= 4096, < 2^31 up to the route's limit kI8MaxK).  Levels s = 0..6 (26 digit pairs) are
// kept: the dropped ones are below 2^-46 of the result, i.e. the Gram matrix is as accurate as
// the FP64-accumulated one (5e-15 measured) and the 1e-5 singular-value gate survives the
// squaring of the condition number (plain 3xTF32 does not).
//
// Two kernels:
//   slice_i8_kernel    fp32 W -> row exponents E[n] and six int8 digit planes S_t[n][Kp],
//                      K-major, K padded to a multiple of 64 with zeros (also transposes the
//                      tall case, so the MMA kernel only ever sees K-major operands).
//   gram_i8_mma_kernel one CTA per 128-row block of one matrix, looping over its 64-column
//                      tiles of the lower triangle (TMEM, barriers and the pipeline are set up once):
//        warp 0    TMA producer: 12 boxes per 64-byte K chunk (6 A planes 128x64B, 6 B planes
//                  64x64B, SWIZZLE_64B) into a 3-stage shared-memory ring (mbarrier full/empty)
//        warp 1    allocates 512 TMEM columns, one elected lane issues 26 pairs x 2
//                  tcgen05.mma (M=128, N=64, K=32) per chunk into 7 int32 accumulators
//                  (one per level s, 64 columns each), tcgen05.commit frees the stage
//        warps 2-5 epilogue: tcgen05.ld the 7 levels, Horner in FP64
//                  (acc = acc*128 + P_s), scale by 2^(E_i+E_j-306), store the packed triangle.
#pragma once

#include <cuda.h>

#include "common.cuh"  // poff()

namespace vsp {

constexpr int kDigits = 6;     // int8 digit planes per matrix
constexpr int kLevels = 7;     // P_s levels kept (s = 0..6)
constexpr int kI8TileM = 128;  // UMMA M
constexpr int kI8TileN = 64;   // UMMA N
constexpr int kI8ChunkK = 64;  // bytes of K per pipeline stage (= SWIZZLE_64B span)
constexpr int kI8Stages = 3;
constexpr int kI8EpiWarps = 8;                    // two per TMEM lane quarter: 32 rows x 32 columns each
constexpr int kI8Threads = 32 * (2 + kI8EpiWarps);  // + TMA producer + MMA issuer
constexpr int kI8StageBytes = kDigits * (kI8TileM + kI8TileN) * kI8ChunkK;  // 73728
constexpr int kI8SmemBytes = kI8Stages * kI8StageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kI8MaxTiles = 64;
// Longest contraction the int8 route takes: |P_s| <= 6 K 64^2 must stay below 2^31 (K < 87 381) and the low Horner
// half P3 128^3 + ... + P6 below 2^53; 65 536 leaves a factor of 1.3.  Longer K: FP64 Gram kernel (vspectra_api.cu).
constexpr int kI8MaxK = 65536;

// Gram class: all items of a plan with the same (n, Kp); they share one digit-plane tensor.
struct I8Class {
    int n, kp;          // Gram order, padded contraction length (bytes per digit row)
    int begin, count;   // item range in the plan's sorted item table (sorted by n, then kp)
    int64_t slice_off;  // byte offset of the [count][6][n][kp] digit planes in the workspace
    int64_t exp_off;    // byte offset of the [count][n] int32 row exponents
    int mtiles;                          // 128-row tiles
    unsigned char nt_count[kI8MaxTiles];  // 64-column tiles computed by the 128-row tile (see xt)
    // When the last 128-row tile has at most 64 rows (n = 192: rows 128..191), its off-diagonal blocks are the transposes
    // of the blocks (earlier row tile) x (column tile xt = columns 128 mtl .. + 63), which fill their MMAs completely:
    // every earlier row tile computes that one extra column tile and stores it transposed, the last row tile only
    // its diagonal block (n = 192: 3 + 1 tiles instead of 2 + 3).  xt < 0: plain lower-triangle enumeration.
    int xt;
};

// column tile computed as tile `ti` of row tile `mt`, and whether it is stored transposed
__host__ __device__ inline int i8_tile_nt(const I8Class& c, int mt, int ti, bool& transposed) {
    transposed = false;
    if (c.xt < 0) return ti;
    if (mt == c.mtiles - 1) return c.xt;  // the last row tile: its diagonal block only
    if (ti == 2 * mt + 2) {
        transposed = true;
        return c.xt;
    }
    return ti;
}

// ---------------------------------------------------------------------------------- slicing
__device__ __forceinline__ int f32_exp_field(unsigned bits) {
    const int ex = (bits >> 23) & 0xff;
    return ex ? ex : 1;  // zero / denormal: exponent field 1, no implicit bit
}

// Six balanced base-128 digits (most significant first) of four consecutive elements of a Gram row, relative to the
// row exponent E, packed as the four bytes of one word per digit plane.
//   q = round(x 2^(167-E)), |q| < 2^41 (elements more than 2^17 below the row maximum lose their low bits here);
//   balanced digits s_t in [-64, 63] (s_0 up to 64) of q = ordinary base-128 digits of qb = q + 64 (128^5 + ... + 1),
//   minus 64: no carry chain.
// One FMA does the scaling, the rounding (to nearest even) and the bias: t = x 2^(167-E) + (1.5 2^52 + bias) has qb
// in its low mantissa bits (0 <= qb < 129 128^5 < 2^43), so the digits are 7-bit fields of the two words of t --
// 32-bit shifts and masks that land each field directly in its byte; the "- 64" is applied per packed word (a 7-bit
// field d minus 64 as a two's-complement byte is d with bit 6 flipped and bit 7 = not bit 6).  About 25 instructions
// per element; the 64-bit integer version this replaces took about 70 and made the slice kernel ALU-bound.
// res2 accumulates the SQUARES of the rounding residuals in units of the row's quantum 2^(E-167) (0 for elements
// within 2^17 of the row maximum, at most 1/4 otherwise).  A Gram row whose residuals add up to more than one quantum
// (a dozen rounded elements -- outlier columns, a single huge entry; a random-init row has none or one) marks the
// matrix "rounded": its Gram matrix is then only accurate to ~K 2^-42 ||G|| and the eigensolve hands it to the FP64
// re-solve earlier (kRefineRatioInexact, bisect_metrics.cuh).  Elements beyond the row end are passed as 0.0f
// (digits 0, no residual).
__device__ __forceinline__ void f32_digits4(const float (&xv)[4], int E, unsigned (&pk)[kDigits], double& res2) {
    constexpr long long kBias = 64LL * ((1LL << 42) - 1) / 127;
    const double magic = 6755399441055744.0 + (double)kBias;  // 1.5 * 2^52 + bias (exact)
    const double scale = __hiloint2double((1190 - E) << 20, 0);  // 2^(167 - E), E = 1..254 (255: the row is NaN/Inf)
#pragma unroll
    for (int t = 0; t < kDigits; ++t) pk[t] = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const double xd = (double)xv[u];
        const double tt = fma(xd, scale, magic);
        const double q = tt - magic;                 // exact: the rounded integer
        const double res = fma(xd, scale, -q);       // exact: |res| <= 1/2
        res2 = fma(res, res, res2);
        const unsigned lo = (unsigned)__double2loint(tt);
        const unsigned hi = (unsigned)__double2hiint(tt) - 0x43380000u;  // qb >> 32
        const int sb = 8 * u;
        // planes 5..2: bits 0-6, 7-13, 14-20, 21-27 of lo, moved straight to byte u
        pk[5] |= (lo << sb) & (0x7fu << sb);
        pk[4] |= (sb >= 7 ? (lo << (sb - 7)) : (lo >> (7 - sb))) & (0x7fu << sb);
        pk[3] |= (sb >= 14 ? (lo << (sb - 14)) : (lo >> (14 - sb))) & (0x7fu << sb);
        pk[2] |= (sb >= 21 ? (lo << (sb - 21)) : (lo >> (21 - sb))) & (0x7fu << sb);
        // plane 1: bits 28-34 (four bits of lo, three of hi)
        pk[1] |= (__funnelshift_r(lo, hi, 28) & 0x7fu) << sb;
        // plane 0: qb >> 35 in 0..128: plain subtraction
        pk[0] |= (((hi >> 3) - 64u) & 0xffu) << sb;
    }
#pragma unroll
    for (int t = 1; t < kDigits; ++t) pk[t] = (pk[t] ^ 0x40404040u) | ((~pk[t] & 0x40404040u) << 1);
}

// One CTA per (item, block of 32 Gram rows).  256 threads.
//   trans == 0 : Gram row i = row i of W (contiguous K): warp per row.
//   trans == 1 : Gram row i = column i of W: lanes on columns, digits transposed through smem.
__global__ void __launch_bounds__(256, 6)
    slice_i8_kernel(const ItemDesc* __restrict__ items, I8Class cls, unsigned char* __restrict__ wsb, int* __restrict__ inexact) {
    const ItemDesc it = items[cls.begin + blockIdx.x];
    const int n = it.n, K = it.kdim, kp = cls.kp;
    const int i0 = blockIdx.y * 32;
    if (i0 >= n) return;
    const float* __restrict__ W = reinterpret_cast<const float*>(it.ptr);
    const int64_t ld = it.ld;
    signed char* __restrict__ planes =
        reinterpret_cast<signed char*>(wsb + cls.slice_off) + (int64_t)blockIdx.x * kDigits * n * kp;
    int* __restrict__ Eout = reinterpret_cast<int*>(wsb + cls.exp_off) + (int64_t)blockIdx.x * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int sE[32];
    __shared__ __align__(16) signed char sdig[2][kDigits][32][17 * 4];  // [chunk parity][plane][Gram row][17 words]: see the transposing branch
    bool any_rounded = false;  // some Gram row of this block accumulated more than one quantum of rounding

    if (!it.trans) {
        for (int r = warp; r < 32; r += 8) {
            const int i = i0 + r;
            if (i >= n) continue;  // warp-uniform
            const float* row = W + (int64_t)i * ld;
            if (K <= 256 && (reinterpret_cast<uintptr_t>(row) & 15) == 0 && (K & 3) == 0) {
                // short aligned row: ONE pass, the row lives in two float4 per lane between its load and the digit stores
                // (the two-pass form below reads every row twice with a shuffle reduction in between: latency-bound)
                float xa[4] = {0.f, 0.f, 0.f, 0.f}, xb[4] = {0.f, 0.f, 0.f, 0.f};
                const int ka = 4 * lane, kb = 128 + 4 * lane;
                if (ka < K) {
                    const float4 f = *reinterpret_cast<const float4*>(row + ka);
                    xa[0] = f.x, xa[1] = f.y, xa[2] = f.z, xa[3] = f.w;
                }
                if (kb < K) {
                    const float4 f = *reinterpret_cast<const float4*>(row + kb);
                    xb[0] = f.x, xb[1] = f.y, xb[2] = f.z, xb[3] = f.w;
                }
                int e = 1;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    e = max(e, (int)((__float_as_uint(xa[u]) >> 23) & 0xff));
                    e = max(e, (int)((__float_as_uint(xb[u]) >> 23) & 0xff));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
                if (lane == 0) Eout[i] = e;
                double res2 = 0.0;
                unsigned pk[kDigits];
                if (ka < kp) {
                    f32_digits4(xa, e, pk, res2);
#pragma unroll
                    for (int t = 0; t < kDigits; ++t) *reinterpret_cast<unsigned*>(planes + ((int64_t)t * n + i) * kp + ka) = pk[t];
                }
                if (kb < kp) {
                    f32_digits4(xb, e, pk, res2);
#pragma unroll
                    for (int t = 0; t < kDigits; ++t) *reinterpret_cast<unsigned*>(planes + ((int64_t)t * n + i) * kp + kb) = pk[t];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) res2 += __shfl_xor_sync(0xffffffffu, res2, o);
                any_rounded |= res2 > 1.0;
                continue;
            }
            int e = 1;
            const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
            if (vec) {  // 128-bit loads, four in flight per lane
                const int K4 = K >> 2;
                const float4* row4 = reinterpret_cast<const float4*>(row);
                int k4 = lane;
                for (; k4 + 96 < K4; k4 += 128) {
                    float4 f[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) f[u] = row4[k4 + 32 * u];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        e = max(e, (int)((__float_as_uint(f[u].x) >> 23) & 0xff));
                        e = max(e, (int)((__float_as_uint(f[u].y) >> 23) & 0xff));
                        e = max(e, (int)((__float_as_uint(f[u].z) >> 23) & 0xff));
                        e = max(e, (int)((__float_as_uint(f[u].w) >> 23) & 0xff));
                    }
                }
                for (; k4 < K4; k4 += 32) {
                    const float4 f = row4[k4];
                    e = max(e, (int)((__float_as_uint(f.x) >> 23) & 0xff));
                    e = max(e, (int)((__float_as_uint(f.y) >> 23) & 0xff));
                    e = max(e, (int)((__float_as_uint(f.z) >> 23) & 0xff));
                    e = max(e, (int)((__float_as_uint(f.w) >> 23) & 0xff));
                }
                for (int k = (K4 << 2) + lane; k < K; k += 32) e = max(e, (int)((__float_as_uint(row[k]) >> 23) & 0xff));
            } else {
                for (int k = lane; k < K; k += 32) e = max(e, (int)((__float_as_uint(row[k]) >> 23) & 0xff));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
            if (lane == 0) Eout[i] = e;
            double res2 = 0.0;
            // four consecutive k per lane: one 128-bit load (when the row is 16-byte aligned), one 32-bit store per plane
            for (int k = 4 * lane; k < kp; k += 128) {
                unsigned pk[kDigits];
                float xv[4] = {0.f, 0.f, 0.f, 0.f};
                if (vec && k + 3 < K) {
                    const float4 f = *reinterpret_cast<const float4*>(row + k);
                    xv[0] = f.x, xv[1] = f.y, xv[2] = f.z, xv[3] = f.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (k + u < K) xv[u] = row[k + u];
                }
                f32_digits4(xv, e, pk, res2);
#pragma unroll
                for (int t = 0; t < kDigits; ++t)
                    *reinterpret_cast<unsigned*>(planes + ((int64_t)t * n + i) * kp + k) = pk[t];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) res2 += __shfl_xor_sync(0xffffffffu, res2, o);
            any_rounded |= res2 > 1.0;
        }
    } else {
        // column maxima: warp w scans rows k = w, w+8, ...; lanes on the 32 columns of this block
        const int i = i0 + lane;
        int e = 1;
        if (i < n) {
            int k = warp;
            for (; k + 56 < K; k += 64) {  // eight loads in flight per lane
                float f[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) f[u] = W[(int64_t)(k + 8 * u) * ld + i];
#pragma unroll
                for (int u = 0; u < 8; ++u) e = max(e, (int)((__float_as_uint(f[u]) >> 23) & 0xff));
            }
            for (; k < K; k += 8) e = max(e, (int)((__float_as_uint(W[(int64_t)k * ld + i]) >> 23) & 0xff));
        }
        if (tid < 32) sE[tid] = 1;
        __syncthreads();
        atomicMax(&sE[lane], e);
        __syncthreads();
        if (tid < 32 && i0 + tid < n) Eout[i0 + tid] = sE[tid];
        const int Ei = sE[lane];
        double res2 = 0.0;  // this thread's share of Gram row i0 + lane
        // digits staged as 32-bit words (four consecutive k of one Gram row), row pitch 17 words: lanes (= Gram rows) fall
        // into different banks (the byte-granular staging at an 80-byte pitch ran at 4-way conflicts: ncu r01)
        // (two staging buffers, alternating by chunk: one barrier per chunk -- the buffer a chunk writes was last read two
        // chunks ago, and every thread has passed the barrier in between)
        for (int k0 = 0; k0 < kp; k0 += 64) {
            unsigned(*sw)[32][17] = reinterpret_cast<unsigned(*)[32][17]>(&sdig[(k0 >> 6) & 1][0][0][0]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int kw = warp + 8 * j;  // word index inside the 64-byte chunk: k = k0 + 4 kw .. + 3
                unsigned pk[kDigits];
                float xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + 4 * kw + u;
                    xv[u] = (k < K && i < n) ? W[(int64_t)k * ld + i] : 0.f;
                }
                f32_digits4(xv, Ei, pk, res2);
#pragma unroll
                for (int t = 0; t < kDigits; ++t) sw[t][lane][kw] = pk[t];
            }
            __syncthreads();
            // 6 planes x 32 rows x 16 words: 64-byte runs per Gram row
            for (int v = tid; v < kDigits * 32 * 16; v += 256) {
                const int t = v >> 9, r = (v >> 4) & 31, c = v & 15;
                if (i0 + r < n) *reinterpret_cast<unsigned*>(planes + ((int64_t)t * n + i0 + r) * kp + k0 + 4 * c) = sw[t][r][c];
            }
        }
        __syncthreads();
        // per Gram row: the eight warps' shares meet in shared memory (the digit staging buffers are free now)
        float* sres = reinterpret_cast<float*>(&sdig[0][0][0][0]);  // [8][32]
        sres[warp * 32 + lane] = (float)res2;
        __syncthreads();
        if (warp == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += sres[w * 32 + lane];
            any_rounded |= tot > 1.f;
        }
    }
    if (inexact != nullptr && __any_sync(0xffffffffu, any_rounded) && lane == 0) atomicOr(&inexact[cls.begin + blockIdx.x], 1);
}

// -------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug must end in a trap (reported as a CUDA error), never in a hang.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (long long spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1LL << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, SWIZZLE_64B operand tile: rows at 64-byte pitch, 8-row atoms of 512 bytes.
__device__ __forceinline__ unsigned long long umma_desc_sw64(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);  // start address
    d |= (unsigned long long)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (unsigned long long)(512 >> 4) << 32;              // stride byte offset: next 8-row atom
    d |= (unsigned long long)1 << 46;                       // descriptor version (Blackwell)
    d |= (unsigned long long)4 << 61;                       // layout type SWIZZLE_64B
    return d;
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M=128, N=64, K=32
__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                        unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// 16 TMEM lanes (rows) x 32 columns in the MMA-fragment distribution: lane l, register 4m + e + 2v holds
// row (l >> 2) + 8 v, column 8 m + 2 (l & 3) + e  (m = 0..3, e = 0..1, v = 0..1).  Four lanes hold the eight columns of
// a row: after conversion to FP64 a warp's registers (m, v) ARE one 8 x 8 tile of the Gram layout, 16 bytes per lane.
__device__ __forceinline__ void tmem_ld_16x256b_x4(unsigned taddr, int (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// exact int32 -> double without the conversion unit: 2^52 + 2^31 + x has x + 2^31 in its low mantissa word
__device__ __forceinline__ double i32_to_f64(int x) {
    return __hiloint2double(0x43300000, x ^ 0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
}

// ------------------------------------------------------------------------------------- MMA
// Persistent: one CTA per SM walks the work units (item, 128-row tile) u = blockIdx.x, + gridDim.x, ...; the three
// roles run their own loops over the same unit / tile sequence and meet only at the mbarriers, whose phases count
// chunks (ring) and tiles (accumulators) across units.  The seven level accumulators take 448 of the 512 TMEM
// columns, so there is ONE accumulator set and the next tile's MMAs cannot start before the epilogue has read it:
// the epilogue therefore drains the whole tile into registers first (64 FP64 values per thread, Horner in exact
// integers on the way), releases TMEM, and only then scales and stores -- the stores of tile t, the launch
// prologue and the TMEM allocation of a non-persistent CTA no longer sit between two tiles' MMAs.
__global__ void __launch_bounds__(kI8Threads, 1)
    gram_i8_mma_kernel(const ItemDesc* __restrict__ items, I8Class cls, unsigned char* __restrict__ wsb,
                       double* __restrict__ ws, const __grid_constant__ CUtensorMap tmA,
                       const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int units = cls.count * cls.mtiles;
    // unit of this CTA in round r: rotated by r, so that with an even grid a CTA alternates between the 128-row tiles of a
    // matrix (2 and 3 column tiles at n = 192) instead of always drawing the same one
#define VSP_I8_UNIT(r, u)                                                        \
    const int u = (r) * (int)gridDim.x + (int)((blockIdx.x + (r)) % gridDim.x); \
    if (u >= units) continue
    const int nchunks = cls.kp / kI8ChunkK;

    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tiles + kI8Stages * kI8StageBytes);
    // bars[0..S) full, [S..2S) empty, [2S] accumulators ready, [2S+1] accumulators drained; then the TMEM base
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * kI8Stages + 2);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + kI8Stages);
    const unsigned accbar = smem_u32(bars + 2 * kI8Stages), drainbar = smem_u32(bars + 2 * kI8Stages + 1);

    if (tid == 0) {
        for (int s = 0; s < kI8Stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accbar, 1);
        mbar_init(drainbar, kI8EpiWarps);  // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // whole warp: allocate all 512 TMEM columns (7 levels x 64 used)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int g = 0;  // chunk counter over all tiles of all units
            for (int rnd = 0; rnd * (int)gridDim.x < units; ++rnd) {
                VSP_I8_UNIT(rnd, u);
                const int item = u / cls.mtiles, mt = u - item * cls.mtiles;  // tiles of one matrix run on neighbouring SMs: L2
                const int n = cls.n;
                const int ntiles = cls.nt_count[mt];
                const int rowA = item * kDigits * n + mt * kI8TileM;  // + t*n per digit plane
                for (int ti = 0; ti < ntiles; ++ti) {
                    bool tr_unused;
                    const int nt = i8_tile_nt(cls, mt, ti, tr_unused);
                    const int rowB = item * kDigits * n + nt * kI8TileN;
                    for (int c = 0; c < nchunks; ++c, ++g) {
                        const int s = g % kI8Stages;
                        if (g >= kI8Stages) mbar_wait(empty0 + 8 * s, ((g / kI8Stages) - 1) & 1);
                        mbar_expect_tx(full0 + 8 * s, kI8StageBytes);
                        const unsigned base = smem_u32(tiles + s * kI8StageBytes);
#pragma unroll
                        for (int t = 0; t < kDigits; ++t) {
                            tma_load_2d(base + t * (kI8TileM * kI8ChunkK), &tmA, full0 + 8 * s, c * kI8ChunkK, rowA + t * n);
                            tma_load_2d(base + kDigits * kI8TileM * kI8ChunkK + t * (kI8TileN * kI8ChunkK), &tmB,
                                        full0 + 8 * s, c * kI8ChunkK, rowB + t * n);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 64, M = 128
            const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kI8TileN >> 3) << 17) |
                                   ((unsigned)(kI8TileM >> 4) << 24);
            const unsigned long long adesc0 = umma_desc_sw64(smem_u32(tiles));
            const unsigned long long bdesc0 = umma_desc_sw64(smem_u32(tiles) + kDigits * kI8TileM * kI8ChunkK);
            int g = 0, tile = 0;  // chunk / tile counters over all units
#ifdef VSP_PHASE_TIMING
            long long tm[3] = {0, 0, 0}, tk = clock64();
#define VSP_GLAP(k) do { const long long now_ = clock64(); tm[k] += now_ - tk; tk = now_; } while (0)
#else
#define VSP_GLAP(k) ((void)0)
#endif
            for (int rnd = 0; rnd * (int)gridDim.x < units; ++rnd) {
                VSP_I8_UNIT(rnd, u);
                const int ntiles = cls.nt_count[u % cls.mtiles];
                for (int ti = 0; ti < ntiles; ++ti, ++tile) {
                    VSP_GLAP(2);
                    if (tile > 0) {  // the epilogue must have drained the accumulators of the previous tile
                        mbar_wait(drainbar, (tile - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    VSP_GLAP(0);
                    for (int c = 0; c < nchunks; ++c, ++g) {
                        const int s = g % kI8Stages;
                        VSP_GLAP(2);
                        mbar_wait(full0 + 8 * s, (g / kI8Stages) & 1);
                        VSP_GLAP(1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        // descriptors differ from the stage-0 ones only in the 16-byte-granular start
                        // address field: one 64-bit add per operand instead of rebuilding them
                        const unsigned long long adS = adesc0 + (unsigned long long)(s * (kI8StageBytes >> 4));
                        const unsigned long long bdS = bdesc0 + (unsigned long long)(s * (kI8StageBytes >> 4));
#pragma unroll
                        for (int lv = 0; lv < kLevels; ++lv) {
                            constexpr int kTop = kDigits - 1;
                            const int tlo = lv > kTop ? lv - kTop : 0;
                            const int thi = lv < kTop ? lv : kTop;
#pragma unroll
                            for (int t = 0; t < kDigits; ++t) {
                                if (t < tlo || t > thi) continue;
#pragma unroll
                                for (int ks = 0; ks < kI8ChunkK / 32; ++ks) {
                                    const unsigned acc = (c > 0 || t > tlo || ks > 0) ? 1u : 0u;
                                    umma_i8(tmem_base + lv * kI8TileN,
                                            adS + (unsigned long long)(((t * kI8TileM * kI8ChunkK) >> 4) + 2 * ks),
                                            bdS + (unsigned long long)((((lv - t) * kI8TileN * kI8ChunkK) >> 4) + 2 * ks),
                                            idesc, acc);
                                }
                            }
                        }
                        umma_commit(empty0 + 8 * s);  // the stage is free once these MMAs have read it
                    }
                    umma_commit(accbar);
                }
            }
#ifdef VSP_PHASE_TIMING
            VSP_GLAP(2);
            if (blockIdx.x == 0)
                printf("[gram mma kp=%d] tiles %d cycles: wait-drain %lld  wait-full %lld  issue %lld\n", cls.kp, tile, tm[0], tm[1], tm[2]);
#endif
#undef VSP_GLAP
        }
    } else {
        // ===== epilogue: warps 2..9.  A warp may read the TMEM lanes 32 (warp % 4) .. +31; the two warps of a lane
        //       quarter split the 64 columns: 32 rows x 32 columns per warp and tile, 32 FP64 values per thread
        const int q = warp & 3, ch = (warp - 2) >> 2;
        const int lr = lane >> 2, lc = 2 * (lane & 3);  // row and first column of this lane inside an 8 x 8 tile
        int tile = 0;
#ifdef VSP_PHASE_TIMING
        long long te[4] = {0, 0, 0, 0}, tk = clock64();
#define VSP_ELAP(k) do { const long long now_ = clock64(); te[k] += now_ - tk; tk = now_; } while (0)
#else
#define VSP_ELAP(k) ((void)0)
#endif
        for (int rnd = 0; rnd * (int)gridDim.x < units; ++rnd) {
            VSP_I8_UNIT(rnd, u);
            const int item = u / cls.mtiles, mt = u - item * cls.mtiles;
            const ItemDesc it = items[cls.begin + item];
            const int n = it.n;
            const int ntiles = cls.nt_count[mt];
            const int* __restrict__ E = reinterpret_cast<const int*>(wsb + cls.exp_off) + (int64_t)item * n;
            double* __restrict__ G = ws + it.gram_off;
            const int layout = it.full;
            const int foff = ((n + 7) & ~7) - n;  // kGramTiled (sbr8.cuh): frame offset of the top-left padding
            if (layout == kGramTiled && foff > 0) {  // the frame padding is zero (one row per thread of warps 2..5)
                const int i = mt * kI8TileM + (tid - 64);
                if (tid - 64 < kI8TileM && i < n) {
                    const int fi = i + foff;
                    const int frow = tile_off(fi >> 3, 0) + ((fi & 7) << 3);
                    for (int c = 0; c < foff; ++c) G[frow + c] = 0.0;
                    if (i == 0)
                        for (int e = 0; e < 8 * foff; ++e) G[e] = 0.0;
                }
            }
            // rows of this thread: R(h, v) = row0 + 16 h + 8 v (h, v = 0..1); their exponents
            const int row0 = mt * kI8TileM + 32 * q + lr;
            int Er[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) Er[k] = (row0 + 8 * k < n) ? E[row0 + 8 * k] : 1;
            const bool fast = layout == kGramTiled && foff == 0;  // n is a multiple of 8: a warp's (h, v, m) block is one tile
            for (int ti = 0; ti < ntiles; ++ti, ++tile) {
                bool transposed;
                const int nt = i8_tile_nt(cls, mt, ti, transposed);
                const int col0 = nt * kI8TileN + 32 * ch + lc;  // columns of this thread: col0 + 8 m + e
                int Ec[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) Ec[k] = (col0 + 8 * (k >> 1) + (k & 1) < n) ? E[col0 + 8 * (k >> 1) + (k & 1)] : 1;
                VSP_ELAP(3);
                mbar_wait(accbar, tile & 1);
                VSP_ELAP(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // ---- drain: all seven levels of 16 rows x 32 columns per wait; Horner in FP64 on exact integers, split
                //   in two so that nothing is rounded before the last step:
                //   hi = P0 128^2 + P1 128 + P2 (< 2^46),  lo = P3 128^3 + ... + P6 (< 2^53),  sum_s P_s 128^(6-s) = hi 2^28 + lo
                double acc[32];  // [h][4 m + e + 2 v]
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned taddr = tmem_base + ((unsigned)(32 * q + 16 * h) << 16) + 32 * ch;
                    int r[kLevels][16];
#pragma unroll
                    for (int lv = 0; lv < kLevels; ++lv) tmem_ld_16x256b_x4(taddr + lv * kI8TileN, r[lv]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const double hi = fma(fma(i32_to_f64(r[0][k]), 128.0, i32_to_f64(r[1][k])), 128.0, i32_to_f64(r[2][k]));
                        const double lo = fma(fma(fma(i32_to_f64(r[3][k]), 128.0, i32_to_f64(r[4][k])), 128.0, i32_to_f64(r[5][k])),
                                              128.0, i32_to_f64(r[6][k]));
                        acc[16 * h + k] = fma(hi, 268435456.0, lo);  // 2^28
                    }
                }
                // every level of this tile has been read: release TMEM to the next tile's MMAs
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(drainbar) : "memory");
                VSP_ELAP(1);
                // ---- scale and store: sum_k q_i q_j = 128^4 * acc ; x = q 2^(E-167)  ->  2^(Ei+Ej-334+28)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int R = row0 + 16 * h + 8 * v, Ei = Er[2 * h + v];
                        const int I = (mt * kI8TileM + 32 * q + 16 * h + 8 * v) >> 3;  // tile row (warp-uniform)
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const int J = (nt * kI8TileN + 32 * ch + 8 * m) >> 3;  // tile column (warp-uniform)
                            double g[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int Ej = Ec[2 * m + e];
                                const int be = Ei + Ej - 306 + 1023;
                                g[e] = (Ei == 255 || Ej == 255) ? __longlong_as_double(0x7ff8000000000000LL)  // NaN/Inf in the input row
                                                                : acc[16 * h + 4 * m + 2 * v + e] * __hiloint2double(be << 20, 0);
                            }
                            if (fast && transposed) {
                                // block above the diagonal (J > I): stored as its transpose, tile (J, I)
                                if (8 * I >= n || 8 * J >= n) continue;  // warp-uniform
                                double* dst = G + tile_off(J, I) + (lc << 3) + lr;
                                dst[0] = g[0];
                                dst[8] = g[1];
                            } else if (fast) {
                                if (8 * I >= n || J > I) continue;  // warp-uniform
                                if (J == I) {
                                    // diagonal tile: both triangles.  Position (lr, lc + e) above the diagonal takes the value
                                    // of (lc + e, lr), held by lane 4 (lc + e) + (lr >> 1), component lr & 1
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        const int src = 4 * (lc + e) + (lr >> 1);
                                        const double m0 = __shfl_sync(0xffffffffu, g[0], src), m1 = __shfl_sync(0xffffffffu, g[1], src);
                                        if (lc + e > lr) g[e] = (lr & 1) ? m1 : m0;
                                    }
                                }
                                // one 8 x 8 tile = 512 contiguous bytes per warp store
                                *reinterpret_cast<double2*>(G + tile_off(I, J) + 2 * lane) = make_double2(g[0], g[1]);
                            } else if (R < n) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int C = nt * kI8TileN + 32 * ch + 8 * m + lc + e;
                                    const int gi = transposed ? C : R, gj = transposed ? R : C;  // stored position (gi, gj), gj <= gi
                                    if (gj > gi || gi >= n) continue;
                                    if (layout == kGramTiled) {
                                        const int fi = gi + foff, fj = gj + foff;
                                        G[tile_off(fi >> 3, fj >> 3) + ((fi & 7) << 3) + (fj & 7)] = g[e];
                                        if ((fj >> 3) == (fi >> 3)) G[tile_off(fi >> 3, fi >> 3) + ((fj & 7) << 3) + (fi & 7)] = g[e];
                                    } else if (layout == kGramFull) {
                                        G[(int64_t)gi * n + gj] = g[e];
                                        G[(int64_t)gj * n + gi] = g[e];
                                    } else {
                                        G[poff(gi) + gj] = g[e];
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
#ifdef VSP_PHASE_TIMING
        VSP_ELAP(2);
        if (blockIdx.x == 0 && tid == 64)
            printf("[gram epilogue kp=%d] tiles %d cycles: wait-acc %lld  drain %lld  store(last) %lld  setup+store %lld\n", cls.kp, tile, te[0], te[1], te[2], te[3]);
#endif
#undef VSP_ELAP
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
#undef VSP_I8_UNIT
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace vsp
