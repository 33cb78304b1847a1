// sbr_band.cuh -- stage 2a-1, n <= kSmemMaxN: blocked Householder reduction of the (scaled)
// Gram matrix to a symmetric band of bandwidth b = 4, in place.  band_tridiag.cuh finishes the
// job.  Together they replace the unblocked tridiag_fused.cuh on the product path: the unblocked
// reduction moves every stored element through shared memory once per *column* (4 DFMA per
// 16 bytes: shared-memory bound at <= 50 % of the FP64 pipe); here an element moves once per
// *panel of four columns* and sees 16 DFMA per 16 bytes, so the pass is FP64-bound.
//
// Elimination order is bottom-up (as in tridiag_fused.cuh): the active matrix is the leading
// m x m block.  One panel step, with p0 = m - 4:
//   (1) mini-pass  (all threads) rows p0..m-1 get the pending rank-8 update  A -= V W^T + W V^T;
//                  their diagonal block goes to the band output, the b x p0 block left of it to P.
//   (2) LQ         (warp 0) four Householder reflectors (row p0+3 first, pivot column p0-1, then
//                  p0+2 / p0-2, ...) reduce P to an upper-triangular 4x4 block R next to the diagonal
//                  block: Q = H_0 H_1 H_2 H_3 = I - U T U^T (compact WY, T built from U^T U).
//   (3) pass       (all warps) over the leading p0 x p0 triangle, once:
//                      A -= V W^T + W V^T   (pending update of the previous panel, 8 DFMA/element)
//                      Y  = A U             (symmetric, both triangles from one read, 8 DFMA/element)
//   (4) W-phase    X = Y T,  S = T^T (U^T X),  W = X - U S / 2;  V <- U.
//
// Pass decomposition: index blocks of 32; warp w owns block w.  At step s = 0..nb/2 warp w works
// on the block pair {w, (w+s) mod nb} (stored tile = rows max, columns min), so in every step all
// row blocks and all column blocks in flight are distinct: the sums for the *own* block stay in
// registers for the whole pass, the sums for the *other* block are added to a shared vector that
// no other warp touches during that step (no atomics, no per-warp scratch).  A 32x32 tile is two
// 16x32 sub-tiles; a lane holds 4 rows x 4 columns (lr = lane>>3 picks rows {2lr,2lr+1,2lr+8,2lr+9},
// lc = lane&7 picks columns {2lc,2lc+1,2lc+16,2lc+17}): every operand load is a conflict-free
// LDS.128 that 4 or 8 lanes share, 112 shared-memory wavefronts per 256 DFMA warp-instructions.
//
// Rows that do not fit into the CTA's shared memory (two CTAs per SM) stay in the global
// workspace and are updated in place through L2 -- once per panel instead of once per column.
#pragma once

#include "band_tridiag.cuh"
#include "common.cuh"
#include "tridiag_fused.cuh"  // poff, fast_rcp, fast_rsqrt, shfl_xor_d

namespace vsp {

constexpr size_t kSbrSmemBudget = 113 * 1024;  // two CTAs per SM
__host__ __device__ inline int sbr_warps(int n) {
    int nw = (n + 31) / 32;
    return nw < 1 ? 1 : nw;
}
// fixed scratch (doubles): V W U Yown Yoth(=P) [4][st] | T 16 | tau 4 | Zpart [nw][16] | pad 12
__host__ __device__ inline size_t sbr_fixed_doubles(int st, int nw) { return (size_t)20 * st + 32 + (size_t)16 * nw; }
constexpr int kSbrPad = 64;  // doubles after the last shared-memory row (masked lanes never read, but keep 16-byte slack)
__host__ __device__ inline size_t sbr_smem_bytes(int rows_smem, int st, int nw) {
    return sizeof(double) * (sbr_fixed_doubles(st, nw) + (size_t)poff(rows_smem) + kSbrPad);
}
// rows [0, rows_smem) of the triangle live in shared memory: all of them if they fit, else a multiple
// of 16 so that a sub-tile never straddles the two address spaces
__host__ __device__ inline int sbr_rows_in_smem(int n, int st, int nw, size_t budget) {
    if (sbr_smem_bytes(n, st, nw) <= budget) return n;
    int r = n & ~15;
    while (r > 0 && sbr_smem_bytes(r, st, nw) > budget) r -= 16;
    return r;
}

#if defined(__CUDACC__)

// 16 values spread over the 8 lanes that differ in lane bits 0..2 -> lane lc keeps the sums 2lc, 2lc+1
__device__ __forceinline__ void fold16_lc(const double (&a)[16], int lane, double& o0, double& o1) {
    double b[8], c[4];
    const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double send = b2 ? a[j] : a[j + 8], keep = b2 ? a[j + 8] : a[j];
        b[j] = keep + shfl_xor_d(send, 4);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double send = b1 ? b[j] : b[j + 4], keep = b1 ? b[j + 4] : b[j];
        c[j] = keep + shfl_xor_d(send, 2);
    }
    {
        const double send = b0 ? c[0] : c[2], keep = b0 ? c[2] : c[0];
        o0 = keep + shfl_xor_d(send, 1);
    }
    {
        const double send = b0 ? c[1] : c[3], keep = b0 ? c[3] : c[1];
        o1 = keep + shfl_xor_d(send, 1);
    }
}
// 16 values spread over the 4 lanes that differ in lane bits 3..4 -> lane lr keeps the sums 4lr..4lr+3
__device__ __forceinline__ void fold16_lr(const double (&a)[16], int lane, double (&o)[4]) {
    double b[8];
    const bool b4 = lane & 16, b3 = lane & 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double send = b4 ? a[j] : a[j + 8], keep = b4 ? a[j + 8] : a[j];
        b[j] = keep + shfl_xor_d(send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double send = b3 ? b[j] : b[j + 4], keep = b3 ? b[j + 4] : b[j];
        o[j] = keep + shfl_xor_d(send, 8);
    }
}

// One 16x32 sub-tile of the fused pass.  Rows R0 + {2lr, 2lr+1, 2lr+8, 2lr+9} (slots 0..3), columns
// C0 + {2lc, 2lc+1} (half 0) and C0 + 16 + {2lc, 2lc+1} (half 1).  rowacc[slot*4 + k],
// colacc[(2*half + xy)*4 + k].  Kinds (compile time, so interior tiles carry no masks at all):
//   kTileFull  off-diagonal, all 16 rows < p0
//   kTileRows  off-diagonal, last row block: rows >= p0 masked
//   kTileDiag0 diagonal block, upper sub-tile (R0 == C0): only half 0 exists, triangular mask
//   kTileDiag1 diagonal block, lower sub-tile (R0 == C0 + 16): half 0 full, half 1 triangular
// On diagonal tiles elements with column > row do not exist and the diagonal enters the row sums only.
enum { kTileFull = 0, kTileRows = 1, kTileDiag0 = 2, kTileDiag1 = 3 };

template <int KIND, bool DO_UPD, bool DO_SYM>
__device__ __forceinline__ void sbr_subtile(double* __restrict__ base, int R0, int C0, int p0,
                                            const double* __restrict__ Vb, const double* __restrict__ Wb,
                                            const double* __restrict__ Ub, int st, int lr, int lc,
                                            double (&rowacc)[16], double (&colacc)[16]) {
    constexpr int NH = (KIND == kTileDiag0) ? 1 : 2;     // halves that exist
    constexpr bool ROWMASK = (KIND != kTileFull);        // rows may be >= p0
    constexpr int TRI = (KIND == kTileDiag0) ? 0 : ((KIND == kTileDiag1) ? 1 : -1);  // half with the triangular mask
    const int rbase = R0 + 2 * lr, cbase = C0 + 2 * lc;
    const int rr[4] = {rbase, rbase + 1, rbase + 8, rbase + 9};
    double2 a[4][NH];
    bool ok[4][NH];  // the double2 (columns cbase + 16h, +1) of row slot s is stored
    double* rowp[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        rowp[s] = base + poff(rr[s]) + cbase;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            ok[s][h] = (!ROWMASK || rr[s] < p0) && (h != TRI || cbase + 16 * h <= rr[s]);
            if (ROWMASK || h == TRI)
                a[s][h] = ok[s][h] ? *reinterpret_cast<const double2*>(rowp[s] + 16 * h) : make_double2(0.0, 0.0);
            else
                a[s][h] = *reinterpret_cast<const double2*>(rowp[s] + 16 * h);
        }
    }
    // ---- pending rank-8 update (V = W = 0 before the first panel)
    if (DO_UPD) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 vr0 = *reinterpret_cast<const double2*>(Vb + k * st + rbase);
        const double2 vr1 = *reinterpret_cast<const double2*>(Vb + k * st + rbase + 8);
        const double2 wr0 = *reinterpret_cast<const double2*>(Wb + k * st + rbase);
        const double2 wr1 = *reinterpret_cast<const double2*>(Wb + k * st + rbase + 8);
        const double vr[4] = {vr0.x, vr0.y, vr1.x, vr1.y};
        const double wr[4] = {wr0.x, wr0.y, wr1.x, wr1.y};
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const double2 vc = *reinterpret_cast<const double2*>(Vb + k * st + cbase + 16 * h);
            const double2 wc = *reinterpret_cast<const double2*>(Wb + k * st + cbase + 16 * h);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                a[s][h].x = fma(-wr[s], vc.x, fma(-vr[s], wc.x, a[s][h].x));
                a[s][h].y = fma(-wr[s], vc.y, fma(-vr[s], wc.y, a[s][h].y));
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            if (ROWMASK || h == TRI) {
                if (ok[s][h]) *reinterpret_cast<double2*>(rowp[s] + 16 * h) = a[s][h];
            } else {
                *reinterpret_cast<double2*>(rowp[s] + 16 * h) = a[s][h];
            }
        }
    }
    // ---- products with the new reflectors: row sums (columns <= row), column sums (columns < row)
    if (DO_SYM) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 ur0 = *reinterpret_cast<const double2*>(Ub + k * st + rbase);
        const double2 ur1 = *reinterpret_cast<const double2*>(Ub + k * st + rbase + 8);
        const double ur[4] = {ur0.x, ur0.y, ur1.x, ur1.y};
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const double2 uc = *reinterpret_cast<const double2*>(Ub + k * st + cbase + 16 * h);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                double2 ar = a[s][h], ac = a[s][h];
                if (h == TRI) {
                    const int c = cbase + 16 * h;
                    ar.x = (ok[s][h] && c <= rr[s]) ? a[s][h].x : 0.0;
                    ar.y = (ok[s][h] && c + 1 <= rr[s]) ? a[s][h].y : 0.0;
                    ac.x = (ok[s][h] && c < rr[s]) ? a[s][h].x : 0.0;
                    ac.y = (ok[s][h] && c + 1 < rr[s]) ? a[s][h].y : 0.0;
                } else if (ROWMASK) {
                    ar.x = ac.x = ok[s][h] ? a[s][h].x : 0.0;
                    ar.y = ac.y = ok[s][h] ? a[s][h].y : 0.0;
                }
                rowacc[s * 4 + k] = fma(ar.x, uc.x, fma(ar.y, uc.y, rowacc[s * 4 + k]));
                colacc[(2 * h) * 4 + k] = fma(ac.x, ur[s], colacc[(2 * h) * 4 + k]);
                colacc[(2 * h + 1) * 4 + k] = fma(ac.y, ur[s], colacc[(2 * h + 1) * 4 + k]);
            }
        }
    }
    }
}

// dispatch on the kind of the sub-tile and on its address space (rows < rows_smem: shared, else the
// global workspace, in place)
template <bool DO_UPD, bool DO_SYM>
__device__ __forceinline__ void sbr_subtile_at(double* __restrict__ A, double* __restrict__ G, int rows_smem, bool diag,
                                               int sub, int R0, int C0, int p0, const double* Vb, const double* Wb,
                                               const double* Ub, int st, int lr, int lc, double (&rowacc)[16],
                                               double (&colacc)[16]) {
#define VSP_TILE_ARGS R0, C0, p0, Vb, Wb, Ub, st, lr, lc, rowacc, colacc
    if (R0 < rows_smem) {
        if (diag) {
            if (sub == 0) sbr_subtile<kTileDiag0, DO_UPD, DO_SYM>(A, VSP_TILE_ARGS);
            else sbr_subtile<kTileDiag1, DO_UPD, DO_SYM>(A, VSP_TILE_ARGS);
        } else if (R0 + 16 <= p0) {
            sbr_subtile<kTileFull, DO_UPD, DO_SYM>(A, VSP_TILE_ARGS);
        } else {
            sbr_subtile<kTileRows, DO_UPD, DO_SYM>(A, VSP_TILE_ARGS);
        }
    } else {
        if (diag) {
            if (sub == 0) sbr_subtile<kTileDiag0, DO_UPD, DO_SYM>(G, VSP_TILE_ARGS);
            else sbr_subtile<kTileDiag1, DO_UPD, DO_SYM>(G, VSP_TILE_ARGS);
        } else if (R0 + 16 <= p0) {
            sbr_subtile<kTileFull, DO_UPD, DO_SYM>(G, VSP_TILE_ARGS);
        } else {
            sbr_subtile<kTileRows, DO_UPD, DO_SYM>(G, VSP_TILE_ARGS);
        }
    }
#undef VSP_TILE_ARGS
}

// update-only sweep over the sub-tiles of the leading p0 x p0 triangle: sub-tile number t (row-major over
// 16-row blocks) is taken by worker (t mod nworkers)
__device__ __forceinline__ void sbr_update_sweep(double* __restrict__ A, double* __restrict__ G, int rows_smem, int p0,
                                                 int worker, int nworkers, const double* Vb, const double* Wb, int st,
                                                 int lr, int lc) {
    double dummy_r[16], dummy_c[16];
    int t = 0;
    const int nrb = (p0 + 15) >> 4;
    for (int i16 = 0; i16 < nrb; ++i16) {
        const int jd = i16 >> 1;  // diagonal column block of this row block
        for (int j = 0; j <= jd; ++j, ++t) {
            if (t % nworkers != worker) continue;
            sbr_subtile_at<true, false>(A, G, rows_smem, j == jd, i16 & 1, 16 * i16, 32 * j, p0, Vb, Wb, Vb, st, lr, lc,
                                        dummy_r, dummy_c);
        }
    }
}

// NQ = 32-column chunks a lane of the LQ warp holds (= max warps); MINB = CTAs per SM
template <int NQ, int MINB>
__global__ void __launch_bounds__(32 * NQ, MINB)
    sbr_band_kernel(const ItemDesc* __restrict__ items, int item_base, double* __restrict__ ws, int st, int rows_smem) {
    extern __shared__ __align__(16) double smem[];
    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    // warp index broadcast from lane 0: the compiler then knows it is warp-uniform, so the `warp == 0` /
    // `warp < nb` regions are uniform branches and the shuffles inside them are plain SHFLs
    const int lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), NW = nthreads >> 5;
    const int lr = lane >> 3, lc = lane & 7;

    double* Vb = smem;            // [4][st] previous reflectors (pending update)
    double* Wb = Vb + 4 * st;     // [4][st] previous W
    double* Ub = Wb + 4 * st;     // [4][st] current reflectors
    double* Yown = Ub + 4 * st;   // [4][st] sums kept by the owner of an index block
    double* Yoth = Yown + 4 * st; // [4][st] sums added by the partner warps; doubles as the panel buffer P
    double* Tm = Yoth + 4 * st;   // [16] T, row-major, upper triangular
    double* tauv = Tm + 16;       // [4] (+12 pad)
    double* Zpart = tauv + 16;    // [NW][16]
    double* A = Zpart + 16 * NW;  // rows < rows_smem (+ kSbrPad)

    // ---- load + condition the Gram matrix (same rules as tridiag_fused_kernel)
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
    if (flags) {  // uniform over the CTA
        for (int i = tid; i < 2 * n; i += nthreads) out[i] = 0.0;
        if (tid == 0) {
            out[2 * n + MISC_SCALE] = 1.0;
            out[2 * n + MISC_FLAGS] = (double)flags;
            out[2 * n + MISC_SLOT] = -1.0;
        }
        return;
    }
    if (tid == 0) {
        out[2 * n + MISC_SCALE] = scale;
        out[2 * n + MISC_FLAGS] = 0.0;
        out[2 * n + MISC_SLOT] = -1.0;
    }
    const int rs_eff = rows_smem < n ? rows_smem : n;
    const int total = poff(n), in_smem = poff(rs_eff);
    for (int i = tid; i < in_smem; i += nthreads) A[i] = G[i] * scale;
    for (int i = in_smem + tid; i < total; i += nthreads) G[i] *= scale;  // rows >= rows_smem: in place
    for (int i = tid; i < kSbrPad; i += nthreads) A[in_smem + i] = 0.0;
    for (int i = tid; i < 20 * st; i += nthreads) Vb[i] = 0.0;
    __syncthreads();

#ifdef VSP_PHASE_TIMING  // per-phase cycle counters of this warp (development builds only)
    long long t_phase[6] = {0, 0, 0, 0, 0, 0};  // mini | LQ (+wait) | pass compute | pass barriers | W-phase | panels
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
    bool pending = false;
    int m = n;
    while (m >= kSbrB + 2) {
        const int p0 = m - kSbrB;
        // ---- (1) mini-pass: panel rows p0..m-1, one column per thread
        double* P = Yoth;
        for (int c = tid; c < st; c += nthreads) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int r = p0 + t;
                double a = 0.0;
                if (c <= r) {
                    a = (r < rs_eff) ? A[poff(r) + c] : G[poff(r) + c];
                    if (pending) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            a = fma(-Wb[k * st + r], Vb[k * st + c], fma(-Vb[k * st + r], Wb[k * st + c], a));
                    }
                    if (c >= p0) G[poff(r) + c] = a;  // diagonal block: final band entries
                }
                P[t * st + c] = (c < p0) ? a : 0.0;
            }
        }
        __syncthreads();
        VSP_LAP(0);

        // ---- (2) LQ of the panel by warp 0, while the other warps apply the pending update to the
        //      leading p0 x p0 triangle (the update does not depend on the new reflectors)
        if (warp != 0) {
            if (pending) sbr_update_sweep(A, G, rs_eff, p0, warp - 1, NW - 1, Vb, Wb, st, lr, lc);
        } else {
            double Pr[4][NQ], Uk[4][NQ];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    Pr[t][q] = (32 * q < st) ? P[t * st + lane + 32 * q] : 0.0;
                    Uk[t][q] = 0.0;
                }
            double tk[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = 3 - k;
                const int pc = p0 - 4 + t;  // pivot column of row t
                if (pc >= 0) {              // warp-uniform
                    // partial sums over columns < pc: |x|^2 and the products with the rows above
                    double red[4] = {0.0, 0.0, 0.0, 0.0};  // [0..t-1]: rows t2 < t, [3]: |x|^2   (t <= 3)
                    double piv[4] = {0.0, 0.0, 0.0, 0.0};  // P[t2][pc], t2 <= t, picked from the owner lane
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int c = lane + 32 * q;
                        const double x = (c < pc) ? Pr[t][q] : 0.0;
                        red[3] = fma(x, x, red[3]);
#pragma unroll
                        for (int t2 = 0; t2 < 3; ++t2)
                            if (t2 < t) red[t2] = fma(Pr[t2][q], x, red[t2]);
                        if (c == pc) {
#pragma unroll
                            for (int t2 = 0; t2 < 4; ++t2)
                                if (t2 <= t) piv[t2] = Pr[t2][q];
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            red[i] += shfl_xor_d(red[i], o);
                            piv[i] += shfl_xor_d(piv[i], o);  // one non-zero contribution: a broadcast
                        }
                    }
                    const double xnorm2 = red[3], alpha = piv[t];
                    double beta = alpha, tau = 0.0, vscale = 0.0;
                    if (xnorm2 > 0.0) {
                        const double s2 = fma(alpha, alpha, xnorm2);
                        if (s2 > 1e-280) {
                            const double rs = fast_rsqrt(s2);
                            const double nrm = s2 * rs;
                            beta = -copysign(nrm, alpha);
                            tau = fma(fabs(alpha), rs, 1.0);
                            vscale = copysign(fast_rcp(fabs(alpha) + nrm), alpha);
                        } else {
                            beta = -copysign(sqrt(s2), alpha);
                            tau = (beta - alpha) / beta;
                            vscale = 1.0 / (alpha - beta);
                        }
                    }
                    tk[k] = tau;
                    double coef[3] = {0.0, 0.0, 0.0};
#pragma unroll
                    for (int t2 = 0; t2 < 3; ++t2)
                        if (t2 < t) coef[t2] = tau * fma(vscale, red[t2], piv[t2]);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int c = lane + 32 * q;
                        double u = 0.0;
                        if (tau != 0.0) u = (c < pc) ? Pr[t][q] * vscale : ((c == pc) ? 1.0 : 0.0);
                        Uk[k][q] = u;
#pragma unroll
                        for (int t2 = 0; t2 < 3; ++t2)
                            if (t2 < t) Pr[t2][q] = fma(-coef[t2], u, Pr[t2][q]);
                        if (c < pc) Pr[t][q] = 0.0;
                        if (c == pc) Pr[t][q] = beta;
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (32 * q < st) Ub[k * st + lane + 32 * q] = Uk[k][q];
            }
            // R block -> band output (rows p0+t, columns max(p0-4+t, 0) .. p0-1)
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int c = lane + 32 * q;
                    if (c < p0 && c >= p0 - 4 + t) G[poff(p0 + t) + c] = Pr[t][q];
                }
            // T from U^T U (forward recurrence in application order)
            double z[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // (0,1) (0,2) (1,2) (0,3) (1,3) (2,3)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                z[0] = fma(Uk[0][q], Uk[1][q], z[0]);
                z[1] = fma(Uk[0][q], Uk[2][q], z[1]);
                z[2] = fma(Uk[1][q], Uk[2][q], z[2]);
                z[3] = fma(Uk[0][q], Uk[3][q], z[3]);
                z[4] = fma(Uk[1][q], Uk[3][q], z[4]);
                z[5] = fma(Uk[2][q], Uk[3][q], z[5]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int i = 0; i < 6; ++i) z[i] += shfl_xor_d(z[i], o);
            if (lane == 0) {
                const double t00 = tk[0], t11 = tk[1], t22 = tk[2], t33 = tk[3];
                const double t01 = -t11 * (t00 * z[0]);
                const double t02 = -t22 * fma(t01, z[2], t00 * z[1]);
                const double t12 = -t22 * (t11 * z[2]);
                const double t03 = -t33 * fma(t02, z[5], fma(t01, z[4], t00 * z[3]));
                const double t13 = -t33 * fma(t12, z[5], t11 * z[4]);
                const double t23 = -t33 * (t22 * z[5]);
                Tm[0] = t00; Tm[1] = t01; Tm[2] = t02; Tm[3] = t03;
                Tm[4] = 0.0; Tm[5] = t11; Tm[6] = t12; Tm[7] = t13;
                Tm[8] = 0.0; Tm[9] = 0.0; Tm[10] = t22; Tm[11] = t23;
                Tm[12] = 0.0; Tm[13] = 0.0; Tm[14] = 0.0; Tm[15] = t33;
            }
            if (NW == 1 && pending) sbr_update_sweep(A, G, rs_eff, p0, 0, 1, Vb, Wb, st, lr, lc);
        }
        __syncthreads();
        VSP_LAP(1);

        // ---- (3) fused pass over the leading p0 x p0 triangle
        const int nb = (p0 + 31) >> 5;
        {
            double colacc[16];
            double ownR0 = 0.0, ownR1 = 0.0, ownR2 = 0.0, ownR3 = 0.0;  // folded row sums of the own block
            bool flushed = false;
#pragma unroll
            for (int i = 0; i < 16; ++i) colacc[i] = 0.0;
            const int nsteps = (nb >> 1) + 1;
            for (int s = 0; s < nsteps; ++s) {
                if (warp < nb) {
                    int o = warp + s;
                    if (o >= nb) o -= nb;
                    const bool half = (2 * s == nb);   // the pair {w, w + nb/2} is met from both sides
                    const bool own_cols = (s == 0) || (o > warp);
                    const int rb = own_cols ? o : warp, cb = own_cols ? warp : o;  // stored tile (rb, cb)
                    if (!own_cols && !flushed) {
                        // the own block's column sums are complete: fold and publish them
                        double f[4];
                        fold16_lr(colacc, lane, f);
                        const int c = 32 * warp + 2 * lc + (lr & 1) + 16 * (lr >> 1);
#pragma unroll
                        for (int k = 0; k < 4; ++k) Yown[k * st + c] = f[k];
#pragma unroll
                        for (int i = 0; i < 16; ++i) colacc[i] = 0.0;
                        flushed = true;
                    }
                    if (!(s > 0 && o == warp)) {  // nb == 1 and s > 0 cannot happen (nsteps == 1), guard anyway
#pragma unroll 1
                        for (int sub = 0; sub < 2; ++sub) {
                            if (half && sub != (own_cols ? 0 : 1)) continue;
                            const int R0 = 32 * rb + 16 * sub, C0 = 32 * cb;
                            if (R0 >= p0) continue;
                            double rowacc[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) rowacc[i] = 0.0;
                            sbr_subtile_at<false, true>(A, G, rs_eff, s == 0, sub, R0, C0, p0, Vb, Wb, Ub, st, lr, lc, rowacc, colacc);
                            double f0, f1;
                            fold16_lc(rowacc, lane, f0, f1);  // row slot lc>>1, k = 2(lc&1), 2(lc&1)+1
                            if (own_cols) {
                                const int slot = lc >> 1;
                                const int r = R0 + 2 * lr + (slot & 1) + 8 * (slot >> 1);
                                const int k0 = 2 * (lc & 1);
                                if (s == 0) {  // first touch of this block's partner sums: plain store
                                    Yoth[k0 * st + r] = f0;
                                    Yoth[(k0 + 1) * st + r] = f1;
                                } else {
                                    Yoth[k0 * st + r] += f0;
                                    Yoth[(k0 + 1) * st + r] += f1;
                                }
                            } else if (sub == 0) {
                                ownR0 += f0;
                                ownR1 += f1;
                            } else {
                                ownR2 += f0;
                                ownR3 += f1;
                            }
                        }
                        if (!own_cols) {  // column sums of the partner block
                            double f[4];
                            fold16_lr(colacc, lane, f);
                            const int c = 32 * cb + 2 * lc + (lr & 1) + 16 * (lr >> 1);
#pragma unroll
                            for (int k = 0; k < 4; ++k) Yoth[k * st + c] += f[k];
#pragma unroll
                            for (int i = 0; i < 16; ++i) colacc[i] = 0.0;
                        }
                    }
                }
                VSP_LAP(2);
                __syncthreads();
                VSP_LAP(3);
            }
            if (warp < nb) {
                if (!flushed) {
                    double f[4];
                    fold16_lr(colacc, lane, f);
                    const int c = 32 * warp + 2 * lc + (lr & 1) + 16 * (lr >> 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) Yown[k * st + c] = f[k];
                }
                __syncwarp();
                const int slot = lc >> 1, k0 = 2 * (lc & 1);
                const int r = 32 * warp + 2 * lr + (slot & 1) + 8 * (slot >> 1);
                Yown[k0 * st + r] += ownR0;
                Yown[(k0 + 1) * st + r] += ownR1;
                Yown[k0 * st + r + 16] += ownR2;
                Yown[(k0 + 1) * st + r + 16] += ownR3;
            }
        }
        __syncthreads();

        // ---- (4) W-phase: X = Y T, Z = U^T X (block reduction), S = T^T Z, W = X - U S / 2
        {
            double T[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) T[i] = Tm[i];
            const int r = tid;
            double x[4] = {0.0, 0.0, 0.0, 0.0}, u[4] = {0.0, 0.0, 0.0, 0.0};
            if (r < p0) {
                double y[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    y[k] = Yown[k * st + r] + Yoth[k * st + r];
                    u[k] = Ub[k * st + r];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k = 0; k <= j; ++k) x[j] = fma(y[k], T[k * 4 + j], x[j]);
            }
            double zp[16];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) zp[i * 4 + j] = u[i] * x[j];
            // fold over the 32 lanes: after four levels lane l holds entry (l >> 1), then pair-sum
            double b8[8], b4[4], b2[2], b1;
            {
                const bool t4 = lane & 16, t3 = lane & 8, t2 = lane & 4, t1 = lane & 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double send = t4 ? zp[j] : zp[j + 8], keep = t4 ? zp[j + 8] : zp[j];
                    b8[j] = keep + shfl_xor_d(send, 16);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double send = t3 ? b8[j] : b8[j + 4], keep = t3 ? b8[j + 4] : b8[j];
                    b4[j] = keep + shfl_xor_d(send, 8);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double send = t2 ? b4[j] : b4[j + 2], keep = t2 ? b4[j + 2] : b4[j];
                    b2[j] = keep + shfl_xor_d(send, 4);
                }
                {
                    const double send = t1 ? b2[0] : b2[1], keep = t1 ? b2[1] : b2[0];
                    b1 = keep + shfl_xor_d(send, 2);
                }
                b1 += shfl_xor_d(b1, 1);
            }
            if ((lane & 1) == 0) Zpart[warp * 16 + (lane >> 1)] = b1;
            __syncthreads();
            // every warp rebuilds S = T^T Z in its lanes 0..15 (lane = 4i + j) and broadcasts it with shuffles
            double ze = 0.0;
            if (lane < 16)
                for (int w = 0; w < NW; ++w) ze += Zpart[w * 16 + lane];
            double se = 0.0;
            {
                const int i = (lane >> 2) & 3, j = lane & 3;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double zk = __shfl_sync(0xffffffffu, ze, 4 * k + j);
                    if (k <= i) se = fma(Tm[k * 4 + i], zk, se);
                }
            }
            double wv[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double sik = __shfl_sync(0xffffffffu, se, 4 * i + k);
                    wv[k] = fma(-0.5 * u[i], sik, wv[k]);
                }
            if (r < st) {
#pragma unroll
                for (int k = 0; k < 4; ++k) Wb[k * st + r] = (r < p0) ? wv[k] : 0.0;
            }
        }
        // V <- U: swap the two buffers (U is rewritten completely by the next LQ)
        {
            double* tmp = Vb;
            Vb = Ub;
            Ub = tmp;
        }
        pending = true;
        m = p0;
        __syncthreads();
        VSP_LAP(4);
#ifdef VSP_PHASE_TIMING
        t_phase[5] += 1;
#endif
    }
#ifdef VSP_PHASE_TIMING
    if (blockIdx.x == 200 && lane == 0)
        printf("[sbr n=%d warp %d] cycles: mini %lld  LQ %lld  pass %lld  pass-barrier %lld  W %lld  panels %lld\n", n, warp,
               t_phase[0], t_phase[1], t_phase[2], t_phase[3], t_phase[4], t_phase[5]);
#endif
#undef VSP_LAP
    // ---- remaining rows (m <= 5): inside the band; bring them up to date and publish
    for (int i = tid; i < m * 8; i += nthreads) {
        const int r = i >> 3, c = i & 7;
        if (c <= r) {
            double a = (r < rs_eff) ? A[poff(r) + c] : G[poff(r) + c];
            if (pending) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a = fma(-Wb[k * st + r], Vb[k * st + c], fma(-Vb[k * st + r], Wb[k * st + c], a));
            }
            G[poff(r) + c] = a;
        }
    }
}

#endif  // __CUDACC__

}  // namespace vsp
