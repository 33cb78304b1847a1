// refine_cluster.cuh -- the ill-conditioned re-solve with one thread-block CLUSTER per flagged matrix.
//
// Same algorithm as refine_bidiag.cuh -- Golub-Kahan Householder bidiagonalisation of the FP64 copy of W (the route
// LAPACK's dgesdd takes behind spectral.py:91) followed by bisection on the Golub-Kahan tridiagonal form -- but the
// K x n matrix is distributed column-cyclically over the B CTAs of a cluster (B SMs work on one matrix), because the
// one-CTA version is a 2.9 ms tail at n = 192 and 150 ms per matrix at n = 768 behind a 1.8 ms bisection kernel
// (round-1 profile).  Column c lives in CTA c mod B, in shared memory when the CTA's share fits, else in the slot's
// global (L2-resident) pool.  One step j needs ONE cluster barrier:
//   every CTA (redundantly) left reflector of column j: all CTAs hold that column (it was published one step earlier)
//   every CTA              applies it to its columns (one warp per column; dot products stay inside the CTA), keeps
//                          the new row j, and publishes: its share of |row j|^2, the pivot entry, its PARTIAL
//                          y' = sum_c X[:, c] row_j[c] over its columns c >= j+2 (unscaled: the right reflector's
//                          scalars are not known yet) and, from its owner, column j+1                 cluster.sync
//   every CTA (redundantly) right reflector scalars, y = tau2 (us sum_b y'_b + column j+1), the rank-one update of its
//                          columns, and the updated column j+1 = column j+1 - y: the next step's left reflector.
// The exchange buffers live in global memory (the pool slot, double-buffered by step parity): cluster.sync is a
// release/acquire barrier at cluster scope, so plain stores before it are visible to plain loads after it; the cluster
// is only used for its hardware barrier and its co-scheduling guarantee (no distributed shared memory).  Rank 0
// finishes alone: bisection on the Golub-Kahan form, metrics, record.
#pragma once

#include <cooperative_groups.h>

#include "bisect_metrics.cuh"
#include "refine_bidiag.cuh"

namespace vsp {

constexpr int kRcThreads = 512;
constexpr int kRcMaxCluster = 16;  // above 8: non-portable cluster size (opt-in per kernel; one GPC holds 16+ SMs on B200)

// pool slot layout (doubles): X [B shares of ceil(n / B) columns][K] | 2 x exchange { scal [B][8] | ypart [B][K] | col [K] }
__host__ __device__ inline int64_t refine_cluster_x_doubles(int64_t K, int64_t n, int B) { return K * (((n + B - 1) / B) * B); }
__host__ __device__ inline int64_t refine_cluster_ex_doubles(int64_t K, int B) { return 8 * (int64_t)B + (int64_t)B * K + K; }
__host__ __device__ inline int64_t refine_cluster_slot_doubles(int64_t K, int64_t n, int B) {
    return refine_cluster_x_doubles(K, n, B) + 2 * refine_cluster_ex_doubles(K, B) + 8;
}
// shared memory (doubles): reduction scratch | dq eq [npad] | v y [Kpad] | part [kRcThreads] | u rowj [nlocpad] | X share
// FUSED (K <= 32 RPL): the right update of step j-1 and the left reflector of step j are ONE sweep over the CTA's
// columns (a column's rows live in RPL registers per lane between its load and its store) instead of two sweeps with
// four passes (n = 768: 14.8 -> 11.6 ms per matrix; n = 192 on the exclusive entry point: 1.5 -> 1.3 ms).  The entry
// point that runs beside the bisection kernel takes it with an 80-register cap, together with a 64-register bisection
// kernel: every register of the re-solve CTA is bisection work that does not fit on its SM (108 of 148 SMs host a
// re-solve CTA with the bench's 36 flagged matrices).  Stage 3 of the Scenario-A sweep, 36 / 28 flagged matrices:
//   re-solve 64 registers, three sweeps, bisection 80 registers   2.30 / 2.25 ms   (one sweep at 64: spills, 2.35)
//   re-solve 72..80, one sweep, bisection 80                      2.44 / 2.13
//   re-solve 80, one sweep, bisection 64                          2.23 / 2.11      <- this build
//   re-solve 96                                                   2.43 either way
// With y' folded in as well (FUSEY: per-warp partial y' through shared memory) it was slower everywhere it was tried.
constexpr int kRcRplWide = 24;  // K <= 768
__host__ __device__ inline size_t refine_cluster_ypw_doubles(bool shared_variant) {
    (void)shared_variant;
    return 0;  // FUSEY is off in both entry points
}
__host__ __device__ inline size_t refine_cluster_fixed_doubles(int npad, int Kpad, int nloc, bool shared_variant) {
    return (size_t)CtaCtx::kScratchDoubles + 2 * (size_t)npad + 2 * (size_t)Kpad + kRcThreads + 2 * (size_t)((nloc + 3) & ~3) + 8 +
           refine_cluster_ypw_doubles(shared_variant);
}
// rank 0 re-uses everything behind the reduction scratch for the bisection: lam [npad] | part | DE [2 npad] | counter
__host__ __device__ inline size_t refine_cluster_tail_doubles(int npad) { return (size_t)CtaCtx::kScratchDoubles + 7 * (size_t)npad + 16; }

#if defined(__CUDACC__)

// WIDE: the CTA's share of X lives in L2 (n > 256): the sweeps keep eight independent loads per lane in flight (an L2
// round trip is ~700 cycles; with the default unrolling a step of n = 768 moved 23 GB/s per SM)
template <typename TIn, bool WIDE, int RPL, bool FUSEY>
__device__ __forceinline__ void refine_cluster_body(const ItemDesc* __restrict__ items, RefineGate gate, RefinePool pool, int npad,
                                                    int Kpad, int nloc_max, int xs_cap, vsp_opts opts, double* __restrict__ sv_out,
                                                    vsp_record* __restrict__ records, double* __restrict__ dist_out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double smem[];
    const int B = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int slot = blockIdx.x / B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = kRcThreads / 32;

    double* red = smem;
    double* dq = red + CtaCtx::kScratchDoubles;
    double* eq = dq + npad;
    double* vs = eq + npad;       // [Kpad] left reflector of this step
    double* ys = vs + Kpad;       // [Kpad] y of this step
    double* part = ys + Kpad;     // [kRcThreads] column-group partial sums of the y' pass
    double* us = part + kRcThreads;  // [nloc] right reflector entries of the local columns
    double* rowj = us + ((nloc_max + 3) & ~3);  // [nloc] updated row j of the local columns
    double* ypw = rowj + ((nloc_max + 3) & ~3) + 8;  // [NWARP][32 RPL] partial y' of the warps (FUSEY)
    double* Xs = ypw + (FUSEY ? (size_t)(kRcThreads / 32) * 32 * RPL : 0);

    const int nflagged = *gate.counter;
    for (int entry = slot; entry < nflagged; entry += pool.slots) {  // uniform over the cluster
        const ItemDesc it = items[gate.slot_items[entry]];
        const int n = it.n, K = it.kdim;
        CtaCtx ctx(red);
        double* Xg = pool.base + (int64_t)slot * pool.slot_doubles;  // the slot's global copy: B shares, column-major
        double* ex0 = Xg + refine_cluster_x_doubles(K, n, B);         // exchange buffers, by step parity
        const int64_t exd = refine_cluster_ex_doubles(K, B);
        const TIn* __restrict__ W = reinterpret_cast<const TIn*>(it.ptr);
        const int nloc = (n - rank + B - 1) / B;  // local columns c = rank + B lc
        // local column lc starts at Xl + lc * K: in shared memory when this matrix's share fits (xs_cap doubles), else in
        // the CTA's share of the slot's global copy
        const int nshare = (n + B - 1) / B;
        double* Xl = ((int64_t)nshare * K <= (int64_t)xs_cap) ? Xs : Xg + (int64_t)rank * nshare * K;
        const bool fused = RPL > 0 && K <= 32 * RPL;  // uniform over the cluster
        bool pending = false;                         // fused: the right update of the previous step has not been applied yet

        // ---- scale by a power of two so that max |x| is in [0.5, 1): cluster-wide maximum
        double mx = 0.0;
        {
            const int64_t total = (int64_t)it.rows * it.cols;
            for (int64_t e = (int64_t)rank * kRcThreads + tid; e < total; e += (int64_t)B * kRcThreads) {
                const int r = (int)(e / it.cols), c = (int)(e % it.cols);
                mx = fmax(mx, fabs((double)W[(int64_t)r * it.ld + c]));
            }
        }
        mx = ctx.max(mx);
        if (tid == 0) ex0[8 * rank] = mx;
        cluster.sync();
        mx = 0.0;
        for (int b = 0; b < B; ++b) mx = fmax(mx, ex0[8 * b]);
        int ex = 0;
        (void)frexp(mx, &ex);
        const double sc = ldexp(1.0, -ex);
        // ---- local columns of X (Gram index on the columns); every CTA also takes column 0 (the first reflector)
        for (int64_t e = tid; e < (int64_t)nloc * K; e += kRcThreads) {
            const int lc = (int)(e / K), k = (int)(e % K);
            const int c = rank + B * lc;
            const double w = it.trans ? (double)W[(int64_t)k * it.ld + c] : (double)W[(int64_t)c * it.ld + k];
            Xl[(int64_t)lc * K + k] = w * sc;
        }
        for (int k = tid; k < K; k += kRcThreads) vs[k] = (it.trans ? (double)W[(int64_t)k * it.ld] : (double)W[k]) * sc;
        cluster.sync();  // everybody has read the maxima before step 0 rewrites the buffer; local columns complete

#ifdef VSP_PHASE_TIMING
        long long tph[6] = {0, 0, 0, 0, 0, 0};
        long long tmk = clock64();
#define VSP_RLAP(k) do { const long long now_ = clock64(); tph[k] += now_ - tmk; tmk = now_; } while (0)
#else
#define VSP_RLAP(k) ((void)0)
#endif
        int par = 0;
        for (int j = 0; j < n; ++j) {
            // ---- left reflector from column j (held by every CTA in vs[j..K)): annihilate X[j+1:K, j]
            double ss = 0.0;
            for (int r = j + 1 + tid; r < K; r += kRcThreads) ss += vs[r] * vs[r];
            const double xn2 = ctx.sum(ss);
            const double alpha = vs[j];
            double beta = alpha, tau = 0.0, vsc = 0.0;
            if (xn2 > 0.0) {
                beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
                tau = (beta - alpha) / beta;
                vsc = 1.0 / (alpha - beta);
            }
            if (tid == 0) dq[j] = beta;
            __syncthreads();  // everybody has read alpha
            for (int r = j + 1 + tid; r < K; r += kRcThreads) vs[r] *= vsc;
            __syncthreads();
            VSP_RLAP(0);
            // apply to the local columns c > j (one warp per column); collect the new row j and |row j|^2 (c >= j+2)
            const int lcf = (j + 1 - rank + B - 1) / B;  // first local column with c >= j + 1
            double s2 = 0.0;
            if (fused) {
                if constexpr (RPL > 0) {
                    // ---- one sweep: right update of step j-1 (pending), left reflector of step j, partial y' of step j.
                    //      Lane l, slot i holds row j + l + 32 i of the column between its load and its store.
                    double yacc[FUSEY ? RPL : 1];
#pragma unroll
                    for (int i = 0; i < (FUSEY ? RPL : 1); ++i) yacc[i] = 0.0;
                    for (int lc = lcf + warp; lc < nloc; lc += NWARP) {
                        const int c = rank + B * lc;
                        double* cc = Xl + (int64_t)lc * K;
                        double x[RPL];
#pragma unroll
                        for (int i = 0; i < RPL; ++i) {
                            const int r = j + lane + 32 * i;
                            x[i] = (r < K) ? cc[r] : 0.0;
                        }
                        if (pending) {
                            const double u = us[lc];
#pragma unroll
                            for (int i = 0; i < RPL; ++i) {
                                const int r = j + lane + 32 * i;
                                if (r < K) x[i] = fma(-u, ys[r], x[i]);
                            }
                        }
                        const double cj = __shfl_sync(0xffffffffu, x[0], 0);  // row j
                        double dot = 0.0;
#pragma unroll
                        for (int i = 0; i < RPL; ++i) {
                            const int r = j + lane + 32 * i;
                            if (r > j && r < K) dot = fma(vs[r], x[i], dot);
                        }
                        const double w = tau * (ctx.warp_sum(dot) + cj);
                        const double rj = cj - w;
                        const bool tail = c >= j + 2;  // uniform over the warp
#pragma unroll
                        for (int i = 0; i < RPL; ++i) {
                            const int r = j + lane + 32 * i;
                            if (r > j && r < K) {
                                x[i] = fma(-w, vs[r], x[i]);
                                cc[r] = x[i];
                                if constexpr (FUSEY)
                                    if (tail) yacc[i] = fma(x[i], rj, yacc[i]);
                            }
                        }
                        if (lane == 0) rowj[lc] = rj;
                        if (tail) s2 += rj * rj;
                    }
                    if constexpr (FUSEY) {
#pragma unroll
                        for (int i = 0; i < RPL; ++i)
                            if (lane + 32 * i < K - j) ypw[warp * (32 * RPL) + lane + 32 * i] = yacc[i];
                    }
                    pending = false;
                }
            } else
            for (int lc = lcf + warp; lc < nloc; lc += NWARP) {
                const int c = rank + B * lc;
                double* cc = Xl + (int64_t)lc * K;
                const double cj = cc[j];
                double w = 0.0;
                if (tau != 0.0) {
                    double dot = 0.0;
                    if constexpr (WIDE) {
                        double d[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                        int r = j + 1 + lane;
                        for (; r + 7 * 32 < K; r += 8 * 32) {
                            double x[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) x[q] = cc[r + 32 * q];
#pragma unroll
                            for (int q = 0; q < 8; ++q) d[q] = fma(vs[r + 32 * q], x[q], d[q]);
                        }
                        for (; r < K; r += 32) d[0] = fma(vs[r], cc[r], d[0]);
                        dot = ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7]));
                        w = tau * (ctx.warp_sum(dot) + cj);
                        r = j + 1 + lane;
                        for (; r + 7 * 32 < K; r += 8 * 32) {
                            double x[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) x[q] = cc[r + 32 * q];
#pragma unroll
                            for (int q = 0; q < 8; ++q) cc[r + 32 * q] = fma(-w, vs[r + 32 * q], x[q]);
                        }
                        for (; r < K; r += 32) cc[r] -= w * vs[r];
                    } else {
                        for (int r = j + 1 + lane; r < K; r += 32) dot += vs[r] * cc[r];
                        w = tau * (ctx.warp_sum(dot) + cj);
                        for (int r = j + 1 + lane; r < K; r += 32) cc[r] -= w * vs[r];
                    }
                }
                if (lane == 0) rowj[lc] = cj - w;
                if (c >= j + 2) s2 += (cj - w) * (cj - w);  // uniform over the warp
            }
            if (lane != 0) s2 = 0.0;
            s2 = ctx.sum(s2);  // (also a CTA barrier: rowj complete, local columns updated)
            if (j + 1 >= n) {
                if (tid == 0) eq[j] = 0.0;
                break;  // uniform over the cluster
            }
            VSP_RLAP(1);
            // ---- publish: |row j|^2 share, pivot entry, partial y' over the local columns c >= j+2, column j+1
            double* exb = ex0 + par * exd;
            double* ypart = exb + 8 * B;
            double* colx = ypart + (int64_t)B * K;
            const int own1 = (j + 1) % B;  // owner of column j + 1
            const int lc2 = (j + 2 - rank + B - 1) / B;  // first local column with c >= j + 2
            if (FUSEY && fused) {
                for (int sl = 1 + tid; sl < K - j; sl += kRcThreads) {  // slot sl <-> row j + sl
                    double y = 0.0;
#pragma unroll
                    for (int wq = 0; wq < NWARP; ++wq) y += ypw[wq * (32 * RPL) + sl];
                    ypart[(int64_t)rank * K + j + sl] = y;
                }
            } else {
                // rows r > j, split over G column groups so that every thread has work
                const int nrows = K - (j + 1);
                int G = kRcThreads / (nrows > 0 ? nrows : 1);
                if (G < 1) G = 1;
                const int ncl = nloc - lc2;
                if (G > ncl) G = ncl < 1 ? 1 : ncl;
                if (G == 1) {
                    for (int r = j + 1 + tid; r < K; r += kRcThreads) {
                        double y = 0.0;
                        if constexpr (WIDE) {
                            double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                            int lc = lc2;
                            for (; lc + 7 < nloc; lc += 8) {
                                double x[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q) x[q] = Xl[(int64_t)(lc + q) * K + r];
#pragma unroll
                                for (int q = 0; q < 8; ++q) a[q] = fma(x[q], rowj[lc + q], a[q]);
                            }
                            for (; lc < nloc; ++lc) a[0] = fma(Xl[(int64_t)lc * K + r], rowj[lc], a[0]);
                            y = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
                        } else {
                            for (int lc = lc2; lc < nloc; ++lc) y += Xl[(int64_t)lc * K + r] * rowj[lc];
                        }
                        ypart[(int64_t)rank * K + r] = y;
                    }
                } else {  // G * nrows <= kRcThreads
                    const int gq = tid / nrows, r = j + 1 + (tid - gq * nrows);
                    if (gq < G) {
                        double y = 0.0;
                        for (int lc = lc2 + gq; lc < nloc; lc += G) y += Xl[(int64_t)lc * K + r] * rowj[lc];
                        part[tid] = y;
                    }
                    __syncthreads();
                    if (gq == 0) {
                        double y = 0.0;
                        for (int q = 0; q < G; ++q) y += part[q * nrows + (r - (j + 1))];
                        ypart[(int64_t)rank * K + r] = y;
                    }
                }
            }
            if (rank == own1) {
                const double* c1 = Xl + (int64_t)((j + 1) / B) * K;
                for (int r = j + 1 + tid; r < K; r += kRcThreads) colx[r] = c1[r];
            }
            if (tid == 0) {
                exb[8 * rank] = s2;
                if (rank == own1) exb[8 * rank + 1] = rowj[(j + 1) / B];
            }
            VSP_RLAP(2);
            cluster.sync();  // the one barrier of the step
            VSP_RLAP(3);
            // ---- right reflector: annihilate X[j, j+2:n].  The loads of the y' shares and of column j+1 are issued first:
            //      they do not depend on the reflector's scalars, whose square root / divisions then overlap their latency
            for (int r = j + 1 + tid; r < K; r += kRcThreads) {
                double y = 0.0;
                for (int b = 0; b < B; ++b) y += ypart[(int64_t)b * K + r];
                ys[r] = y;
                vs[r] = colx[r];
            }
            double yn2 = 0.0;
            for (int b = 0; b < B; ++b) yn2 += exb[8 * b];
            const double a2 = exb[8 * own1 + 1];
            double b2 = a2, tau2 = 0.0, usc = 0.0;
            if (yn2 > 0.0) {
                b2 = -copysign(sqrt(a2 * a2 + yn2), a2);
                tau2 = (b2 - a2) / b2;
                usc = 1.0 / (a2 - b2);
            }
            if (tid == 0) eq[j] = b2;
            // y = tau2 (us sum_b y'_b + column j+1); the next step's column: column j+1 - y  (u_{j+1} = 1)
            for (int r = j + 1 + tid; r < K; r += kRcThreads) {  // same thread <-> row mapping as above: no barrier needed
                const double c1 = vs[r];
                const double y = tau2 * fma(usc, ys[r], c1);
                ys[r] = y;
                vs[r] = c1 - y;
            }
            for (int lc = lcf + tid; lc < nloc; lc += kRcThreads) us[lc] = (rank + B * lc == j + 1) ? 1.0 : rowj[lc] * usc;
            __syncthreads();
            VSP_RLAP(4);
            if (fused) {
                pending = tau2 != 0.0;  // applied by the next step's sweep
            } else if (tau2 != 0.0) {  // uniform over the cluster
                // rank-one update of the local columns c >= j+1, one warp per column
                for (int lc = lcf + warp; lc < nloc; lc += NWARP) {
                    double* cc = Xl + (int64_t)lc * K;
                    const double u = us[lc];
                    if constexpr (WIDE) {
                        int r = j + 1 + lane;
                        for (; r + 7 * 32 < K; r += 8 * 32) {
                            double x[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) x[q] = cc[r + 32 * q];
#pragma unroll
                            for (int q = 0; q < 8; ++q) cc[r + 32 * q] = fma(-u, ys[r + 32 * q], x[q]);
                        }
                        for (; r < K; r += 32) cc[r] -= ys[r] * u;
                    } else {
                        for (int r = j + 1 + lane; r < K; r += 32) cc[r] -= ys[r] * u;
                    }
                }
            }
            __syncthreads();
            VSP_RLAP(5);
            par ^= 1;
        }
#ifdef VSP_PHASE_TIMING
        if (slot == 0 && tid == 0)
            printf("[refine n=%d K=%d rank %d] cycles: reflector %lld  left-apply %lld  publish %lld  cluster.sync %lld  right-scalars+y %lld  update %lld\n",
                   n, K, rank, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5]);
#endif
#undef VSP_RLAP
        cluster.sync();  // the exchange buffers are free for the next entry; dq / eq complete in every CTA
        if (rank == 0) {
            // ---- singular values of the bidiagonal, metrics, record (this CTA alone; the shared arrays behind dq / eq
            //      are re-used: lam | DE)
            double* lam = eq + npad;
            DE* de = reinterpret_cast<DE*>(lam + npad);
            int iters = gk_singular_values(ctx, dq, eq, n, de, lam);
            iters = ctx.max_i(iters);  // barrier: lam[] complete
            double* sv = (opts.want_sv != 0 && sv_out != nullptr) ? sv_out + it.sv_off : nullptr;
            double* aux = dist_out != nullptr ? dist_out + (int64_t)it.item * VSP_AUX_STRIDE(opts.dist_k, opts.clauset) : nullptr;
            const MetricOut mo = spectral_metrics(ctx, lam, n, sc * sc, 0, opts.fit_start, opts.fit_end, opts.hill_k, sv, aux, opts.dist_k,
                                                  (aux != nullptr && opts.clauset) ? aux + VSP_AUX_STRIDE(opts.dist_k, 0) : nullptr);
            if (tid == 0) {
                vsp_record r;
                r.item = it.item;
                r.status = mo.status | VSP_ST_ILLCOND | VSP_ST_REFINED;
                r.m = mo.m;
                r.start = mo.start;
                r.end = mo.end;
                r.k = mo.k;
                r.n = n;
                r.iters = iters;
                r.metrics[0] = mo.metrics[0];
                r.metrics[1] = mo.metrics[1];
                r.metrics[2] = mo.metrics[2];
                r.metrics[3] = mo.metrics[3];
                records[it.item] = r;
            }
        }
        cluster.sync();  // rank 0's shared arrays are rewritten by the next entry's fill
    }
}

// Two entry points around the same body:
//   refine_cluster_kernel        n > 256: the re-solve is the long pole (29 ms per ViT-Base chunk against 2 ms of
//                                bisection), all registers, default carve-out (the L2-resident shares like the L1);
//   refine_cluster_shared_kernel n <= 256: runs BESIDE the bisection kernel, so it must leave room on its SMs: at most
//                                80 registers (124 x 512 threads took an SM's whole register file) and, set by the
//                                host, the maximum shared-memory carve-out (with the default split no bisection CTA
//                                fitted next to the 120 KB of a re-solve CTA).  Both were needed: stage 3 of the
//                                Scenario-A sweep 2.65 -> 2.22 ms although the re-solve itself slows from 1.5 to 2.1 ms.
template <typename TIn>
__global__ void __launch_bounds__(kRcThreads)
    refine_cluster_kernel(const ItemDesc* __restrict__ items, RefineGate gate, RefinePool pool, int npad, int Kpad, int nloc_max,
                          int xs_cap, vsp_opts opts, double* __restrict__ sv_out, vsp_record* __restrict__ records,
                          double* __restrict__ dist_out) {
    refine_cluster_body<TIn, true, kRcRplWide, false>(items, gate, pool, npad, Kpad, nloc_max, xs_cap, opts, sv_out, records, dist_out);
}
// n <= 256, small classes: all registers like the entry point above, but the one-sweep step sized for K <= 256 (the
// K <= 768 instantiation walks 24 register slots per lane of which a 192-row column fills six)
template <typename TIn>
__global__ void __launch_bounds__(kRcThreads)
    refine_cluster_small_kernel(const ItemDesc* __restrict__ items, RefineGate gate, RefinePool pool, int npad, int Kpad,
                                int nloc_max, int xs_cap, vsp_opts opts, double* __restrict__ sv_out,
                                vsp_record* __restrict__ records, double* __restrict__ dist_out) {
    refine_cluster_body<TIn, false, 8, false>(items, gate, pool, npad, Kpad, nloc_max, xs_cap, opts, sv_out, records, dist_out);
}
template <typename TIn>
__global__ void __maxnreg__(80)
    refine_cluster_shared_kernel(const ItemDesc* __restrict__ items, RefineGate gate, RefinePool pool, int npad, int Kpad,
                                 int nloc_max, int xs_cap, vsp_opts opts, double* __restrict__ sv_out,
                                 vsp_record* __restrict__ records, double* __restrict__ dist_out) {
    refine_cluster_body<TIn, false, 8, false>(items, gate, pool, npad, Kpad, nloc_max, xs_cap, opts, sv_out, records, dist_out);
}

#endif  // __CUDACC__

}  // namespace vsp
