// sbr_band.cuh -- stage 2a-1, n <= kSmemMaxN: blocked Householder reduction of the (scaled) Gram matrix to a
// symmetric band of bandwidth b = 4, in place, on the FP64 tensor cores.  band_tridiag.cuh finishes the job.
//
// Why blocked: an unblocked tridiagonalisation moves every stored element through shared memory once per
// *column* (4 DFMA per 16 bytes: shared-memory bound at <= 50 % of the FP64 pipe, and one block-wide dependency
// chain per column).  Here an element moves once per *panel of four columns*, all O(n^3) work is rank-8 updates
// and 4-column symmetric products, i.e. small GEMMs that run as DMMA.8x8x4 (37.1 TFLOP/s measured on B200 with
// one warp per SM sub-partition, scripts/micro/dmma_bench.cu), and the serial chain is per panel.
//
// Elimination order is bottom-up: the active matrix is the leading m x m block.  One panel step, p0 = m - 4:
//   (1) mini-pass  (all threads) rows p0..m-1 get the pending rank-8 update  A -= V W^T + W V^T;
//                  their diagonal block goes to the band output, the 4 x p0 block left of it to P.
//   (2) LQ         (warp 0) four Householder reflectors (row p0+3 first, pivot column p0-1, then
//                  p0+2 / p0-2, ...) reduce P to an upper-triangular 4x4 block R next to the diagonal
//                  block: Q = H_0 H_1 H_2 H_3 = I - U T U^T (compact WY, T from U^T U).
//       update     (warps 1..) meanwhile the pending update is applied to the leading p0 x p0 triangle: it does not
//                  need the new reflectors, so the one-warp LQ chain hides behind it.  Two DMMAs per 8x8 tile.
//   (3) products   (all warps) Y = A U over the leading p0 x p0 triangle, both triangles from one read of the
//                  stored one: four DMMAs per 8x8 tile (b = 4 fills half of the MMA's n = 8).
//   (4) W-phase    X = Y T,  S = T^T (U^T X),  W = X - U S / 2;  V <- U.
//
// Decomposition of (3): index blocks of 32; warp w owns block w.  At step s = 0..nb/2 warp w works on the block
// pair {w, (w+s) mod nb} (stored tile = rows max, columns min), so in every step all row blocks and all column
// blocks in flight are distinct: the sums for the *own* block stay in registers for the whole pass, the sums for
// the *other* block are added to a shared vector that no other warp touches during that step (no atomics, no
// per-warp scratch).  The stored values of an 8x8 tile are its C fragment (lane (g, t): row g, columns 2t, 2t+1,
// one 128-bit access); the same two registers are the A fragments of the k-slices {2t} and {2t+1} for the row
// sums, and four shuffles per slice give the transposed A fragments for the column sums.
//
// Rows that do not fit into the CTA's shared memory (two CTAs per SM while the active order is large) stay in
// the global workspace and are updated in place through L2 -- once per panel, with the next unit's tiles
// fetched while the current one is in the pipe.  The host launches the kernel per order range (n -> 96 -> 48 ->
// end, vspectra_api.cu): a smaller active block means a smaller CTA and more matrices per SM while the steps
// are latency-bound.
#pragma once

#include "band_tridiag.cuh"
#include "common.cuh"

namespace vsp {

constexpr size_t kSbrSmemBudget = 113 * 1024;  // two CTAs per SM
__host__ __device__ inline int sbr_warps(int n) {
    int nw = (n + 31) / 32;
    return nw < 1 ? 1 : nw;
}
// fixed scratch (doubles): V W U Y(=P) [4][st] | T 16 | tau 4 | Zpart [nw][16] | pad 12
__host__ __device__ inline size_t sbr_fixed_doubles(int st, int nw) { return (size_t)16 * st + 32 + (size_t)16 * nw; }
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

// FP64 tensor-core MMA (SASS: DMMA.8x8x4), D[8x8] += A[8x4] B[4x8].  Lane (g = lane>>2, t = lane&3) holds
//   a = A[g][t],  b = B[t][g],  d0 = D[g][2t], d1 = D[g][2t+1].
// Measured on B200 (scripts/micro/dmma_bench.cu): 37.1 TFLOP/s with one warp per SM sub-partition, the
// same peak as the DFMA pipe at 1/8 of the instructions, and the k-reduction happens inside the pipe.
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---- pending rank-8 update  A -= V W^T + W V^T,  in 16 x 32 units (rows R0.., columns C0..) of up to 8 tiles.
// Per 8x8 tile the stored values are the C fragment (lane: row g, columns 2t, 2t+1 = one 128-bit access), the
// row operands -V[t][row g], -W[t][row g] are A fragments and W[t][col g], V[t][col g] are B fragments: two
// DMMAs per tile.  Loads and compute are separate so that the sweep can fetch the next unit (possibly from the
// global workspace, i.e. L2) while the current one is in the pipe.
struct SbrUnit {
    int u;       // row-major unit number: 16-row blocks 2a and 2a+1 have a+1 units each
    int i16, j;  // 16-row block, 32-column block (j <= i16 >> 1)
    __device__ __forceinline__ bool valid() const { return u >= 0; }
    __device__ __forceinline__ static int first_of(int i16) {  // number of the first unit of a 16-row block
        const int a = i16 >> 1;
        return (i16 & 1) ? (a + 1) * (a + 1) : a * (a + 1);
    }
    __device__ __forceinline__ void locate() {  // (i16, j) of unit u; i16 only ever moves down
        if (u < 0) return;
        while (first_of(i16) > u) --i16;
        j = u - first_of(i16);
    }
};

__device__ __forceinline__ bool sbr_tile_ok(bool diag, int trr, int tc, int g, int t, bool rv) {
    return rv && (!(diag && tc == trr) || 2 * t <= g);  // diagonal tile: the pair (2t, 2t+1) starts at or left of the diagonal
}

__device__ __forceinline__ void sbr_unit_load(const double* __restrict__ base, SbrUnit u, int p0, int g, int t,
                                              double2 (&c)[8]) {
    const int R0 = 16 * u.i16, C0 = 32 * u.j;
    const bool diag = u.j == (u.i16 >> 1);
#pragma unroll
    for (int tr = 0; tr < 2; ++tr) {
        const int row = R0 + 8 * tr + g;
        const bool rv = row < p0;
        const double* rowp = base + poff(row) + C0 + 2 * t;
        const int trr = 2 * (u.i16 & 1) + tr;  // 8-row block index inside the 32 x 32 block
#pragma unroll
        for (int tc = 0; tc < 4; ++tc) {
            const bool ok = sbr_tile_ok(diag, trr, tc, g, t, rv) && !(diag && tc > trr);
            c[tr * 4 + tc] = ok ? *reinterpret_cast<const double2*>(rowp + 8 * tc) : make_double2(0.0, 0.0);
        }
    }
}

__device__ __forceinline__ void sbr_unit_update(double* __restrict__ base, SbrUnit u, int p0, const double* __restrict__ Vb,
                                                const double* __restrict__ Wb, int st, int g, int t, double2 (&c)[8]) {
    const int R0 = 16 * u.i16, C0 = 32 * u.j;
    const bool diag = u.j == (u.i16 >> 1);
    double bw[4], bv[4];
#pragma unroll
    for (int tc = 0; tc < 4; ++tc) {
        bw[tc] = Wb[t * st + C0 + 8 * tc + g];
        bv[tc] = Vb[t * st + C0 + 8 * tc + g];
    }
#pragma unroll
    for (int tr = 0; tr < 2; ++tr) {
        const int row = R0 + 8 * tr + g;
        if (R0 + 8 * tr >= p0) continue;  // warp-uniform
        const bool rv = row < p0;
        const double av = -Vb[t * st + row], aw = -Wb[t * st + row];
        double* rowp = base + poff(row) + C0 + 2 * t;
        const int trr = 2 * (u.i16 & 1) + tr;
#pragma unroll
        for (int tc = 0; tc < 4; ++tc) {
            if (diag && tc > trr) continue;  // above the diagonal: not stored (warp-uniform)
            dmma(c[tr * 4 + tc].x, c[tr * 4 + tc].y, av, bw[tc]);
            dmma(c[tr * 4 + tc].x, c[tr * 4 + tc].y, aw, bv[tc]);
            if (sbr_tile_ok(diag, trr, tc, g, t, rv)) *reinterpret_cast<double2*>(rowp + 8 * tc) = c[tr * 4 + tc];
        }
    }
}

// update-only sweep over the units of the leading p0 x p0 triangle, bottom-up (the rows that live in the
// global workspace first); worker w takes the units w, w + nworkers, ... counted from the last one
__device__ __noinline__ void sbr_update_sweep(double* __restrict__ A, double* __restrict__ G, int rows_smem, int p0,
                                                 int worker, int nworkers, const double* Vb, const double* Wb, int st,
                                                 int g, int t) {
    SbrUnit cur;
    cur.i16 = ((p0 + 15) >> 4) - 1;
    cur.u = SbrUnit::first_of(cur.i16) + (cur.i16 >> 1) - worker;  // last unit, minus this worker's offset
    cur.locate();
    double2 c[8], cn[8];
    if (cur.valid()) sbr_unit_load(16 * cur.i16 < rows_smem ? A : G, cur, p0, g, t, c);
    while (cur.valid()) {
        SbrUnit nxt = cur;
        nxt.u -= nworkers;
        nxt.locate();
        if (nxt.valid()) sbr_unit_load(16 * nxt.i16 < rows_smem ? A : G, nxt, p0, g, t, cn);
        sbr_unit_update(16 * cur.i16 < rows_smem ? A : G, cur, p0, Vb, Wb, st, g, t, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = cn[i];
        cur = nxt;
    }
}

// ---- Y = A U contributions of one stored 32 x 32 block (rows RB0.., columns CB0..), 8-row strips [tr0, tre).
// All tiles are fetched first (one 128-bit access per tile and lane, C-fragment layout), so a block that lives in
// the global workspace pays the L2 latency once.  Per tile:
//   rows:    D[row g][kk] += sum_cols tile[g][col] U[kk][col]   tile as A fragment: its two registers are the
//                                                              k-slices {2t} and {2t+1}
//   columns: D[col g][kk] += sum_rows tile[row][g] U[kk][row]   the transposed tile as A fragment, built with four
//                                                              shuffles per k-slice (rows 0-3 and 4-7)
// (kk = 0..3: the B fragments carry U in the columns n < 4, zero in the rest).  On diagonal tiles elements with
// column > row do not exist and the diagonal enters the row sums only.
__device__ __forceinline__ void sbr_symm_block(const double* __restrict__ base, bool diag, int RB0, int CB0, int tr0,
                                               int tre, int p0, const double* __restrict__ Ub, int st, int lane, int g,
                                               int t, double (&rowsum)[4][2], double (&colsum)[4][2]) {
    double2 c[4][4];
#pragma unroll
    for (int tr = 0; tr < 4; ++tr) {
        const int row = RB0 + 8 * tr + g;
        const bool rv = tr >= tr0 && tr < tre && row < p0;
        const double* rowp = base + poff(row) + CB0 + 2 * t;
#pragma unroll
        for (int tc = 0; tc < 4; ++tc) {
            const bool dt = diag && tc == tr;
            const bool ok = rv && !(diag && tc > tr) && (!dt || 2 * t <= g);
            c[tr][tc] = ok ? *reinterpret_cast<const double2*>(rowp + 8 * tc) : make_double2(0.0, 0.0);
            if (dt && !(2 * t + 1 <= g)) c[tr][tc].y = 0.0;
        }
    }
    double bu0[4], bu1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double2 uu = (g < 4) ? *reinterpret_cast<const double2*>(Ub + g * st + CB0 + 8 * i + 2 * t)
                                   : make_double2(0.0, 0.0);
        bu0[i] = uu.x;
        bu1[i] = uu.y;
    }
    const int src0 = 4 * t + (g >> 1), src1 = src0 + 16;  // lanes that hold tile[t][g], tile[4 + t][g]
    const bool odd = g & 1;
#pragma unroll
    for (int tr = 0; tr < 4; ++tr) {
        if (tr < tr0 || tr >= tre) continue;  // warp-uniform
        const int R0 = RB0 + 8 * tr;
        const double bur0 = (g < 4) ? Ub[g * st + R0 + t] : 0.0, bur1 = (g < 4) ? Ub[g * st + R0 + 4 + t] : 0.0;
        double da0 = 0.0, da1 = 0.0, db0 = 0.0, db1 = 0.0;  // two chains for the row sums
#pragma unroll
        for (int tc = 0; tc < 4; ++tc) {
            if (diag && tc > tr) continue;  // warp-uniform
            if (tc & 1) {
                dmma(db0, db1, c[tr][tc].x, bu0[tc]);
                dmma(db0, db1, c[tr][tc].y, bu1[tc]);
            } else {
                dmma(da0, da1, c[tr][tc].x, bu0[tc]);
                dmma(da0, da1, c[tr][tc].y, bu1[tc]);
            }
            const double x0 = __shfl_sync(0xffffffffu, c[tr][tc].x, src0), y0 = __shfl_sync(0xffffffffu, c[tr][tc].y, src0);
            const double x1 = __shfl_sync(0xffffffffu, c[tr][tc].x, src1), y1 = __shfl_sync(0xffffffffu, c[tr][tc].y, src1);
            double e0 = odd ? y0 : x0, e1 = odd ? y1 : x1;
            if (diag && tc == tr) {  // strictly below the diagonal only
                if (!(g < t)) e0 = 0.0;
                if (!(g < 4 + t)) e1 = 0.0;
            }
            dmma(colsum[tc][0], colsum[tc][1], e0, bur0);
            dmma(colsum[tc][0], colsum[tc][1], e1, bur1);
        }
        rowsum[tr][0] += da0 + db0;
        rowsum[tr][1] += da1 + db1;
    }
}

// ---- LQ of the 4 x p0 panel P by one warp, in place in shared memory: four Householder reflectors, row 3 first
// (pivot column p0-1), then row 2 (pivot p0-2), ...  Lane l works on the column pairs {2l, 2l+1} + 64q.  One
// butterfly of four values per reflector: |x|^2, the products of x with the rows still to be reduced, and the
// products of x with the earlier reflectors (they give U^T U, hence T, without a reduction of their own).  Writes
// U (zero beyond each pivot), T and the R block of the band.  Nothing but scalars lives in registers.
template <int NP>
__device__ __forceinline__ void sbr_panel_lq(double* __restrict__ P, double* __restrict__ Ub, double* __restrict__ Tm,
                                             double* __restrict__ G, int p0, int st, int lane) {
    double tk[4] = {0.0, 0.0, 0.0, 0.0};
    double zz[4][4];  // zz[k][j] = u_j . u_k, j < k
#ifdef VSP_PHASE_TIMING
    long long lq_t[20];
    int lq_n = 0;
    lq_t[lq_n++] = clock64();
#define VSP_LQ_MARK() lq_t[lq_n++] = clock64()
#else
#define VSP_LQ_MARK() ((void)0)
#endif
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t = 3 - k;
        const int pc = p0 - 4 + t;  // pivot column of row t
        if (pc >= 0) {              // warp-uniform
            // v[0] = |x|^2, v[1 + t2] = P[t2] . x (t2 < t), v[1 + t + j] = u_j . x (j < k): always four values
            double v[4] = {0.0, 0.0, 0.0, 0.0};
            // branch-free over the column chunks (all loads are issued together); pairs at or beyond the pivot
            // are masked, chunks beyond the buffer re-read its last pair
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int c0 = 2 * lane + 64 * q;
                const int cl = c0 < st ? c0 : st - 2;
                double2 xm = *reinterpret_cast<const double2*>(P + t * st + cl);
                xm.x = (c0 < pc) ? xm.x : 0.0;
                xm.y = (c0 + 1 < pc) ? xm.y : 0.0;
                v[0] = fma(xm.x, xm.x, fma(xm.y, xm.y, v[0]));
#pragma unroll
                for (int t2 = 0; t2 < 3; ++t2)
                    if (t2 < t) {
                        const double2 r2 = *reinterpret_cast<const double2*>(P + t2 * st + cl);
                        v[1 + t2] = fma(r2.x, xm.x, fma(r2.y, xm.y, v[1 + t2]));
                    }
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < k) {
                        const double2 uj = *reinterpret_cast<const double2*>(Ub + j * st + cl);
                        v[1 + t + j] = fma(uj.x, xm.x, fma(uj.y, xm.y, v[1 + t + j]));
                    }
            }
            VSP_LQ_MARK();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] += shfl_xor_d(v[i], o);
            VSP_LQ_MARK();
            const double xnorm2 = v[0], alpha = P[t * st + pc];
            double beta = alpha, tau = 0.0, vscale = 0.0;
            // |x| <= n after the power-of-four scaling; a row below 1e-140 is dropped (set to zero)
            if (xnorm2 > 1e-280) {
                const double s2 = fma(alpha, alpha, xnorm2);
                const double rs = fast_rsqrt(s2);  // 1/||x||
                const double nrm = s2 * rs;        // ||x||
                beta = -copysign(nrm, alpha);
                tau = fma(fabs(alpha), rs, 1.0);  // (beta - alpha)/beta = 1 + |alpha|/||x||
                vscale = copysign(fast_rcp(fabs(alpha) + nrm), alpha);  // 1/(alpha - beta)
            }
            tk[k] = tau;
            double coef[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int t2 = 0; t2 < 3; ++t2)
                if (t2 < t) coef[t2] = tau * fma(vscale, v[1 + t2], P[t2 * st + pc]);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (j < k) zz[k][j] = (tau != 0.0) ? fma(vscale, v[1 + t + j], Ub[j * st + pc]) : 0.0;
            VSP_LQ_MARK();
            __syncwarp();  // the pivot column has been read by everybody
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int c0 = 2 * lane + 64 * q;
                const int cl = c0 < st ? c0 : st - 2;
                const bool live = c0 < st;
                double2 x = *reinterpret_cast<const double2*>(P + t * st + cl);
                double2 u;
                u.x = (tau == 0.0 || c0 > pc) ? 0.0 : ((c0 < pc) ? x.x * vscale : 1.0);
                u.y = (tau == 0.0 || c0 + 1 > pc) ? 0.0 : ((c0 + 1 < pc) ? x.y * vscale : 1.0);
#pragma unroll
                for (int t2 = 0; t2 < 3; ++t2)
                    if (t2 < t) {
                        double2 r2 = *reinterpret_cast<const double2*>(P + t2 * st + cl);
                        r2.x = fma(-coef[t2], u.x, r2.x);
                        r2.y = fma(-coef[t2], u.y, r2.y);
                        if (live) *reinterpret_cast<double2*>(P + t2 * st + c0) = r2;
                    }
                x.x = (c0 < pc) ? 0.0 : ((c0 == pc) ? beta : x.x);
                x.y = (c0 + 1 < pc) ? 0.0 : ((c0 + 1 == pc) ? beta : x.y);
                if (live) {
                    *reinterpret_cast<double2*>(P + t * st + c0) = x;
                    *reinterpret_cast<double2*>(Ub + k * st + c0) = u;
                }
            }
            __syncwarp();
            VSP_LQ_MARK();
        } else {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int c0 = 2 * lane + 64 * q;
                if (c0 < st) *reinterpret_cast<double2*>(Ub + k * st + c0) = make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (j < k) zz[k][j] = 0.0;
        }
    }
#ifdef VSP_PHASE_TIMING
    if (blockIdx.x == 200 && lane == 0 && (p0 == 92 || p0 == 20)) {
        printf("[lq p0=%d]", p0);
        for (int i = 1; i < lq_n; ++i) printf(" %lld", lq_t[i] - lq_t[i - 1]);
        printf("\n");
    }
#endif
#undef VSP_LQ_MARK
    // R block -> band output (rows p0+t, columns max(p0-4+t, 0) .. p0-1): 16 entries
    if (lane < 16) {
        const int t = lane >> 2, c = p0 - 4 + (lane & 3);
        if (c >= 0 && c >= p0 - 4 + t) G[poff(p0 + t) + c] = P[t * st + c];
    }
    // T from U^T U (forward recurrence in application order)
    if (lane == 0) {
        const double t00 = tk[0], t11 = tk[1], t22 = tk[2], t33 = tk[3];
        const double t01 = -t11 * (t00 * zz[1][0]);
        const double t02 = -t22 * fma(t01, zz[2][1], t00 * zz[2][0]);
        const double t12 = -t22 * (t11 * zz[2][1]);
        const double t03 = -t33 * fma(t02, zz[3][2], fma(t01, zz[3][1], t00 * zz[3][0]));
        const double t13 = -t33 * fma(t12, zz[3][2], t11 * zz[3][1]);
        const double t23 = -t33 * (t22 * zz[3][2]);
        Tm[0] = t00; Tm[1] = t01; Tm[2] = t02; Tm[3] = t03;
        Tm[4] = 0.0; Tm[5] = t11; Tm[6] = t12; Tm[7] = t13;
        Tm[8] = 0.0; Tm[9] = 0.0; Tm[10] = t22; Tm[11] = t23;
        Tm[12] = 0.0; Tm[13] = 0.0; Tm[14] = 0.0; Tm[15] = t33;
    }
}

// NQ = 32-column chunks a lane of the LQ warp holds (= max warps); MINB = CTAs per SM
// BPW = index blocks per warp in the symmetric products (1: the matrix is small enough for one warp per block;
// 3: orders up to 768 with 8 warps, the matrix in the global workspace throughout)
template <int NQ, int MINB, int BPW>
__global__ void __launch_bounds__(32 * NQ, MINB)
    sbr_band_kernel(const ItemDesc* __restrict__ items, int item_base, double* __restrict__ ws, int st, int rows_smem,
                    int m_start, int m_stop) {
    extern __shared__ __align__(16) double smem[];
    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    // warp index broadcast from lane 0: the compiler then knows it is warp-uniform, so the `warp == 0` /
    // `warp < nb` regions are uniform branches and the shuffles inside them are plain SHFLs
    const int lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), NW = nthreads >> 5;
    const int g = lane >> 2, t = lane & 3;  // DMMA fragment coordinates

    double* Vb = smem;            // [4][st] previous reflectors (pending update)
    double* Wb = Vb + 4 * st;     // [4][st] previous W
    double* Ub = Wb + 4 * st;     // [4][st] current reflectors
    double* Yoth = Ub + 4 * st;   // [4][st] Y = A U: partner sums during the pass, own sums added at its end; also the panel buffer P
    double* Tm = Yoth + 4 * st;   // [16] T, row-major, upper triangular
    double* tauv = Tm + 16;       // [4] (+12 pad)
    double* Zpart = tauv + 16;    // [NW][16]
    double* A = Zpart + 16 * NW;  // rows < rows_smem (+ kSbrPad)

    // The reduction of one matrix may be split over several launches (m_start > 0: resume with the leading
    // m_start x m_start block, already scaled and up to date in the workspace; m_stop > 0: hand over as soon as the
    // active order is <= m_stop): the small trailing steps are latency-bound, and a smaller footprint lets more
    // matrices share an SM.
    double* __restrict__ G = ws + it.gram_off;
    double* out = ws + it.de_off;
    const int n0 = m_start > 0 ? m_start : n;  // order this launch starts from
    const int rs_eff = rows_smem < n0 ? rows_smem : n0;
    const int total = poff(n0), in_smem = poff(rs_eff);
    if (m_start > 0) {
        if (out[2 * n + MISC_FLAGS] != 0.0) return;  // non-finite / all-zero: flagged by the first launch
        for (int i = tid; i < in_smem; i += nthreads) A[i] = G[i];
    } else {
        // ---- load + condition the Gram matrix (power-of-four scale so that |G_ij| <= 1 and the singular values
        //      un-scale exactly; NaN/Inf anywhere in W shows on the Gram diagonal)
        double md = 0.0;
        int bad = 0;
        for (int c = lane; c < n; c += 32) {
            const double gd = G[poff(c) + c];
            if (!isfinite(gd)) bad = 1;
            md = fmax(md, gd);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            md = fmax(md, shfl_xor_d(md, o));
            bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        }
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
        for (int i = tid; i < in_smem; i += nthreads) A[i] = G[i] * scale;
        for (int i = in_smem + tid; i < total; i += nthreads) G[i] *= scale;  // rows >= rows_smem: in place
    }
    for (int i = tid; i < kSbrPad; i += nthreads) A[in_smem + i] = 0.0;
    for (int i = tid; i < 16 * st; i += nthreads) Vb[i] = 0.0;
    __syncthreads();

#ifdef VSP_PHASE_TIMING  // per-phase cycle counters of this warp (development builds only)
    long long t_phase[6] = {0, 0, 0, 0, 0, 0};  // mini | LQ (+wait) | pass compute | pass barriers | W-phase | panels
    long long t_prev[5] = {0, 0, 0, 0, 0};
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
    int m = n0;
    while (m >= kSbrB + 2 && m > m_stop) {
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
            if (pending) sbr_update_sweep(A, G, rs_eff, p0, warp - 1, NW - 1, Vb, Wb, st, g, t);
        } else {
            sbr_panel_lq<(NQ * BPW + 1) / 2>(P, Ub, Tm, G, p0, st, lane);
            if (NW == 1 && pending) sbr_update_sweep(A, G, rs_eff, p0, 0, 1, Vb, Wb, st, g, t);
        }
        __syncthreads();
        VSP_LAP(1);

        // ---- (3) symmetric products Y = A U over the leading p0 x p0 triangle (cyclic block schedule)
        const int nb = (p0 + 31) >> 5;
        {
            double own[BPW][4][2];  // sums of the own index blocks (warp, warp + NW, ...), as D fragments:
                                    // [block][8-row/col group][kk = 2t, 2t+1]
#pragma unroll
            for (int b = 0; b < BPW; ++b)
#pragma unroll
                for (int i = 0; i < 4; ++i) own[b][i][0] = own[b][i][1] = 0.0;
            const int nsteps = (nb >> 1) + 1;
            // The partner sums of step s go to buffer s mod 3: besides Y itself, the buffers of the previous V and W
            // are free from here to the W-phase (the update sweep and the mini-pass were their last readers).  Every
            // step touches every block of its buffer exactly once, so steps 0..2 store, need no barrier between them
            // and the warps stream through their blocks; a barrier before steps 3, 6, ... re-opens the buffers.
            double* const Ybuf[3] = {Yoth, Wb, Vb};
            for (int s = 0; s < nsteps; ++s) {
                if (s > 0 && s % 3 == 0) __syncthreads();
#pragma unroll
                for (int b = 0; b < BPW; ++b) {
                    const int blk_own = warp + b * NW;
                    if (blk_own >= nb) continue;  // warp-uniform
                    int o = blk_own + s;
                    if (o >= nb) o -= nb;
                    const bool half = (2 * s == nb);   // the pair {w, w + nb/2} is met from both sides
                    const bool own_cols = (s == 0) || (o > blk_own);
                    const int rb = own_cols ? o : blk_own, cb = own_cols ? blk_own : o;  // stored tile (rb, cb)
                    const int tr0 = (half && !own_cols) ? 2 : 0, tr1 = (half && own_cols) ? 2 : 4;
                    const int RB0 = 32 * rb, CB0 = 32 * cb;
                    // the sums of the own index block accumulate in `own` across the steps, those of the partner
                    // block in `oth` (stored / added to this step's buffer below): own = columns <=> own_cols
                    double oth[4][2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) oth[i][0] = oth[i][1] = 0.0;
                    {
                        int tre = tr1;  // strips [tr0, tre) exist
                        while (tre > tr0 && RB0 + 8 * (tre - 1) >= p0) --tre;
                        const double* blk = (RB0 + 8 * tr0 < rs_eff) ? A : G;  // rs_eff is a multiple of 16 or >= p0:
                        if (tre - tr0 > 2 && RB0 + 16 >= rs_eff && RB0 < rs_eff) {
                            // the block straddles the two address spaces: two half blocks
                            if (own_cols) {
                                sbr_symm_block(A, s == 0, RB0, CB0, tr0, 2, p0, Ub, st, lane, g, t, oth, own[b]);
                                sbr_symm_block(G, s == 0, RB0, CB0, 2, tre, p0, Ub, st, lane, g, t, oth, own[b]);
                            } else {
                                sbr_symm_block(A, s == 0, RB0, CB0, tr0, 2, p0, Ub, st, lane, g, t, own[b], oth);
                                sbr_symm_block(G, s == 0, RB0, CB0, 2, tre, p0, Ub, st, lane, g, t, own[b], oth);
                            }
                        } else if (own_cols) {
                            sbr_symm_block(blk, s == 0, RB0, CB0, tr0, tre, p0, Ub, st, lane, g, t, oth, own[b]);
                        } else {
                            sbr_symm_block(blk, s == 0, RB0, CB0, tr0, tre, p0, Ub, st, lane, g, t, own[b], oth);
                        }
                    }
                    // partner block: into this step's buffer (no other warp touches the block during the step)
                    if (t < 2) {
                        double* const yb = Ybuf[s % 3];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            double* y0 = yb + (2 * t) * st + 32 * o + 8 * i + g;
                            if (s < 3) {  // first touch of this block in this buffer: plain store
                                y0[0] = oth[i][0];
                                y0[st] = oth[i][1];
                            } else {
                                y0[0] += oth[i][0];
                                y0[st] += oth[i][1];
                            }
                        }
                    }
                }
                VSP_LAP(2);
            }
            __syncthreads();
            VSP_LAP(3);
            if (t < 2) {
#pragma unroll
                for (int b = 0; b < BPW; ++b) {
                    const int blk_own = warp + b * NW;
                    if (blk_own >= nb) continue;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double* y0 = Yoth + (2 * t) * st + 32 * blk_own + 8 * i + g;  // all steps are done: exclusive again
                        y0[0] += own[b][i][0];
                        y0[st] += own[b][i][1];
                    }
                }
            }
        }
        __syncthreads();

        // ---- (4) W-phase: X = Y T, Z = U^T X (block reduction), S = T^T Z, W = X - U S / 2
        {
            double T[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) T[i] = Tm[i];
            // thread tid handles the rows tid, tid + nthreads, ... (BPW of them)
            double x[BPW][4], u[BPW][4], zp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zp[i] = 0.0;
#pragma unroll
            for (int q = 0; q < BPW; ++q) {
                const int r = tid + q * nthreads;
#pragma unroll
                for (int k = 0; k < 4; ++k) x[q][k] = u[q][k] = 0.0;
                if (r < p0) {
                    double y[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        y[k] = Yoth[k * st + r];
                        if (nb >= 2) y[k] += Wb[k * st + r];  // partner sums of the steps 1, 4, ...
                        if (nb >= 4) y[k] += Vb[k * st + r];  // partner sums of the steps 2, 5, ...
                        u[q][k] = Ub[k * st + r];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int k = 0; k <= j; ++k) x[q][j] = fma(y[k], T[k * 4 + j], x[q][j]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) zp[i * 4 + j] = fma(u[q][i], x[q][j], zp[i * 4 + j]);
            }
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
            double sik[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sik[i] = __shfl_sync(0xffffffffu, se, i);
#pragma unroll
            for (int q = 0; q < BPW; ++q) {
                const int r = tid + q * nthreads;
                if (r < st) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        double wv = x[q][k];
#pragma unroll
                        for (int i = 0; i < 4; ++i) wv = fma(-0.5 * u[q][i], sik[4 * i + k], wv);
                        Wb[k * st + r] = (r < p0) ? wv : 0.0;
                    }
                }
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
        if (blockIdx.x == 200 && tid == 0)
            printf("[sbr panel] m=%d  cycles: mini %lld LQ %lld pass %lld passbar %lld W %lld\n", m + 4, t_phase[0] - t_prev[0],
                   t_phase[1] - t_prev[1], t_phase[2] - t_prev[2], t_phase[3] - t_prev[3], t_phase[4] - t_prev[4]);
        for (int q_ = 0; q_ < 5; ++q_) t_prev[q_] = t_phase[q_];
        t_phase[5] += 1;
#endif
    }
#ifdef VSP_PHASE_TIMING
    if (blockIdx.x == 200 && lane == 0)
        printf("[sbr n=%d warp %d] cycles: mini %lld  LQ %lld  pass %lld  pass-barrier %lld  W %lld  panels %lld\n", n, warp,
               t_phase[0], t_phase[1], t_phase[2], t_phase[3], t_phase[4], t_phase[5]);
#endif
#undef VSP_LAP
    if (m >= kSbrB + 2) {
        // ---- hand-over: bring the leading m x m block up to date and return it to the workspace
        if (pending) sbr_update_sweep(A, G, rs_eff, m, warp, NW, Vb, Wb, st, g, t);
        __syncthreads();
        const int back = poff(rs_eff < m ? rs_eff : m);
        for (int i = tid; i < back; i += nthreads) G[i] = A[i];
        return;
    }
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
