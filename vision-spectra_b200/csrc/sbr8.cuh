// sbr8.cuh -- stage 2a-1 for orders up to kSbr8MaxN: blocked Householder reduction of the (scaled) Gram matrix to a
// symmetric band of bandwidth 8 on the FP64 tensor cores, the whole matrix resident in shared memory.  chase8.cuh
// finishes the job (band -> tridiagonal).  Round-2 successor of sbr_band.cuh (bandwidth 4), which stays for larger
// orders.  What changed and why (ncu of the round-1 kernel: 39 % of its shared-memory wavefronts were bank conflicts,
// tensor pipe 24 % active, the rows that did not fit two CTAs per SM cost ~20 k cycles per panel through L2):
//
//   * panel width 8 = DMMA's n: the symmetric products fill the MMA (b = 4 wasted half of it), half as many panels,
//     half as many passes over the matrix, half as many barriers;
//   * the triangle is stored as 8 x 8 tiles, tile (I, J), J <= I, at ((I (I + 1) / 2 + J) * 64 doubles, row-major inside
//     the tile = the DMMA C-fragment order: one conflict-free 128-bit access per lane and tile (the row-packed layout
//     gave two-way conflicts on 7 of 8 tile rows).  Diagonal tiles are stored full (both triangles), so they enter the
//     products through their row sums only;
//   * operand arrays [8][st] with st = 4 mod 8 doubles: the four k-rows of a fragment fall into different bank groups;
//   * the panel LQ runs out of registers: lane l, slot q holds column p0 - 1 - (l + 32 q) of the eight panel rows, so the
//     pivot of reflector k is always (lane k, slot 0) -- compile-time register indices, no local memory -- and the
//     eight inner products of a reflector are reduced together by one folding butterfly (17 shuffles, not 40);
//   * the W-phase (X = Y T, Z = U^T X, S = T^T Z, W = X - U S / 2) is six DMMAs per 8-row strip instead of scalar
//     code and a 16-value shuffle fold per thread;
//   * no global-memory path: an order range runs with as many CTAs per SM as fit entirely in shared memory
//     (vspectra_api.cu picks the ranges), so no step ever waits for L2.
//
// Internal order N = round_up(n, 8); the matrix sits at the bottom-right of the N x N frame (off = N - n zero rows /
// columns at the top-left: they stay zero through the reduction and decouple).  Elimination is bottom-up: the active
// matrix is the leading m x m block, the panel is tile row Ip = p0 / 8, p0 = m - 8.  One panel step:
//   (1) mini-pass (all warps)   panel tiles get the pending update A -= V W^T + W V^T (four DMMAs per tile); the
//                               diagonal tile goes to the band output, the rest to the panel buffer, REVERSED: buffer
//                               row k = panel row 7 - k, so reflector k (application order) is built from buffer
//                               row k and overwrites it: the panel buffer becomes U in place.
//   (2) LQ (last warp)          eight reflectors, pivot column of reflector k = p0 - 1 - k; T from U^T U.
//       update (other warps)    the pending update on the leading p0 x p0 triangle, four DMMAs per tile.
//   (3) products (all warps)    Y = A U, cyclic block-pair schedule over index blocks of 8 TB (one warp per block): per
//                               off-diagonal tile two DMMAs for the row sums and two for the column sums (transposed
//                               fragment by four shuffles), partner sums through the two free operand buffers.
//   (4) W-phase (all warps)     per 8-row strip, DMMA; V <- U.
// Band output: Bd[r * 9 + j] = A[r][r - j], j = 0..8, frame indices.
#pragma once

#include "common.cuh"

namespace vsp {

#ifndef VSP_SBR8_QUIET_LQ
#define VSP_SBR8_QUIET_LQ 1
#endif
constexpr bool kQuietLq = VSP_SBR8_QUIET_LQ != 0;
constexpr int kSbr8MaxN = 200;  // whole triangle + operand arrays within 227 KB (N = 200: 166.4 + 39.2 KB)
constexpr int kBandW = 9;       // band row: diagonal + 8 sub-diagonals

VSP_HD int sbr8_order(int n) { return (n + 7) & ~7; }
VSP_HD int sbr8_stride(int N) { return N + 4; }  // N is a multiple of 8: = 4 mod 8
VSP_HD int64_t sbr8_band_doubles(int n) { return (int64_t)kBandW * sbr8_order(n); }
VSP_HD int sbr8_band_off(int n) { return tile_off(sbr8_order(n) >> 3, 0); }  // doubles: behind the tiled triangle
// shared memory (doubles): V W U [8][st] | T 64 | Zpart [NW][64] | S scratch [NW][64] | tiles
__host__ __device__ inline size_t sbr8_fixed_doubles(int st, int nw) { return (size_t)24 * st + 64 + (size_t)128 * nw; }
__host__ __device__ inline size_t sbr8_smem_bytes_tiles(int tiles, int st, int nw) {
    return sizeof(double) * (sbr8_fixed_doubles(st, nw) + (size_t)tiles * 64);
}
__host__ __device__ inline size_t sbr8_smem_bytes(int N_active, int st, int nw) {
    const int nt = N_active >> 3;
    return sbr8_smem_bytes_tiles((nt * (nt + 1)) >> 1, st, nw);
}

#if defined(__CUDACC__)

__device__ __forceinline__ void dmma8(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Where the tiles live: the first ts64 / 64 tiles (the top tile rows: they survive longest) in shared memory, the rest
// in place in the item's global workspace region (L2).  All-shared launches pass ts64 >= the triangle's size and never
// take the global branch; the first order range of n > 136 runs two CTAs per SM with its bottom tile rows in L2 --
// they are eliminated first, so the global share shrinks to nothing by the time the range hands over.  A tile is 512
// contiguous bytes: one fully coalesced 128-bit access per lane either way.  `off` is warp-uniform.
template <bool L2>  // false: every tile in shared memory (no branch, no extra registers in the capped instantiations)
struct TileMem {
    double* A;
    double* G;
    int ts64;
    __device__ __forceinline__ double2 load(int off, int lane) const {
        if (!L2) return *(reinterpret_cast<const double2*>(A + off) + lane);
        return off < ts64 ? *(reinterpret_cast<const double2*>(A + off) + lane) : __ldcg(reinterpret_cast<const double2*>(G + off) + lane);
    }
    __device__ __forceinline__ void store(int off, int lane, double2 v) const {
        if (!L2 || off < ts64)
            *(reinterpret_cast<double2*>(A + off) + lane) = v;
        else
            *(reinterpret_cast<double2*>(G + off) + lane) = v;
    }
};

// A -= V W^T + W V^T on one tile (rows 8I.., columns 8J..): operands as fragments
//   a*: -V / -W [k = 4 s + t][row g]      b*: W / V [k = 4 s + t][col g]
struct Frag8 {
    double v0, v1, w0, w1;
};
__device__ __forceinline__ Frag8 sbr8_frag(const double* __restrict__ Vb, const double* __restrict__ Wb, int st, int i0, int g,
                                           int t) {
    Frag8 f;
    f.v0 = Vb[t * st + i0 + g];
    f.v1 = Vb[(4 + t) * st + i0 + g];
    f.w0 = Wb[t * st + i0 + g];
    f.w1 = Wb[(4 + t) * st + i0 + g];
    return f;
}
__device__ __forceinline__ void sbr8_tile_update(double2& c, const Frag8& ra, const Frag8& cb) {
    dmma8(c.x, c.y, -ra.v0, cb.w0);
    dmma8(c.x, c.y, -ra.v1, cb.w1);
    dmma8(c.x, c.y, -ra.w0, cb.v0);
    dmma8(c.x, c.y, -ra.w1, cb.v1);
}

// pending update of the leading nt x nt tile triangle; worker w of nwk takes tiles w, w + nwk, ... (linear order), TWO
// at a time: the four dependent DMMAs of a tile are a 104-cycle chain, a second tile's chain fills the pipe meanwhile;
// the next pair is fetched while the current one is in the pipe (it may come from L2)
struct TilePos {
    int I, J;
    __device__ __forceinline__ void advance(int by) {
        J += by;
        while (J > I) {
            J -= I + 1;
            ++I;
        }
    }
};
template <bool PAIR>
__device__ __forceinline__ void sbr8_update_sweep(const TileMem<PAIR>& tm, int nt, int worker, int nwk, const double* Vb,
                                                  const double* Wb, int st, int lane, int g, int t) {
    const int total = (nt * (nt + 1)) >> 1;
    if (worker >= total) return;
    if (!PAIR) {  // one tile at a time, all in shared memory (the register-capped instantiations)
        TilePos p{0, 0};
        p.advance(worker);
        for (int tau = worker; tau < total; tau += nwk) {
            double2 c = tm.load(tile_off(p.I, p.J), lane);
            const Frag8 ra = sbr8_frag(Vb, Wb, st, 8 * p.I, g, t);
            const Frag8 cb = sbr8_frag(Vb, Wb, st, 8 * p.J, g, t);
            sbr8_tile_update(c, ra, cb);
            tm.store(tile_off(p.I, p.J), lane, c);
            p.advance(nwk);
        }
        return;
    }
    TilePos p0{0, 0}, p1;
    p0.advance(worker);
    p1 = p0;
    p1.advance(nwk);
    bool two = worker + nwk < total;
    double2 c0 = tm.load(tile_off(p0.I, p0.J), lane);
    double2 c1 = two ? tm.load(tile_off(p1.I, p1.J), lane) : make_double2(0.0, 0.0);
    for (int tau = worker; tau < total; tau += 2 * nwk) {
        TilePos q0 = p1, q1;
        q0.advance(nwk);
        q1 = q0;
        q1.advance(nwk);
        const bool n0 = tau + 2 * nwk < total, n1 = tau + 3 * nwk < total;
        double2 d0 = make_double2(0.0, 0.0), d1 = make_double2(0.0, 0.0);
        if (n0) d0 = tm.load(tile_off(q0.I, q0.J), lane);
        if (n1) d1 = tm.load(tile_off(q1.I, q1.J), lane);
        const Frag8 ra0 = sbr8_frag(Vb, Wb, st, 8 * p0.I, g, t), cb0 = sbr8_frag(Vb, Wb, st, 8 * p0.J, g, t);
        if (two) {
            const Frag8 ra1 = sbr8_frag(Vb, Wb, st, 8 * p1.I, g, t), cb1 = sbr8_frag(Vb, Wb, st, 8 * p1.J, g, t);
            dmma8(c0.x, c0.y, -ra0.v0, cb0.w0);
            dmma8(c1.x, c1.y, -ra1.v0, cb1.w0);
            dmma8(c0.x, c0.y, -ra0.v1, cb0.w1);
            dmma8(c1.x, c1.y, -ra1.v1, cb1.w1);
            dmma8(c0.x, c0.y, -ra0.w0, cb0.v0);
            dmma8(c1.x, c1.y, -ra1.w0, cb1.v0);
            dmma8(c0.x, c0.y, -ra0.w1, cb0.v1);
            dmma8(c1.x, c1.y, -ra1.w1, cb1.v1);
            tm.store(tile_off(p0.I, p0.J), lane, c0);
            tm.store(tile_off(p1.I, p1.J), lane, c1);
        } else {
            sbr8_tile_update(c0, ra0, cb0);
            tm.store(tile_off(p0.I, p0.J), lane, c0);
        }
        c0 = d0;
        c1 = d1;
        p0 = q0;
        p1 = q1;
        two = n1;
    }
}

// sum of eight values over the warp, every lane gets every total: three folding levels (4 + 2 + 1 exchanges), two
// plain ones, eight broadcasts
__device__ __forceinline__ void warp_sum8(double (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    double f4[4], f2[2], f1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double send = b4 ? v[j] : v[j + 4], keep = b4 ? v[j + 4] : v[j];
        f4[j] = keep + shfl_xor_d(send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double send = b3 ? f4[j] : f4[j + 2], keep = b3 ? f4[j + 2] : f4[j];
        f2[j] = keep + shfl_xor_d(send, 8);
    }
    {
        const double send = b2 ? f2[0] : f2[1], keep = b2 ? f2[1] : f2[0];
        f1 = keep + shfl_xor_d(send, 4);
    }
    f1 += shfl_xor_d(f1, 2);
    f1 += shfl_xor_d(f1, 1);
    // lane l now holds the total of index 4 * bit4 + 2 * bit3 + bit2
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, f1, ((i & 4) << 2) | ((i & 2) << 2) | ((i & 1) << 2));
}

// LQ of the 8 x p0 panel (buffer rows k = 0..7 of Ub, k-major, stride st) by one warp, in registers.
// Writes U over the panel, T (row-major 8 x 8, upper triangular) and the R block of the band output.
template <int NC>
__device__ __forceinline__ void sbr8_panel_lq(double* __restrict__ Ub, double* __restrict__ Tm, double* __restrict__ Bd,
                                              int p0, int st, int lane) {
    double p[8][NC];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int c = p0 - 1 - (lane + 32 * q);
            p[k][q] = (c >= 0) ? Ub[k * st + c] : 0.0;
        }
    double bandv[8];
    double tr[8];  // row `lane` of T (lanes 0..7), built column by column: T[:k, k] = -tau_k T[:k, :k] (U^T u_k)
#pragma unroll
    for (int j = 0; j < 8; ++j) tr[j] = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // x: row k left of its pivot (relative column index lane + 32 q > k)
        double x[NC];
        x[0] = (lane > k) ? p[k][0] : 0.0;
#pragma unroll
        for (int q = 1; q < NC; ++q) x[q] = p[k][q];
        double v[8];
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < NC; ++q) s = fma(x[q], p[k2][q], s);
            v[k2] = s;
        }
        double pv[8];  // the pivot column's entries (lane k, slot 0)
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) pv[k2] = __shfl_sync(0xffffffffu, p[k2][0], k);
        warp_sum8(v, lane);
        const double xnorm2 = v[k], alpha = pv[k];
        double beta = alpha, tau = 0.0, vscale = 0.0;
        if (xnorm2 > 1e-280) {  // |x| <= n after the power-of-four scaling; a row below 1e-140 is left alone
            const double s2 = fma(alpha, alpha, xnorm2);
            const double rs = fast_rsqrt(s2);
            const double nrm = s2 * rs;
            beta = -copysign(nrm, alpha);
            tau = fma(fabs(alpha), rs, 1.0);                        // (beta - alpha) / beta
            vscale = copysign(fast_rcp(fabs(alpha) + nrm), alpha);  // 1 / (alpha - beta)
        }
        // band row r = p0 + 7 - k: beta at distance 8 (lane k), the entries right of the pivot (lanes < k) are final;
        // kept in a register and stored after the chain
        bandv[k] = (lane == k) ? beta : p[k][0];
#pragma unroll
        for (int q = 0; q < NC; ++q) x[q] *= vscale;  // u
        if (lane == k) x[0] = (tau != 0.0) ? 1.0 : 0.0;
        double ts = 0.0;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            const double dot = fma(vscale, v[k2], pv[k2]);  // row k2 . u_k (pivot entry included)
            if (k2 > k) {
                const double coef = tau * dot;
#pragma unroll
                for (int q = 0; q < NC; ++q) p[k2][q] = fma(-coef, x[q], p[k2][q]);
            } else if (k2 < k) {
                ts = fma(tr[k2], dot, ts);  // (T[lane][:k]) . (U^T u_k)
            }
        }
        tr[k] = (k > lane) ? -tau * ts : ((k == lane) ? tau : 0.0);
#pragma unroll
        for (int q = 0; q < NC; ++q) p[k][q] = x[q];
    }
    if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) Tm[lane * 8 + j] = tr[j];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (lane <= k) Bd[(p0 + 7 - k) * kBandW + 8 - k + lane] = bandv[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int c = p0 - 1 - (lane + 32 * q);
            if (c >= 0) Ub[k * st + c] = p[k][q];
        }
}

// Y = A U contributions of one stored block of TB x TB tiles (tile rows 4... RT0.., tile columns CT0..).
//   rows:    D[row g][kk] += sum_cols tile[g][col] U[kk][col]   the tile's two registers are the A fragments of the
//                                                              k-slices {cols 2t} and {cols 2t+1}
//   columns: D[col g][kk] += sum_rows tile[row][g] U[kk][row]   transposed fragments by four shuffles
// Diagonal tiles are stored full: row sums only.  Tile rows >= nt do not exist.
template <int TB>
__device__ __forceinline__ void sbr8_symm_block(const TileMem<TB == 4>& tm, bool diag, int RT0, int CT0, int tr0, int tre,
                                                int nt, const double* __restrict__ Ub, int st, int lane, int g, int t,
                                                double (&rowsum)[TB][2], double (&colsum)[TB][2]) {
    double2 c[TB][TB];
#pragma unroll
    for (int tr = 0; tr < TB; ++tr)
#pragma unroll
        for (int tc = 0; tc < TB; ++tc) {
            const bool ok = tr >= tr0 && tr < tre && RT0 + tr < nt && !(diag && tc > tr);
            c[tr][tc] = ok ? tm.load(tile_off(RT0 + tr, CT0 + tc), lane) : make_double2(0.0, 0.0);
        }
    double2 bu[TB];  // U[kk = g][cols 2t, 2t+1 of tile column tc]
#pragma unroll
    for (int tc = 0; tc < TB; ++tc) bu[tc] = *reinterpret_cast<const double2*>(Ub + g * st + 8 * (CT0 + tc) + 2 * t);
    const int src0 = 4 * t + (g >> 1), src1 = src0 + 16;  // lanes that hold tile[t][g], tile[4 + t][g]
    const bool odd = g & 1;
#pragma unroll
    for (int tr = 0; tr < TB; ++tr) {
        if (tr < tr0 || tr >= tre || RT0 + tr >= nt) continue;  // warp-uniform
        const int R0 = 8 * (RT0 + tr);
        const double bur0 = Ub[g * st + R0 + t], bur1 = Ub[g * st + R0 + 4 + t];
        double da0 = 0.0, da1 = 0.0, db0 = 0.0, db1 = 0.0;  // two chains for the row sums
#pragma unroll
        for (int tc = 0; tc < TB; ++tc) {
            if (diag && tc > tr) continue;  // warp-uniform
            if (tc & 1) {
                dmma8(db0, db1, c[tr][tc].x, bu[tc].x);
                dmma8(db0, db1, c[tr][tc].y, bu[tc].y);
            } else {
                dmma8(da0, da1, c[tr][tc].x, bu[tc].x);
                dmma8(da0, da1, c[tr][tc].y, bu[tc].y);
            }
            if (diag && tc == tr) continue;  // full symmetric tile: the row sums are its whole contribution
            const double x0 = __shfl_sync(0xffffffffu, c[tr][tc].x, src0), y0 = __shfl_sync(0xffffffffu, c[tr][tc].y, src0);
            const double x1 = __shfl_sync(0xffffffffu, c[tr][tc].x, src1), y1 = __shfl_sync(0xffffffffu, c[tr][tc].y, src1);
            const double e0 = odd ? y0 : x0, e1 = odd ? y1 : x1;
            dmma8(colsum[tc][0], colsum[tc][1], e0, bur0);
            dmma8(colsum[tc][0], colsum[tc][1], e1, bur1);
        }
        rowsum[tr][0] += da0 + db0;
        rowsum[tr][1] += da1 + db1;
    }
}

// NW warps, MINB CTAs per SM, NC = 32-column slots of the LQ warp (>= ceil((m_start - 8) / 32)), TB = tiles per
// index-block side of the cyclic products schedule (2: blocks of 16, 4: blocks of 32): NW * 8 TB >= m_start - 8.
template <int NW, int MINB, int NC, int TB>
__global__ void __launch_bounds__(32 * NW, MINB)
    sbr8_kernel(const ItemDesc* __restrict__ items, int item_base, double* __restrict__ ws, int st, int m_start, int m_stop,
                int tiles_smem) {
    extern __shared__ __align__(16) double smem[];
    const ItemDesc it = items[item_base + blockIdx.x];
    const int n = it.n;
    const int N = sbr8_order(n), off = N - n;
    const int tid = threadIdx.x;
    constexpr int nthreads = 32 * NW;
    const int lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, t = lane & 3;  // DMMA fragment coordinates
    constexpr bool kPair = (TB == 4);       // the two-CTA-per-SM instantiation (168 registers): two tiles in flight

    double* Vb = smem;              // [8][st] previous reflectors (pending update); partner sums during the products
    double* Wb = Vb + 8 * st;       // [8][st] previous W; partner sums / Y / X / new W
    double* Ub = Wb + 8 * st;       // [8][st] panel buffer -> current reflectors
    double* Tm = smem + 24 * st;    // [64]
    double* Zpart = Tm + 64;        // [NW][64]
    double* Sscr = Zpart + 64 * NW; // [NW][64] per-warp scratch (Z, S as B fragments)
    double* A = Sscr + 64 * NW;     // tiles

    double* __restrict__ G = ws + it.gram_off;
    double* out = ws + it.de_off;
    double* __restrict__ Bd = G + sbr8_band_off(n);  // band output, frame indices
    const int n0 = m_start > 0 ? m_start : N;            // order this launch starts from
    const int nt0 = n0 >> 3;
    const int tot64 = tile_off(nt0, 0);                  // doubles of this launch's triangle
    const int ts64 = (tiles_smem << 6) < tot64 ? (tiles_smem << 6) : tot64;  // ... of which in shared memory
    const TileMem<TB == 4> tm{A, G, ts64};
    if (m_start > 0) {
        if (out[2 * n + MISC_FLAGS] != 0.0) return;  // non-finite / all-zero: flagged by the first launch
        for (int i = tid; i < ts64; i += nthreads) A[i] = G[i];  // handed over in tile order
    } else {
        // ---- condition the Gram matrix (power-of-four scale so that |G_ij| <= 1 and the singular values un-scale
        //      exactly; NaN/Inf anywhere in W shows on the Gram diagonal).  The Gram kernels wrote the tile layout.
        double md = 0.0;
        int bad = 0;
        for (int c = lane; c < n; c += 32) {
            const int f = c + off;
            const double gd = G[tile_off(f >> 3, f >> 3) + 9 * (f & 7)];
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
        double2* G2 = reinterpret_cast<double2*>(G);  // flat, coalesced 128-bit copy; the global share is scaled in place
        double2* A2 = reinterpret_cast<double2*>(A);
        for (int i = tid; i < (tot64 >> 1); i += nthreads) {
            double2 v = G2[i];
            v.x *= scale;
            v.y *= scale;
            if (2 * i < ts64)
                A2[i] = v;
            else
                G2[i] = v;
        }
    }
    __syncthreads();

#ifdef VSP_PHASE_TIMING  // per-phase cycle counters of warp 0 / warp 1 (development builds only)
    long long t_phase[6] = {0, 0, 0, 0, 0, 0};  // mini | LQ or update | wait for the other | products | W-phase | panels
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
    while (m >= 16 && m > m_stop) {
        const int p0 = m - 8, Ip = p0 >> 3;
        // ---- (1) mini-pass: tile row Ip with the pending update; diagonal tile -> band, the rest -> panel buffer
        {
            // a warp's tiles J = warp, warp + NW, ... (at most kMini of them): all loads first (they may come from L2)
            constexpr int kMini = kPair ? (kSbr8MaxN / 8 + NW - 1) / NW : 1;  // (register-capped instantiations: one at a time)
            Frag8 ra;
            if (pending) ra = sbr8_frag(Vb, Wb, st, p0, g, t);
            for (int Jb = warp; Jb <= Ip; Jb += kMini * NW) {
            double2 cm[kMini];
#pragma unroll
            for (int q = 0; q < kMini; ++q) {
                const int J = Jb + q * NW;
                cm[q] = (J <= Ip) ? tm.load(tile_off(Ip, J), lane) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < kMini; ++q) {
                const int J = Jb + q * NW;
                if (J > Ip) continue;  // warp-uniform
                double2 c = cm[q];
                if (pending) {
                    const Frag8 cb = sbr8_frag(Vb, Wb, st, 8 * J, g, t);
                    sbr8_tile_update(c, ra, cb);
                }
                if (J < Ip) {
                    *reinterpret_cast<double2*>(Ub + (7 - g) * st + 8 * J + 2 * t) = c;
                } else {
                    double* brow = Bd + (p0 + g) * kBandW + g;
                    if (2 * t <= g) brow[-2 * t] = c.x;
                    if (2 * t + 1 <= g) brow[-2 * t - 1] = c.y;
                }
            }
            }
        }
        __syncthreads();
        VSP_LAP(0);

        // ---- (2) LQ of the panel by warp 0, the pending update of the leading p0 x p0 triangle by the others
        // The LQ warp is the LAST one (the issue arbiter of an SM sub-partition prefers the highest warp id, and the LQ
        // chain is what the panel step waits for), and the warps that share its sub-partition (warp id = 3 mod 4)
        // sit the update out: DMMA and DFMA use the same FP64 units, a DMMA holds them for 16 cycles, and every
        // one of the ~130 dependent FP64 instructions of a reflector queued behind the neighbours' DMMAs
        // (measured: LQ 10.4 k cycles per panel with eleven update warps).
        if (warp == NW - 1) {
            sbr8_panel_lq<NC>(Ub, Tm, Bd, p0, st, lane);
            if (NW == 1 && pending) sbr8_update_sweep<kPair>(tm, Ip, 0, 1, Vb, Wb, st, lane, g, t);
        } else if (pending) {
            if (kQuietLq && NW >= 8) {
                if ((warp & 3) != 3) sbr8_update_sweep<kPair>(tm, Ip, warp - (warp >> 2), NW - NW / 4, Vb, Wb, st, lane, g, t);
            } else {
                sbr8_update_sweep<kPair>(tm, Ip, warp, NW - 1, Vb, Wb, st, lane, g, t);
            }
        }
        VSP_LAP(1);
        __syncthreads();
        VSP_LAP(2);

        // ---- (3) symmetric products Y = A U over the leading Ip x Ip tile triangle (cyclic block-pair schedule)
        const int nb = (Ip + TB - 1) / TB;  // index blocks of 8 TB
        {
            double own[TB][2];  // sums of the own index block as D fragments: [8-row group][kk = 2t, 2t+1]
#pragma unroll
            for (int i = 0; i < TB; ++i) own[i][0] = own[i][1] = 0.0;
            const int nsteps = (nb >> 1) + 1;
            // partner sums of step s go to buffer s mod 2 (the previous V and W are free from here to the W-phase):
            // every step touches every block of its buffer exactly once, so steps 0 and 1 store, later ones add,
            // with a barrier before every even step
            double* const Ybuf[2] = {Wb, Vb};
            for (int s = 0; s < nsteps; ++s) {
                if (s > 0 && (s & 1) == 0) __syncthreads();
                if (warp < nb) {
                    int o = warp + s;
                    if (o >= nb) o -= nb;
                    const bool half = (2 * s == nb);  // the pair {w, w + nb/2} is met from both sides
                    const bool own_cols = (s == 0) || (o > warp);
                    const int rb = own_cols ? o : warp, cb = own_cols ? warp : o;  // stored block (rb, cb)
                    const int tr0 = (half && !own_cols) ? TB / 2 : 0, tre = (half && own_cols) ? TB / 2 : TB;
                    double oth[TB][2];
#pragma unroll
                    for (int i = 0; i < TB; ++i) oth[i][0] = oth[i][1] = 0.0;
                    if (own_cols)
                        sbr8_symm_block<TB>(tm, s == 0, TB * rb, TB * cb, tr0, tre, Ip, Ub, st, lane, g, t, oth, own);
                    else
                        sbr8_symm_block<TB>(tm, s == 0, TB * rb, TB * cb, tr0, tre, Ip, Ub, st, lane, g, t, own, oth);
                    double* const yb = Ybuf[s & 1];
#pragma unroll
                    for (int i = 0; i < TB; ++i) {
                        if (TB * o + i >= Ip) continue;  // strip beyond the active block (warp-uniform)
                        double* y0 = yb + (2 * t) * st + 8 * (TB * o + i) + g;
                        if (s < 2) {  // first touch of this block in this buffer
                            y0[0] = oth[i][0];
                            y0[st] = oth[i][1];
                        } else {
                            y0[0] += oth[i][0];
                            y0[st] += oth[i][1];
                        }
                    }
                }
            }
            __syncthreads();
            if (warp < nb) {
#pragma unroll
                for (int i = 0; i < TB; ++i) {
                    if (TB * warp + i >= Ip) continue;
                    double* y0 = Wb + (2 * t) * st + 8 * (TB * warp + i) + g;  // all steps are done: exclusive again
                    y0[0] += own[i][0];
                    y0[st] += own[i][1];
                }
            }
        }
        __syncthreads();
        VSP_LAP(3);

        // ---- (4) W-phase per 8-row strip: X = Y T, Z = U^T X (CTA reduction), S = T^T Z, W = X - U S / 2
        {
            constexpr int kMaxStrips = (8 * TB * NW / 8 + NW - 1) / NW;  // = TB
            double2 xs[kMaxStrips];
            double z0 = 0.0, z1 = 0.0;  // this warp's part of Z as a D fragment: Z[i = g][j = 2t, 2t+1]
            const bool two = nb >= 2;   // partner sums of the odd steps live in Vb
            double tb0 = Tm[t * 8 + g], tb1 = Tm[(4 + t) * 8 + g];  // T as B fragments: T[k = 4 s + t][j = g]
#pragma unroll
            for (int q = 0; q < kMaxStrips; ++q) {
                const int R = warp + q * NW;
                xs[q] = make_double2(0.0, 0.0);
                if (R < Ip) {  // warp-uniform
                    const int i0 = 8 * R;
                    double y0 = Wb[t * st + i0 + g], y1 = Wb[(4 + t) * st + i0 + g];
                    if (two) {
                        y0 += Vb[t * st + i0 + g];
                        y1 += Vb[(4 + t) * st + i0 + g];
                    }
                    dmma8(xs[q].x, xs[q].y, y0, tb0);
                    dmma8(xs[q].x, xs[q].y, y1, tb1);
                    __syncwarp();  // every lane has read this strip's Y before X overwrites it
                    Wb[(2 * t) * st + i0 + g] = xs[q].x;  // X, k-major, for the B fragments below
                    Wb[(2 * t + 1) * st + i0 + g] = xs[q].y;
                    __syncwarp();
                    // Z[i][j] += sum_rows U[i][row] X[row][j]
                    dmma8(z0, z1, Ub[g * st + i0 + t], Wb[g * st + i0 + t]);
                    dmma8(z0, z1, Ub[g * st + i0 + 4 + t], Wb[g * st + i0 + 4 + t]);
                }
            }
            *(reinterpret_cast<double2*>(Zpart + 64 * warp) + lane) = make_double2(z0, z1);
            __syncthreads();
            double2 zt = make_double2(0.0, 0.0);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const double2 zp = *(reinterpret_cast<const double2*>(Zpart + 64 * w) + lane);
                zt.x += zp.x;
                zt.y += zp.y;
            }
            double* myS = Sscr + 64 * warp;
            *(reinterpret_cast<double2*>(myS) + lane) = zt;  // Z row-major
            __syncwarp();
            // S[i][j] = sum_k T[k][i] Z[k][j]: A = T^T (lane: T[4 s + t][g]), B = Z[4 s + t][g]
            double s0 = 0.0, s1 = 0.0;
            dmma8(s0, s1, tb0, myS[t * 8 + g]);
            dmma8(s0, s1, tb1, myS[(4 + t) * 8 + g]);
            __syncwarp();
            *(reinterpret_cast<double2*>(myS) + lane) = make_double2(-0.5 * s0, -0.5 * s1);  // -S / 2 row-major
            __syncwarp();
            const double sb0 = myS[t * 8 + g], sb1 = myS[(4 + t) * 8 + g];  // B fragments: (-S/2)[i = 4 s + t][j = g]
#pragma unroll
            for (int q = 0; q < kMaxStrips; ++q) {
                const int R = warp + q * NW;
                if (R < Ip) {
                    const int i0 = 8 * R;
                    // W[row][j] = X[row][j] + sum_i U[i][row] (-S/2)[i][j]
                    dmma8(xs[q].x, xs[q].y, Ub[t * st + i0 + g], sb0);
                    dmma8(xs[q].x, xs[q].y, Ub[(4 + t) * st + i0 + g], sb1);
                    Wb[(2 * t) * st + i0 + g] = xs[q].x;
                    Wb[(2 * t + 1) * st + i0 + g] = xs[q].y;
                }
            }
        }
        // V <- U; the old V buffer becomes the next panel buffer
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
    if (blockIdx.x == 200 && lane == 0 && (warp == 0 || warp == NW - 1))
        printf("[sbr8 n=%d from %d warp %d] cycles: mini %lld  LQ/update %lld  wait %lld  products %lld  W %lld  panels %lld\n", n, n0,
               warp, t_phase[0], t_phase[1], t_phase[2], t_phase[3], t_phase[4], t_phase[5]);
#endif
#undef VSP_LAP
    const int ntm = m >> 3;
    if (m >= 16) {
        // ---- hand-over: bring the leading m x m block up to date and return it to the workspace in tile order
        if (pending) sbr8_update_sweep<kPair>(tm, ntm, warp, NW, Vb, Wb, st, lane, g, t);
        __syncthreads();
        const int cnt = tile_off(ntm, 0) < ts64 ? tile_off(ntm, 0) : ts64;  // the global share is already in place
        for (int i = tid; i < cnt; i += nthreads) G[i] = A[i];
        return;
    }
    // ---- m == 8: the leading tile is inside the band
    if (warp == 0) {
        double2 c = *(reinterpret_cast<const double2*>(A) + lane);
        if (pending) {
            const Frag8 ra = sbr8_frag(Vb, Wb, st, 0, g, t);
            sbr8_tile_update(c, ra, ra);
        }
        double* brow = Bd + g * kBandW + g;
        if (2 * t <= g) brow[-2 * t] = c.x;
        if (2 * t + 1 <= g) brow[-2 * t - 1] = c.y;
    }
}

#endif  // __CUDACC__

}  // namespace vsp
