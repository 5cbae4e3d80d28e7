// rmx_tc.cuh -- FP64 tensor-core (DMMA.8x8x4, mma.sync m8n8k4 f64) forms of the two hot phases of the one-warp forward
// kernel: assembly of the Newton matrix from the composite blocks, and its partial-pivot LU.
//
// Why: the scalar forms (rmx_fast.cuh eval_columns2 / lu_solve_warp_t) spend 2 x 32-bit shuffles plus register moves per
// trailing LU entry and 18 broadcast multiply-adds per matrix entry; measured on B200 (profiles/r01_dmma_dfma_shfl_probe.log)
// one DMMA (256 multiply-adds, 1 issue slot) costs the FP64 pipe what 8 DFMAs cost, and a shuffle occupies the SM-wide
// shuffle unit for 2 cycles.  Here
//   * H = [L | S] [Rt ; Z] restricted by the tree relation is 8x8 output tiles of k = 12 (18 with ground contact) + 8:
//     5 (7) DMMAs per tile, operands fetched as fragments straight from two joint-major arrays in shared memory;
//   * the LU is right-looking in 8-column panels: rows never move (perm[] / rem[] indirection), a panel is factored in
//     registers (8 entries per lane) with LAPACK's idamax tie rule, U12 by one forward substitution per trailing column,
//     and the trailing update A22 -= L21 U12 by 2 DMMAs per 8x8 tile on the shared-memory image of H.
// H is padded with an identity block to a multiple of 8 rows/columns (the assembly writes the padding), its leading
// dimension and all strides are compile-time constants, and every index is a 32-bit offset into shared memory, so the
// loops below carry no bounds checks and almost no address arithmetic.
// tools/proto_dmma_lu.py emulates the LU lane by lane (fragment layouts included) and checks pivots and factors against
// LAPACK.  PTX fragment layout (g = lane >> 2, t = lane & 3): A[g][t], B[t][g], C[g][2t], C[g][2t+1].
#pragma once
#include "rmx_fast.cuh"

namespace rmx {

// Reciprocal of a pivot on the factorisation's critical path: the hardware seed (MUFU.RCP64H, relative error < 2^-20) and two
// Newton steps, accurate to an ulp or so -- __drcp_rn's correctly rounded result costs two more dependent FMAs and a
// slow-path test per pivot, and the multipliers it would make bitwise equal to LAPACK's are not compared with anything.
__device__ __forceinline__ double rcp_pivot(double x) {
#ifdef RMX_PIVOT_EXACT
    return __drcp_rn(x);
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
#endif
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // volatile: a .sync.aligned instruction must stay where the warp is converged (never sunk into a divergent consumer)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// eval_columns_tc: same result as eval_columns2 (out = scale * (sq dg/dq + sqd dg/dqdot + sd dg/d(dqtmp))), NW = 1 or 2
// warps (thread = joint in the per-joint part, the tile rows are dealt round-robin to the warps), plus the identity padding
// of rows / columns nr .. 8*ceil(nr/8)-1.
// ---------------------------------------------------------------------------------------------
// T: the shared-memory layout (TcLayout<GROUND, NW> of the forward kernels, TcLayoutA of the adjoint forward kernel).
// TOGLOBAL: the tiles go straight to `out` in GLOBAL memory, nr x nr ROW-major (the adjoint tape's M and D), with the
// joint-level diagonal term folded into the epilogue; otherwise `out` is the shared-memory image of H (column-major, T::LD).
template <int NW, int GROUND, class T, bool TOGLOBAL>
__device__ __forceinline__ void eval_columns_tc(Ctx2& c, double sq, double sqd, double sd, double scale, double* out) {
    typedef typename T::F F;
    typedef typename TcMask<NW>::type mask_t;
    constexpr int NL = F::NL, NWD = F::NW_, LD = T::LD;
    constexpr int KS1 = (NL + 3) / 4;  // k-steps of the subtree part (k = NL, zero padded)
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int n = c.n, nr = c.nr;
    const double cc = c.c;
    __builtin_assume(__isShared(c.sa));
    if (!TOGLOBAL) __builtin_assume(__isShared(out));
    const int myidx = (tid < n) ? c.ie_s[tid].x : -1;
    double* __restrict__ Wb = c.sa + T::W_OFF;
    double* __restrict__ RZb = c.sa + T::RZ_OFF;
    // Two-warp forward kernels keep 32 rows of RZ at a time (T::RZ_ROWS): the column joints are taken in two halves, the upper
    // threads holding their RZ rows in registers through the first one -- 4.6 KB less per block, which is what lets a fourth
    // 64-link rollout reside on an SM.
    constexpr bool HALF = T::RZ_ROWS < T::CAP;
    double Rt[NL], Z[6];
    auto put_rz = [&](int row, bool has_dof) {
        double* RZ = RZb + row * NWD;
#pragma unroll
        for (int i = 0; i < NL; ++i) RZ[i] = has_dof ? Rt[i] : 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) RZ[NL + i] = has_dof ? Z[i] : 0.0;
    };
    {
        double L[NL], s[6];
        columns_joint<NW, GROUND, T::KEEP>(c, tid, myidx, sq, sqd, sd, L, s, Rt, Z);
        bsync<NW>();  // every thread has read S, V, U and its composite blocks: the SoA block may be overwritten
        if (tid < n) {
            double* W = Wb + tid * NWD;
            if (myidx >= 0) {
#pragma unroll
                for (int i = 0; i < NL; ++i) W[i] = L[i];
#pragma unroll
                for (int i = 0; i < 6; ++i) W[NL + i] = s[i];
            } else {  // fixed joint: no row / column of H
#pragma unroll
                for (int i = 0; i < NWD; ++i) W[i] = 0.0;
            }
            if (!HALF || tid < 32) put_rz(tid, myidx >= 0);
        }
    }
    bsync<NW>();
    const int nT = (n + 7) >> 3;
    const int* __restrict__ idxs = c.tcidx_s;
    const mask_t* __restrict__ subs = reinterpret_cast<const mask_t*>(c.tcsub_s);
    const mask_t* __restrict__ ancs = reinterpret_cast<const mask_t*>(c.tcanc_s);
    __builtin_assume(__isShared(idxs));
    __builtin_assume(__isShared(subs));
    __builtin_assume(__isShared(ancs));
    // rows >= n of W / RZ are never written: whatever is there only reaches C rows / columns >= n, which are not stored
    for (int half = 0; half < (HALF ? 2 : 1); ++half) {
        if (HALF && half == 1) {
            if (nT <= 4) break;  // block-uniform
            bsync<NW>();         // every tile of the first half has read its RZ rows
            if (tid >= 32 && tid < n) put_rz(tid - 32, myidx >= 0);
            bsync<NW>();
        }
        const int J0 = HALF ? 4 * half : 0;
        const int J1 = HALF ? (nT < 4 * half + 4 ? nT : 4 * half + 4) : nT;
    for (int I = (NW == 1 ? 0 : warp); I < nT; I += NW) {
        const int k = 8 * I + g;  // row joint of this lane's A fragments and C elements
        const double* wk = Wb + k * NWD + t4;
        double a1[KS1], a2[2];
#pragma unroll
        for (int ks = 0; ks < KS1; ++ks) a1[ks] = (4 * ks + 3 < NL || 4 * ks + t4 < NL) ? wk[4 * ks] : 0.0;
        a2[0] = wk[NL];
        a2[1] = (t4 < 2) ? wk[NL + 4] : 0.0;
        // tree relation of row joint k with every column joint, and its reduced index (no row if fixed or k >= n)
        const int idxk = (k < n) ? idxs[k] : -1;
        const mask_t subk = (idxk >= 0) ? subs[k] : (mask_t)0;
        const mask_t anck = (idxk >= 0) ? ancs[k] : (mask_t)0;
        double* orow = TOGLOBAL ? out + (idxk >= 0 ? idxk : 0) * nr : out + idxk;
        // joint stiffness / damping / limit terms Kr, Dr (Joint.m:470-481) of this row's diagonal entry
        const double dgk = (TOGLOBAL && idxk >= 0) ? -cc * (sq * c.sp2[idxk] + sqd * c.sp1[idxk]) : 0.0;
        // epilogue of one tile: pick subtree / ancestor / zero per entry (one bit test each) and store
        auto store_tile = [&](int J, double s0, double s1, double z0, double z1) {
            const int i0 = 8 * J + 2 * t4;  // column joints i0, i0+1 of this lane's C elements (tcidx_s has CAP entries)
            const int2 ix = *reinterpret_cast<const int2*>(idxs + i0);
            const unsigned sb = (unsigned)(subk >> i0), ab = (unsigned)(anck >> i0);
            const double v0 = (sb & 1u) ? s0 : ((ab & 1u) ? z0 : 0.0);  // k in sub(i): L_k . Rt_i ; k proper ancestor of i: s_k . Z_i
            const double v1 = (sb & 2u) ? s1 : ((ab & 2u) ? z1 : 0.0);
            if (idxk >= 0) {
                if (TOGLOBAL) {
                    if (i0 < n && ix.x >= 0) orow[ix.x] = scale * (ix.x == idxk ? v0 + dgk : v0);
                    if (i0 + 1 < n && ix.y >= 0) orow[ix.y] = scale * (ix.y == idxk ? v1 + dgk : v1);
                } else {
                    if (i0 < n && ix.x >= 0) orow[ix.x * LD] = scale * v0;
                    if (i0 + 1 < n && ix.y >= 0) orow[ix.y * LD] = scale * v1;
                }
            }
        };
        // two tiles of the row at a time: their DMMA chains (3 or 5 + 2 dependent instructions each, 26 cycles apart) interleave
        for (int J = J0; J < J1; J += 2) {
            const bool two = J + 1 < J1;  // warp-uniform
            const double* rzA = RZb + (8 * (J - J0) + g) * NWD + t4;  // column joint 8J+g of this lane's B fragments
            const double* rzB = two ? rzA + 8 * NWD : rzA;
            double sA0 = 0.0, sA1 = 0.0, zA0 = 0.0, zA1 = 0.0, sB0 = 0.0, sB1 = 0.0, zB0 = 0.0, zB1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS1; ++ks) {
                const bool in = 4 * ks + 3 < NL || 4 * ks + t4 < NL;
                const double bA = in ? rzA[4 * ks] : 0.0;
                const double bB = in ? rzB[4 * ks] : 0.0;
                dmma884(sA0, sA1, a1[ks], bA);
                if (two) dmma884(sB0, sB1, a1[ks], bB);
                if (ks < 2) {  // the ancestor part (k = 6, two steps) rides along
                    const bool inz = ks == 0 || t4 < 2;
                    const double cA = inz ? rzA[NL + 4 * ks] : 0.0;
                    const double cB = inz ? rzB[NL + 4 * ks] : 0.0;
                    dmma884(zA0, zA1, a2[ks], cA);
                    if (two) dmma884(zB0, zB1, a2[ks], cB);
                }
            }
            store_tile(J, sA0, sA1, zA0, zA1);
            if (two) store_tile(J + 1, sB0, sB1, zB0, zB1);
        }
    }
    }
    bsync<NW>();
    if (TOGLOBAL) return;
    // diagonal: joint stiffness / damping / limit terms Kr, Dr (Joint.m:470-481)
    if (myidx >= 0) out[myidx * (LD + 1)] += scale * (-cc * (sq * c.sp2[myidx] + sqd * c.sp1[myidx]));
    if (GROUND == 2 && c.npf > 0) {  // off-diagonal blocks of the point forces; RZ row = [c2 ; c1 ; sq s ; Z]
        const double* rzl = RZb + tid * NWD;
        pf_cross_pass(c, tid, myidx, rzl + 6, rzl + 12, scale, out, LD, Wb, NWD, NL);
    }
    // identity padding up to a multiple of 8 (the blocked LU runs whole panels and whole tiles)
    const int np8 = (nr + 7) & ~7;
    for (int col = nr; col < np8; ++col) out[col * LD + tid] = (tid == col) ? 1.0 : 0.0;
    if (tid < nr)
        for (int row = nr; row < np8; ++row) out[tid * LD + row] = 0.0;
    bsync<NW>();
}

// ---------------------------------------------------------------------------------------------
// lu_solve_tc: dx = scale * H \ rhs for a block of NW = 1 or 2 warps, nr <= 32 NW.  H: column-major, leading dimension
// 32 NW + 1, padded with an identity block to np8 = 8*ceil(nr/8) rows and columns, in shared memory; overwritten by its factors
// (rows stay where they are: row perm[k] is the k-th pivot row, unit-lower multipliers left of the diagonal position).
// perm, rem: int[32 NW] shared; rowbuf: 2 x 5 double2 shared (pivot-row broadcast inside a panel).
// The panel factorisation (the pivot search is one dependent chain per column) runs in warp 0, lane l holding rows l and, with
// two warps, l + 32; U12 and the trailing update are spread over all warps.
// ---------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void lu_solve_tc(int nr, double* H, int* perm, int* rem, double2* rowbuf, const double* rhs,
                                            double scale, double* dx) {
    constexpr int LD = 32 * NW + 1;
    constexpr int R = NW;             // rows per lane of warp 0
    constexpr int MAXI = 4 * NW - 1;  // trailing tile rows / columns at most
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const unsigned ltmask = (1u << lane) - 1u;
    __builtin_assume(__isShared(H));
    __builtin_assume(__isShared(perm));
    __builtin_assume(__isShared(rem));
    __builtin_assume(__isShared(rowbuf));
    __builtin_assume(__isShared(rhs));
    __builtin_assume(__isShared(dx));
    const int NP = (nr + 7) >> 3;
    const int np8 = 8 * NP;
    const bool w0 = (NW == 1) || warp == 0;
    bool done[R];  // rows beyond the padded matrix never take part
    int pos[R], mypos[R];  // pos: LAPACK position of each row (RMX_PIVOT_EXACT only)
    unsigned kmask[R], kcode[R];  // pivot-search key of a row: (high word of the entry & kmask) | kcode
    double b[R], rdiag[R];
#pragma unroll
    for (int h = 0; h < R; ++h) {
        const int row = lane + 32 * h;
        done[h] = row >= np8;
        pos[h] = row;
        (void)pos[h];
        mypos[h] = -1;
        b[h] = (row < nr) ? scale * rhs[row] : 0.0;
        rdiag[h] = 1.0;
        kmask[h] = done[h] ? 0u : 0x7fffffc0u;
        kcode[h] = done[h] ? 0u : (0x80000000u | (unsigned)(63 - row));
        (void)kmask[h];
        (void)kcode[h];
    }
    double* Hrow = H + lane;  // this lane's rows: entry c of row lane + 32 h at Hrow[c * LD + 32 h]
    for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        const int nt = NP - p - 1;  // trailing tile rows == trailing tile columns
        if (w0) {
            // ---- panel: columns c0 .. c0+7 of every row, 8 entries per row ---------------------------------------------
            double* Hp = Hrow + c0 * LD;
            double a[R][8];
#pragma unroll
            for (int h = 0; h < R; ++h)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[h][i] = Hp[i * LD + 32 * h];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = c0 + i;
#ifndef RMX_PIVOT_EXACT
                // Every lane takes the reciprocal of its own candidate(s) now, while the reduction below is in flight: the one of
                // the row that wins is published with the pivot row, so the MUFU + two Newton steps (~50 cycles) leave the chain
                // reduction -> broadcast -> multiplier -> update.  Same number as the reciprocal of the broadcast pivot (the seed
                // instruction has no slow path, so zero / padding entries cost nothing), same instruction count.
                double rp_own[R];
#pragma unroll
                for (int h = 0; h < R; ++h) rp_own[h] = rcp_pivot(a[h][i]);
                // Pivot row: ONE warp reduction per column.  Key = [not a pivot yet | top 25 bits of |a[i]| (sign, exponent, 14
                // mantissa bits dropped of the IEEE high word's 20) | 63 - row]: its maximum names a row whose entry is within
                // 2^-14 of the largest in the column (the lowest such row), which is all partial pivoting needs -- the growth
                // bound changes by that factor.  The exact dgetf2 rule (idamax, first maximum in LAPACK's current row order: a second
                // reduction, two votes and a position exchange per column, all on the factorisation's critical path) is kept
                // behind -DRMX_PIVOT_EXACT; both give factors of the same quality, neither is bitwise MATLAB's.
                // On the chain: one LOP3 per row ((hi & kmask) | kcode; both words are zeroed once a row is a pivot), the
                // reduction, one compare -- the keys are distinct, so the winner recognises itself and nobody decodes the row.
                unsigned key[R];
#pragma unroll
                for (int h = 0; h < R; ++h) key[h] = ((unsigned)__double2hiint(a[h][i]) & kmask[h]) | kcode[h];
                const unsigned kmax = __reduce_max_sync(FULL, R == 1 ? key[0] : max(key[0], key[R - 1]));
                bool win[R];
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    win[h] = key[h] == kmax;  // (kmax != 0: some row is not a pivot yet)
                    if (win[h]) {
                        done[h] = true;
                        mypos[h] = k;
                        kmask[h] = 0u;
                        kcode[h] = 0u;
                        perm[k] = lane + 32 * h;
                    }
                }
#else
                // argmax |a[i]| over the rows that are not pivots yet; ties -> smallest LAPACK position (dgetf2's idamax).
                // |v| >= 0, so the IEEE bit pattern orders like the value: one warp reduction on the high words decides unless two
                // rows agree in their top 32 bits; only then the low words and the positions are consulted (warp-uniform branch).
                double v[R];
                unsigned hi[R];
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    v[h] = fabs(a[h][i]);
                    hi[h] = done[h] ? 0u : (unsigned)__double2hiint(v[h]);
                }
                const unsigned mh = __reduce_max_sync(FULL, R == 1 ? hi[0] : max(hi[0], hi[R - 1]));
                bool c1[R];
                unsigned cand[R];
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    c1[h] = !done[h] && hi[h] == mh;
                    cand[h] = __ballot_sync(FULL, c1[h]);
                }
                const bool multi = (R == 1) ? ((cand[0] & (cand[0] - 1)) != 0u)
                                            : (__popc(cand[0]) + __popc(cand[R - 1]) > 1);
                if (multi) {
                    unsigned lo[R];
#pragma unroll
                    for (int h = 0; h < R; ++h) lo[h] = c1[h] ? (unsigned)__double2loint(v[h]) : 0u;
                    const unsigned ml = __reduce_max_sync(FULL, R == 1 ? lo[0] : max(lo[0], lo[R - 1]));
                    bool c2[R];
                    unsigned pl = 0xffffu;
#pragma unroll
                    for (int h = 0; h < R; ++h) {
                        c2[h] = c1[h] && lo[h] == ml;
                        pl = c2[h] ? min(pl, (unsigned)pos[h]) : pl;
                    }
                    const unsigned pm = __reduce_min_sync(FULL, pl);
#pragma unroll
                    for (int h = 0; h < R; ++h) cand[h] = __ballot_sync(FULL, c2[h] && (unsigned)pos[h] == pm);
                }
                const int sh = (R == 2 && cand[0] == 0u) ? 1 : 0;  // pivot row = src + 32 sh (warp-uniform)
                const int src = __ffs(sh ? cand[R - 1] : cand[0]) - 1;
                // LAPACK swaps the pivot row with the row in position k: that row takes over the pivot row's position
                const int pos_src = __shfl_sync(FULL, sh ? pos[R - 1] : pos[0], src);
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    if (!done[h] && pos[h] == k) pos[h] = pos_src;
                    if (lane == src && h == sh) {
                        pos[h] = k;
                        done[h] = true;
                        mypos[h] = k;
                    }
                }
                if (lane == (k & 31)) perm[k] = src + 32 * sh;
                bool win[R];
#pragma unroll
                for (int h = 0; h < R; ++h) win[h] = lane == src && h == sh;
#endif
                // the pivot lane publishes its panel row (entries i.. as 128-bit pairs) and right-hand side; two alternating buffers,
                // so one __syncwarp per pivot step orders both the read-after-write and the next write-after-read
                double2* buf = rowbuf + 5 * (i & 1);
#ifndef RMX_PIVOT_EXACT
#define RMX_RP_SLOT(h) rp_own[h]
#else
#define RMX_RP_SLOT(h) 0.0
#endif
                if (win[0]) {
#pragma unroll
                    for (int j = i / 2; j < 4; ++j) buf[j] = make_double2(a[0][2 * j], a[0][2 * j + 1]);
                    buf[4] = make_double2(b[0], RMX_RP_SLOT(0));
                } else if (R == 2 && win[R - 1]) {
#pragma unroll
                    for (int j = i / 2; j < 4; ++j) buf[j] = make_double2(a[R - 1][2 * j], a[R - 1][2 * j + 1]);
                    buf[4] = make_double2(b[R - 1], RMX_RP_SLOT(R - 1));
                }
#undef RMX_RP_SLOT
                __syncwarp();
                double2 u2[4];
#pragma unroll
                for (int j = i / 2; j < 4; ++j) u2[j] = buf[j];
                const double2 ubrp = buf[4];
                const double ub = ubrp.x;
#ifndef RMX_PIVOT_EXACT
                const double rp = ubrp.y;
#else
                const double piv = (i & 1) ? u2[i / 2].y : u2[i / 2].x;
                const double rp = rcp_pivot(piv);
#endif
#pragma unroll
                for (int h = 0; h < R; ++h) {
                    rdiag[h] = win[h] ? rp : rdiag[h];
                    const double l = done[h] ? 0.0 : a[h][i] * rp;  // l == 0 for rows that are already pivots
                    a[h][i] = done[h] ? a[h][i] : l;
                    if (!(i & 1)) a[h][i + 1] = fma(-l, u2[i / 2].y, a[h][i + 1]);
#pragma unroll
                    for (int j = i / 2 + 1; j < 4; ++j) {
                        a[h][2 * j] = fma(-l, u2[j].x, a[h][2 * j]);
                        a[h][2 * j + 1] = fma(-l, u2[j].y, a[h][2 * j + 1]);
                    }
                    b[h] = fma(-l, ub, b[h]);
                }
            }
#pragma unroll
            for (int h = 0; h < R; ++h)
#pragma unroll
                for (int i = 0; i < 8; ++i) Hp[i * LD + 32 * h] = a[h][i];
            if (nt > 0) {
                // ---- compact list of the rows still to be eliminated: exactly 8*nt of them ----------------------------------
                const unsigned live0 = __ballot_sync(FULL, !done[0]);
                if (!done[0]) rem[__popc(live0 & ltmask)] = lane;
                if (R == 2) {
                    const unsigned live1 = __ballot_sync(FULL, !done[R - 1]);
                    if (!done[R - 1]) rem[__popc(live0) + __popc(live1 & ltmask)] = lane + 32;
                }
            }
        }
        if (nt == 0) break;  // block-uniform
        bsync<NW>();
        // ---- U12 = L11^-1 A12(pivot rows): thread = trailing column ------------------------------------------------
        int pr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pr[i] = perm[c0 + i];
        if (tid < 8 * nt) {
            double* colp = H + (c0 + 8 + tid) * LD;
            const double* Lp = H + c0 * LD;
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = colp[pr[i]];
#pragma unroll
            for (int i = 1; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < i; ++j) x[i] = fma(-Lp[j * LD + pr[i]], x[j], x[i]);
            }
#pragma unroll
            for (int i = 1; i < 8; ++i) colp[pr[i]] = x[i];
        }
        bsync<NW>();
        // ---- A22 -= L21 U12 : 8x8 tiles, 2 DMMAs each; the tile columns are dealt round-robin to the warps ---------------
        // pivot rows of this lane's k-indices t4 and 4 + t4 (B fragments)
        const int prt0 = (t4 == 0) ? pr[0] : (t4 == 1) ? pr[1] : (t4 == 2) ? pr[2] : pr[3];
        const int prt1 = (t4 == 0) ? pr[4] : (t4 == 1) ? pr[5] : (t4 == 2) ? pr[6] : pr[7];
        const double* La = H + (c0 + t4) * LD;  // multipliers of k-index t4 (column c0 + t4): row r at La[r]; 4 + t4 at La[4 LD + r]
        if (NW == 1) {
            // one warp, at most 3 trailing tile rows: the A fragments of all of them stay in registers across the tile columns
            double af0[MAXI], af1[MAXI];
            int rI[MAXI];
    #pragma unroll
            for (int I = 0; I < MAXI; ++I) {
                if (I < nt) {
                    const int r = rem[8 * I + g];
                    rI[I] = r;
                    af0[I] = -La[r];
                    af1[I] = -La[4 * LD + r];
                }
            }
            for (int J = (NW == 1 ? 0 : warp); J < nt; J += NW) {
                const int cJ = c0 + 8 + 8 * J;
                const double* Ub = H + (cJ + g) * LD;  // B fragments: U12[k-index][column cJ + g]
                const double bf0 = Ub[prt0], bf1 = Ub[prt1];
                double* Cc = H + (cJ + 2 * t4) * LD;   // C elements: columns cJ + 2 t4, + 1
                double v0[MAXI], v1[MAXI];
    #pragma unroll
                for (int I = 0; I < MAXI; ++I)
                    if (I < nt) {
                        v0[I] = Cc[rI[I]];
                        v1[I] = Cc[LD + rI[I]];
                    }
    #pragma unroll
                for (int I = 0; I < MAXI; ++I)  // first k-step of every tile of this column, then the second: independent chains
                    if (I < nt) dmma884(v0[I], v1[I], af0[I], bf0);
    #pragma unroll
                for (int I = 0; I < MAXI; ++I)
                    if (I < nt) dmma884(v0[I], v1[I], af1[I], bf1);
    #pragma unroll
                for (int I = 0; I < MAXI; ++I)
                    if (I < nt) {
                        Cc[rI[I]] = v0[I];
                        Cc[LD + rI[I]] = v1[I];
                    }
            }
        } else {
            // two warps, up to 7 trailing tile rows: a predicated, fully unrolled form would issue all 7 row slots for every
            // tile column whatever nt is; here the row tiles are a plain loop (A fragments loaded once per row tile), the tile
            // columns are dealt round-robin to the warps and taken two at a time so that their DMMA chains overlap
            for (int I = 0; I < nt; ++I) {
                const int r = rem[8 * I + g];
                const double a0 = -La[r], a1 = -La[4 * LD + r];
                for (int J = warp; J < nt; J += 2 * NW) {
                    const bool two = J + NW < nt;  // warp-uniform
                    const int cJ = c0 + 8 + 8 * J;
                    const double* Ub = H + (cJ + g) * LD;  // B fragments: U12[k-index][column cJ + g]
                    double* Cc = H + (cJ + 2 * t4) * LD + r;  // C elements: row r, columns cJ + 2 t4, + 1
                    const double b0 = Ub[prt0], b1 = Ub[prt1];
                    double v0 = Cc[0], v1 = Cc[LD];
                    double e0 = 0.0, e1 = 0.0, w0v = 0.0, w1v = 0.0;
                    if (two) {
                        e0 = Ub[8 * NW * LD + prt0];
                        e1 = Ub[8 * NW * LD + prt1];
                        w0v = Cc[8 * NW * LD];
                        w1v = Cc[8 * NW * LD + LD];
                    }
                    dmma884(v0, v1, a0, b0);
                    if (two) dmma884(w0v, w1v, a0, e0);
                    dmma884(v0, v1, a1, b1);
                    if (two) dmma884(w0v, w1v, a1, e1);
                    Cc[0] = v0;
                    Cc[LD] = v1;
                    if (two) {
                        Cc[8 * NW * LD] = w0v;
                        Cc[8 * NW * LD + LD] = w1v;
                    }
                }
            }
        }
        bsync<NW>();
    }
    bsync<NW>();
    // ---- back substitution U x = y: the row with pivot position k is row perm[k], y there is b (padding rows: x = 0) --------
    if (w0) {
#pragma unroll 4
        for (int k = nr - 1; k >= 0; --k) {
            const int srow = perm[k];
            const double mine = (R == 2 && (srow & 32)) ? b[R - 1] * rdiag[R - 1] : b[0] * rdiag[0];
            const double xk = __shfl_sync(FULL, mine, srow & 31);
#pragma unroll
            for (int h = 0; h < R; ++h) {
                const double u = Hrow[k * LD + 32 * h];
                b[h] = (mypos[h] >= 0 && mypos[h] < k) ? fma(-u, xk, b[h]) : b[h];
            }
            if (lane == (k & 31)) dx[k] = xk;
        }
    }
    bsync<NW>();
}

}  // namespace rmx
