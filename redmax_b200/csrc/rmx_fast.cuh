// rmx_fast.cuh -- the composite ("v2") evaluation of the implicit-step residual and Newton matrix, and the warp-register LU.
//
// Same mathematics as rmx_device.cuh's eval_base / eval_columns (world-frame screws), reorganised so
// that nothing is a serial sweep over the joints any more:
//   * forward kinematics, V and U are prefix products / sums along root paths -> pointer-jumping scans, log2(depth) rounds
//     (ancestor tables anc[r][j] = 2^r-th ancestor of j come with the scene);
//   * the tangent wrench of body j for column i is linear in three per-column vectors,
//         T^i_j = B_j c2_i + A_j c1_i + sq C_j s_i ,     B_j = X' M X,  A_j = -c X'(M ad(phi) + dfcor/dphi + Dext) X,  C_j = -c X'(Kgrav + Kext) X
//     so with subtree-composite blocks B^C_k, A^C_k, C^C_k (one leaves->root accumulation, parallel over components)
//         H[k][i] = (B^C_k s_k).c2_i + (A^C_k' s_k).c1_i + sq (C^C_k' s_k).s_i          k in sub(i)
//                 = s_k.(B^C_i c2_i + A^C_i c1_i + sq C^C_i s_i + sq ad*(s_i) F^C_i)    k a proper ancestor of i
//     i.e. 12 (18 with ground contact) multiply-adds per matrix entry instead of a 330-flop body evaluation per (i,j) pair;
//   * without external forces A_j, B_j, C_j have closed forms needing 22 numbers per body (rotated inertia, m p, m, a 3x3
//     block and the linear momentum); ground contact adds two dense 6x6 blocks per body.
// Shared memory is structure-of-arrays (field-major, joint index fastest): every phase is bank-conflict free.
// The prototype tools/proto_composite.py checks these formulas against the dense oracle (1e-15 on g, H, M, D).
#pragma once
#include "rmx_device.cuh"

namespace rmx {

// Field offsets in the SoA block (units: NS doubles).  KEEP = 1: body frames and twists stay in shared memory and H has its
// own storage (test hooks, Krylov and external-force adjoint kernels); KEEP = 2: body frames only (tensor-core adjoint kernel);
// KEEP = 0: nothing kept, H aliases the fields that are dead once the Newton matrix is assembled.
// GROUND (everywhere in these files) is the level of external forces compiled in: 0 none, 1 ForceGroundCuboid, 2 ground contact
// and the forces between body points (springs, cables) -- scenes without those must not carry their code, stack frame and
// registers.
template <int GROUND, int KEEP>
struct Fld {
    static constexpr int NL = GROUND ? 18 : 12;  // length of L_k
    static constexpr int NW_ = NL + 6;           // W_k = [L_k ; s_k], stored joint-major (AoS) in region XA
    static constexpr int RW = 0;                 // 9  joint frame rotation (scan only; region XA)
    static constexpr int PW = 9;                 // 3
    static constexpr int XA = NW_;               // size of region XA in fields
    static constexpr int S = XA;                 // 6  world screw
    static constexpr int V = S + 6;              // 6
    static constexpr int U = V + 6;              // 6
    static constexpr int CF = U + 6;             // 6  F (composite after the accumulation)
    static constexpr int JB = CF + 6;            // 6  sum (R I3 R' - m [p][p]) : xx xy xz yy yz zz
    static constexpr int MP = JB + 6;            // 3  sum m p
    static constexpr int MS = MP + 3;            // 1  sum m
    static constexpr int ATL = MS + 1;           // 9  sum -c (R Ptl R' + 2 m [p][vc])
    static constexpr int MV = ATL + 9;           // 3  sum m vc
    static constexpr int AEXT = MV + 3;          // 36 sum -c X' Dext X   (GROUND only)
    static constexpr int CEXT = AEXT + 36;       // 36 sum -c X' Kext X   (GROUND only)
    static constexpr int NCOMP = GROUND ? 100 : 28;  // composite components starting at CF
    static constexpr int RB = CF + NCOMP;        // 9  body frame            (KEEP only)
    static constexpr int PB = RB + 9;            // 3
    static constexpr int PHI = PB + 3;           // 6  body twist            (KEEP == 1 only)
    static constexpr int TOTAL = CF + NCOMP + (KEEP == 1 ? 18 : (KEEP == 2 ? 12 : 0));
    static constexpr int HALIAS = V;             // H may live in [V, CF + NCOMP) when !KEEP
    static constexpr int HROOM = 12 + NCOMP;
};

// Joint stride of the SoA block: 33 for one-warp kernels (n, nr <= 32), 65 for two-warp kernels (n, nr <= 64), so that every
// field offset is a compile-time immediate (const int NS = NW == 1 ? 33 : 65): no address arithmetic per shared-memory access.
__host__ __device__ inline int soa_stride(int n, int nr) { return (n <= 32 && nr <= 32) ? 33 : 65; }
// leading dimension of H: the tensor-core forward path (!keep) fixes it at 33 / 65 for the same reason
__host__ __device__ inline int h_ld2(int n, int nr, bool keep) { return keep ? h_ld(nr) : soa_stride(n, nr); }

__host__ __device__ inline int fld_total(bool ground, bool keep) {
    return ground ? (keep ? Fld<true, 1>::TOTAL : Fld<true, 0>::TOTAL)
                  : (keep ? Fld<false, 1>::TOTAL : Fld<false, 0>::TOTAL);
}
// Tensor-core forward path (rmx_tc.cuh; one or two warps, n, nr <= 64, !KEEP): during the assembly the SoA block is overlaid by
//   W  [CAP][NW_] at 0 (rows [L_k ; s_k]),  RZ [CAP][NW_] behind it (rows [Rt_i ; Z_i]),  H [CAP][LD] behind both (column-major)
__host__ __device__ inline bool tc_layout(int n, int nr, bool keep) { return !keep && n <= 64 && nr <= 64; }
__host__ __device__ inline size_t soa_doubles(int n, int nr, bool ground, bool keep) {
    const size_t d = (size_t)soa_stride(n, nr) * fld_total(ground, keep);
    return (d + 1) & ~(size_t)1;
}

// Static shared-memory layout of the tensor-core forward kernels (capacity n = nr = 32 NW whatever the scene): every vector
// and table sits at a compile-time offset from the block's base, so no pointer lives in a register and no address is computed.
template <int GROUND, int NW>
struct TcLayout {
    typedef Fld<GROUND, 0> F;
    static constexpr int KEEP = 0;
    static constexpr int CAP = 32 * NW, NS = CAP + 1, LD = CAP + 1;
    static constexpr int W_OFF = 0;                   // W  [CAP][NW_]  rows [L_k ; s_k]
    static constexpr int RZ_OFF = CAP * F::NW_;       // RZ [32][NW_]   rows [Rt_i ; Z_i], 32 column joints at a time
    static constexpr int RZ_ROWS = 32;
    static constexpr int H_OFF = (CAP + RZ_ROWS) * F::NW_;  // H  [CAP][LD]   column-major, identity padded to whole tiles
    static constexpr int SOA_FIELDS = NS * F::TOTAL;  // the SoA block of eval_base2 (overlaid by W, RZ, H during assembly + LU)
    static constexpr int SOA = ((SOA_FIELDS > H_OFF + CAP * LD ? SOA_FIELDS : H_OFF + CAP * LD) + 1) & ~1;
    static constexpr int NV = 9;                      // q g (= dx) tau hq0 hqd0 hq1 hqd1 sp1 sp2
    static constexpr int VEC = SOA;                   // NV vectors of CAP
    static constexpr int RED = VEC + NV * CAP;        // 16 + 8 reduction scratch (cross-warp reductions: NW > 1 only)
    static constexpr int ROWBUF = RED + (NW == 1 ? 0 : 24);  // 2 x 5 double2 pivot-row buffers (16B aligned: all terms even)
    static constexpr int IE = ROWBUF + 20;            // int2 ie_s[CAP]
    static constexpr int PAR = IE + CAP;              // int par_s[CAP]
    static constexpr int REM = PAR + CAP / 2;         // int rem_s[CAP]
    static constexpr int TIDX = REM + CAP / 2;        // int tcidx_s[CAP]
    static constexpr int MASKD = (NW == 1) ? CAP / 2 : CAP;  // tree-relation masks: one bit per joint (32 / 64 bits)
    static constexpr int TSUB = TIDX + CAP / 2;       // mask tcsub_s[CAP]
    static constexpr int TANC = TSUB + MASKD;         // mask tcanc_s[CAP]
    static constexpr int TOTAL = (TANC + MASKD + 1) & ~1;
};
// Tensor-core adjoint forward kernel (one warp, no external forces): the composite blocks must survive the assembly, because
// the tape needs M and D (two more passes with other seeds) besides the Newton matrix.  W overlays region XA (the joint frames
// of the scans, dead after eval_base2), RZ has its own storage, the M and D tiles go straight to the tape in global memory and
// the H pass comes last, so H may overlay the fields from S on; the body frames (task Jacobian rows) sit behind it.
struct TcLayoutA {
    typedef Fld<false, 2> F;
    static constexpr int KEEP = 2;
    static constexpr int CAP = 32, NS = 33, LD = 33;
    static constexpr int W_OFF = 0;
    static constexpr int RZ_ROWS = 32;
    static constexpr int H_OFF = F::XA * NS;          // 594: behind W, over S, V, U, composite blocks (ends before RB)
    static constexpr int SOA = (NS * F::TOTAL + 1) & ~1;
    static constexpr int RZ_OFF = SOA;
    // 6 x CAP rows of J of the task body: over the kept body frames, which are dead from the moment those rows are formed
    // (adjoint_tape_md) to the next base evaluation -- the objective reads its row before that evaluation (rollout_fwd_kernel).
    // Without a region of their own the block is 28 160 B with the kernel's static part: eight blocks per SM instead of seven.
    static constexpr int JROWS = NS * F::RB;
    static constexpr int NV = 9;                      // q g (= dx) tau hq0 hqd0 hq1 hqd1 sp1 sp2
    static constexpr int VEC = RZ_OFF + CAP * F::NW_;
    static constexpr int RED = VEC + NV * CAP;
    static constexpr int ROWBUF = RED;                 // (one warp: no reduction scratch)
    static constexpr int IE = ROWBUF + 20;
    static constexpr int PAR = IE + CAP;
    static constexpr int REM = PAR + CAP / 2;
    static constexpr int TIDX = REM + CAP / 2;
    static constexpr int TSUB = TIDX + CAP / 2;
    static constexpr int TANC = TSUB + CAP / 2;
    static constexpr int TOTAL = (TANC + CAP / 2 + 1) & ~1;
    static_assert(CAP * F::NW_ <= F::XA * NS, "W must fit region XA");
    static_assert(H_OFF + CAP * LD <= NS * F::RB, "H must end before the kept body frames");
    static_assert(JROWS + 6 * CAP <= SOA, "J rows must fit the body-frame fields");
    static_assert((TOTAL * 8 + 128 + 1024) * 8 <= 228 * 1024, "eight blocks per SM");
};
// the adjoint forward kernel of a scene runs on the tensor-core path if ...
__host__ __device__ inline bool tc_adjoint(int n, int nr, bool ground) { return !ground && n <= 32 && nr <= 32; }
template <int NW> struct TcMask { typedef unsigned type; };
template <> struct TcMask<2> { typedef unsigned long long type; };
constexpr int LUBUF = 2 * (32 / 2 + 1) * 2;  // doubles: two pivot-row buffers of the warp LU (lu_solve_warp_sm), 16B aligned

__host__ __device__ inline size_t smem_doubles2(int n, int nr, bool ground, bool keep) {
    if (tc_layout(n, nr, keep)) {
        if (n <= 32 && nr <= 32) return ground ? TcLayout<true, 1>::TOTAL : TcLayout<false, 1>::TOTAL;
        return ground ? TcLayout<true, 2>::TOTAL : TcLayout<false, 2>::TOTAL;
    }
    size_t d = soa_doubles(n, nr, ground, keep) + LUBUF + (size_t)NVEC * nr + 16 + 8;
    d += (size_t)nr * h_ld2(n, nr, keep);   // H has its own storage
    d += (size_t)(3 * n + 1) / 2 + 1 + 16;  // int tables {idx,end}/parent, rem[32]
    return (d + 1) & ~(size_t)1;
}

struct Ctx2 : Ctx {
    double* sa;  // SoA block
    double2* lubuf;  // [LUBUF/2] pivot-row buffers of the warp LU
    int NS;      // stride (= n|1: odd, so that component-major accesses are conflict free too)
    int2* ie_s;  // [n] {reduced index or -1, subtree end}
    int* par_s;  // [n] parent
    int* rem_s;  // [CAP] rows still to be eliminated (blocked LU, rmx_tc.cuh)
    int* tcidx_s;         // [CAP] reduced index of joint i or -1            (tensor-core path only)
    void* tcsub_s;        // [CAP] TcMask<NW>: bit i set: joint k is in sub(i)
    void* tcanc_s;        // [CAP] bit i set: joint k is a proper ancestor of i
    double* jrows;        // [6][CAP] rows of J of the task body (adjoint kernels)
    double2* tcrow_s;     // [2][5] pivot-row buffers of the blocked LU
    const int* __restrict__ anc;  // [nrounds][n] ancestor tables (global)
    int nrounds;
    int anyext;    // external-force variants: some body of this rollout has contact / an attached force in the current evaluation
    int anc_r[6];  // this thread's 2^r-th ancestors (nrounds <= 6 for n <= 64), read once per kernel: the scans' rounds are
                   // a dependent chain and must not wait for a global load each
};

// The pointer-jumping scans of eval_base2 run their rounds as a real loop (one copy of the round's code, the ancestor picked
// from the register array by a select chain): unrolled six times they were 800 of the kernel's 6 800 instructions, and the
// one-warp kernels are sensitive to their executed code footprint (instruction fetch from the GPC-level cache runs at 90 % of
// its peak, profiles/r02_residency_ab.log).  -DRMX_SCAN_UNROLLED restores the unrolled rounds.
#ifdef RMX_SCAN_UNROLLED
#define RMX_SCAN_UNROLL _Pragma("unroll")
#else
#define RMX_SCAN_UNROLL _Pragma("unroll 1")
#endif
__device__ __forceinline__ int anc_round(const int (&a)[6], int r) {
#ifdef RMX_SCAN_UNROLLED
    return a[r];
#else
    int v = a[0];
    v = r == 1 ? a[1] : v;
    v = r == 2 ? a[2] : v;
    v = r == 3 ? a[3] : v;
    v = r == 4 ? a[4] : v;
    v = r == 5 ? a[5] : v;
    return v;
#endif
}

__device__ __forceinline__ void ctx2_carve(Ctx2& c, double* sm, int n, int nr, bool ground, bool keep) {
    c.n = n;
    c.nr = nr;
    c.ld = h_ld2(n, nr, keep);
    c.NS = soa_stride(n, nr);
    double* p = sm;
    c.sa = p;
    p += soa_doubles(n, nr, ground, keep);  // even number of doubles: p stays 16B aligned
    c.lubuf = reinterpret_cast<double2*>(p);
    p += LUBUF;
    c.rec1 = nullptr;
    c.rec2 = nullptr;
    c.KD = nullptr;
    c.q = p; p += nr;
    c.qd = p; p += nr;
    c.dq = p; p += nr;
    c.g = p; p += nr;
    c.dx = p; p += nr;
    c.x0 = p; p += nr;
    c.tau = p; p += nr;
    c.hq0 = p; p += nr;
    c.hqd0 = p; p += nr;
    c.hq1 = p; p += nr;
    c.hqd1 = p; p += nr;
    c.sp0 = p; p += nr;
    c.sp1 = p; p += nr;
    c.sp2 = p; p += nr;
    c.red = p;
    p += 16;
    c.ie_s = reinterpret_cast<int2*>(p);
    c.par_s = reinterpret_cast<int*>(p) + 2 * n;
    p += (size_t)(3 * n + 1) / 2 + 1;
    c.rem_s = reinterpret_cast<int*>(p);
    p += 16;
    c.tcidx_s = nullptr;
    c.tcsub_s = nullptr;
    c.tcanc_s = nullptr;
    c.tcrow_s = nullptr;
    c.jrows = nullptr;
    p = (double*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    c.H = p;
}

// Tensor-core forward kernels: the static layout above (nothing depends on n or nr).
template <int GROUND, int NW>
__device__ __forceinline__ void ctx2_carve_tc(Ctx2& c, double* sm, int n, int nr) {
    typedef TcLayout<GROUND, NW> T;
    c.n = n;
    c.nr = nr;
    c.ld = T::LD;
    c.NS = T::NS;
    c.sa = sm;
    c.lubuf = nullptr;
    c.rec1 = nullptr;
    c.rec2 = nullptr;
    c.KD = nullptr;
    double* v = sm + T::VEC;
    c.q = v;
    c.qd = nullptr;  // formed per joint in eval_base2
    c.dq = nullptr;
    c.g = v + 1 * T::CAP;
    c.dx = c.g;      // the solve reads its right-hand side into registers before it writes the solution
    c.x0 = nullptr;
    c.sp0 = nullptr;
    c.tau = v + 2 * T::CAP;
    c.hq0 = v + 3 * T::CAP;
    c.hqd0 = v + 4 * T::CAP;
    c.hq1 = v + 5 * T::CAP;
    c.hqd1 = v + 6 * T::CAP;
    c.sp1 = v + 7 * T::CAP;
    c.sp2 = v + 8 * T::CAP;
    c.red = sm + T::RED;
    c.tcrow_s = reinterpret_cast<double2*>(sm + T::ROWBUF);
    c.ie_s = reinterpret_cast<int2*>(sm + T::IE);
    c.par_s = reinterpret_cast<int*>(sm + T::PAR);
    c.rem_s = reinterpret_cast<int*>(sm + T::REM);
    c.tcidx_s = reinterpret_cast<int*>(sm + T::TIDX);
    c.tcsub_s = sm + T::TSUB;
    c.tcanc_s = sm + T::TANC;
    c.H = sm + T::H_OFF;
    c.jrows = nullptr;
}

// Tensor-core adjoint forward kernel: TcLayoutA
__device__ __forceinline__ void ctx2_carve_tca(Ctx2& c, double* sm, int n, int nr) {
    typedef TcLayoutA T;
    c.n = n;
    c.nr = nr;
    c.ld = T::LD;
    c.NS = T::NS;
    c.sa = sm;
    c.lubuf = nullptr;
    c.rec1 = nullptr;
    c.rec2 = nullptr;
    c.KD = nullptr;
    double* v = sm + T::VEC;
    c.q = v;
    c.qd = nullptr;
    c.dq = nullptr;
    c.g = v + 1 * T::CAP;
    c.dx = c.g;
    c.x0 = nullptr;
    c.sp0 = nullptr;
    c.tau = v + 2 * T::CAP;
    c.hq0 = v + 3 * T::CAP;
    c.hqd0 = v + 4 * T::CAP;
    c.hq1 = v + 5 * T::CAP;
    c.hqd1 = v + 6 * T::CAP;
    c.sp1 = v + 7 * T::CAP;
    c.sp2 = v + 8 * T::CAP;
    c.red = sm + T::RED;
    c.tcrow_s = reinterpret_cast<double2*>(sm + T::ROWBUF);
    c.ie_s = reinterpret_cast<int2*>(sm + T::IE);
    c.par_s = reinterpret_cast<int*>(sm + T::PAR);
    c.rem_s = reinterpret_cast<int*>(sm + T::REM);
    c.tcidx_s = reinterpret_cast<int*>(sm + T::TIDX);
    c.tcsub_s = sm + T::TSUB;
    c.tcanc_s = sm + T::TANC;
    c.H = sm + T::H_OFF;
    c.jrows = sm + T::JROWS;
}

#define SA(f, k, j) c.sa[(size_t)((f) + (k)) * NS + (j)]

// X' M X (scaled) for a dense body-frame 6x6 M (row-major), X = Ad(E^-1): column m of the result is
// xf_b2w(M * xm_w2b(e_m)).  Adds into the SoA field `fld` (row-major 6x6) of joint j.
__device__ __forceinline__ void xtmx_store(double* sa, int NS, int fld, int j, const double* R, const double* p, const double* M,
                                           double scale) {
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        double e[6] = {0, 0, 0, 0, 0, 0}, xe[6], y[6], wv[6];
        e[m] = 1.0;
        xm_w2b(R, p, e, xe);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double acc = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc += M[6 * r + k] * xe[k];
            y[r] = acc;
        }
        xf_b2w(R, p, y, wv);
#pragma unroll
        for (int r = 0; r < 6; ++r) sa[(size_t)(fld + 6 * r + m) * NS + j] = scale * wv[r];
    }
}

// same as xtmx_store, accumulating into the field
__device__ __forceinline__ void xtmx_add(double* sa, int NS, int fld, int j, const double* R, const double* p, const double* M,
                                         double scale) {
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        double e[6] = {0, 0, 0, 0, 0, 0}, xe[6], y[6], wv[6];
        e[m] = 1.0;
        xm_w2b(R, p, e, xe);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double acc = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc += M[6 * r + k] * xe[k];
            y[r] = acc;
        }
        xf_b2w(R, p, y, wv);
#pragma unroll
        for (int r = 0; r < 6; ++r) sa[(size_t)(fld + 6 * r + m) * NS + j] += scale * wv[r];
    }
}

// ---------------------------------------------------------------------------------------------
// ForcePointPoint (ForcePointPoint.m:48-113) in the composite formulation (prototype: tools/proto_pointforce.py).
// The force on "me" (a body point xl) from the other end is f = ks (xw_o - xw_me) + kd (vw_o - vw_me); its body-frame wrench and
// the diagonal blocks Km_aa, Dm_aa go where ground contact goes (fb, K, D); the off-diagonal blocks are kept per ordered pair as
// world-frame 6x6 matrices Aext_ab = -c X_a' Dm_ab X_b, Cext_ab = -c X_a' Km_ab X_b for the cross-term pass of the assembly.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pf_point(const double* R, const double* p, const double* phi, const double* xl, double* xw, double* vl,
                                         double* vw) {
    mat3_vec(R, xl, xw);
    xw[0] += p[0]; xw[1] += p[1]; xw[2] += p[2];
    cross3(phi, xl, vl);  // Gamma(xl) phi = w x xl + v
    vl[0] += phi[3]; vl[1] += phi[4]; vl[2] += phi[5];
    mat3_vec(R, vl, vw);
}

// dst (row-major 6x6) = scale * X_a' M X_o,  X = Ad(E^-1)
__device__ __forceinline__ void xtmy_store(double* dst, const double* Ra, const double* pa, const double* Ro, const double* po,
                                           const double* M, double scale) {
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        double e[6] = {0, 0, 0, 0, 0, 0}, xe[6], y[6], wv[6];
        e[m] = 1.0;
        xm_w2b(Ro, po, e, xe);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double acc = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc += M[6 * r + k] * xe[k];
            y[r] = acc;
        }
        xf_b2w(Ra, pa, y, wv);
#pragma unroll
        for (int r = 0; r < 6; ++r) dst[6 * r + m] = scale * wv[r];
    }
}

// rows 0..2 of the 6x6 block = [xl] M3x6, rows 3..5 = M3x6   (Gamma(xl)' M)
__device__ __forceinline__ void pf_gammaT(const double* xl, const double* M36, double* out, bool accumulate) {
#pragma unroll
    for (int cI = 0; cI < 6; ++cI) {
        const double m0 = M36[cI], m1 = M36[6 + cI], m2 = M36[12 + cI];
        const double t0 = xl[1] * m2 - xl[2] * m1, t1 = xl[2] * m0 - xl[0] * m2, t2 = xl[0] * m1 - xl[1] * m0;
        if (accumulate) {
            out[cI] += t0; out[6 + cI] += t1; out[12 + cI] += t2;
            out[18 + cI] += m0; out[24 + cI] += m1; out[30 + cI] += m2;
        } else {
            out[cI] = t0; out[6 + cI] = t1; out[12 + cI] = t2;
            out[18 + cI] = m0; out[24 + cI] = m1; out[30 + cI] = m2;
        }
    }
}

// ForceSpringGeneric.computeValues_ (ForceSpringGeneric.m:35-143) with ForceSpringDamper.computeSpringForce
// (ForceSpringDamper.m:65-72) for one end ("me" = side sd) of a spring: body-frame wrench f6, diagonal blocks Kown / Down and
// off-diagonal blocks Kab / Dab (rows: me, columns: the other body), all 6x6 row-major, written out as the reference forms
// its 12x12 K = K1 + K2 and D.  E[0], E[1] are the two ends in the reference's order (R p phi x); a world end has R = I, p = 0,
// phi = 0.
struct PfEnd {
    double R[9], p[3], phi[6], x[3];
};
static __device__ __noinline__ void pf_spring(const PfEnd* E, int sd, double ks, double kd, double L, double* f6, double* Kown, double* Down,
                                       double* Kab, double* Dab, bool deriv) {
    double xw[2][3], vl[2][3], vw[2][3];
#pragma unroll
    for (int e = 0; e < 2; ++e) pf_point(E[e].R, E[e].p, E[e].phi, E[e].x, xw[e], vl[e], vw[e]);
    double dx[3], dv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dx[i] = xw[1][i] - xw[0][i];
        dv[i] = vw[1][i] - vw[0][i];
    }
    const double l2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    const double l = sqrt(l2);
    const double ldot = (dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2]) / l;
    const double strain = (l - L) / L, dstrain = ldot / L;
    const double fs = ks * strain + kd * dstrain, dfsdl = ks / L, dfsdldot = kd / L;
    // fx = [G1'R1'dx ; -G2'R2'dx]
    double fx[12];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double y[3], t3[3];
        mat3T_vec(E[e].R, dx, y);
        cross3(E[e].x, y, t3);
        const double sg = e ? -1.0 : 1.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            fx[6 * e + i] = sg * t3[i];
            fx[6 * e + 3 + i] = sg * y[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) f6[i] = (fs / l) * fx[6 * sd + i];
    if (!deriv) return;
    // A = [-R1 G1, R2 G2] (3 x 12), G = [-[x], I]:  R G = [-R [x], R]
    double A[36];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double X[9], RX[9];
        brac3(E[e].x, X);
        mat3_mul(E[e].R, X, RX);
        const double sg = e ? 1.0 : -1.0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cI = 0; cI < 3; ++cI) {
                A[12 * r + 6 * e + cI] = -sg * RX[3 * r + cI];
                A[12 * r + 6 * e + 3 + cI] = sg * E[e].R[3 * r + cI];
            }
    }
    double dldq[12], dldotdq[12], w3[3];
    {
        // ((dx'dx I - dx dx') / l^3 dv)
        const double dd = dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2];
        const double il3 = 1.0 / (l2 * l);
#pragma unroll
        for (int i = 0; i < 3; ++i) w3[i] = (l2 * dv[i] - dx[i] * dd) * il3;
    }
#pragma unroll
    for (int cI = 0; cI < 12; ++cI) {
        dldq[cI] = (dx[0] * A[cI] + dx[1] * A[12 + cI] + dx[2] * A[24 + cI]) / l;
        dldotdq[cI] = w3[0] * A[cI] + w3[1] * A[12 + cI] + w3[2] * A[24 + cI];
    }
    // rotation columns: dldotdq(ax) += dx'/l * (-R1 [e_ax] G1 phi1),  dldotdq(6+ax) += dx'/l * (R2 [e_ax] G2 phi2);  G phi = vl
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double y[3];
        mat3T_vec(E[e].R, dx, y);  // R' dx
        // dx' R [e_ax] vl = (R'dx) . (e_ax x vl) = e_ax . (vl x R'dx)
        double t3[3];
        cross3(vl[e], y, t3);
        const double sg = e ? 1.0 : -1.0;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) dldotdq[6 * e + ax] += sg * t3[ax] / l;
    }
    double dfsdq[12];
#pragma unroll
    for (int cI = 0; cI < 12; ++cI) dfsdq[cI] = dfsdl * dldq[cI] + dfsdldot * dldotdq[cI];
    // K2 (12 x 12), ForceSpringGeneric.m:96-116: only the six rows of this end are kept
    double K2[72];
    for (int i = 0; i < 72; ++i) K2[i] = 0.0;
    {
        const double* R1 = E[0].R;
        const double* R2 = E[1].R;
        double x1b[9], x2b[9], R2R1[9], R1R2[9], B[9], T[9], d3[3], a3[3];
        brac3(E[0].x, x1b);
        brac3(E[1].x, x2b);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cI = 0; cI < 3; ++cI) {
                R2R1[3 * r + cI] = R2[r] * R1[cI] + R2[3 + r] * R1[3 + cI] + R2[6 + r] * R1[6 + cI];
            }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cI = 0; cI < 3; ++cI) R1R2[3 * r + cI] = R2R1[3 * cI + r];
        auto put = [&](int r0, int c0, const double* M, double sg) {
            if (r0 / 6 != sd) return;
            for (int r = 0; r < 3; ++r)
                for (int cI = 0; cI < 3; ++cI) K2[12 * (r0 - 6 * sd + r) + c0 + cI] = sg * M[3 * r + cI];
        };
        const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        // column block 1:3
        for (int i = 0; i < 3; ++i) d3[i] = E[0].p[i] - xw[1][i];
        mat3T_vec(R1, d3, a3);
        brac3(a3, B);             // K2(4:6,1:3) = [R1'(p1 - xw2)]
        put(3, 0, B, 1.0);
        mat3_mul(x1b, B, T);
        put(0, 0, T, 1.0);        // K2(1:3,1:3) = x1b * K2(4:6,1:3)
        mat3_mul(R2R1, x1b, B);
        put(9, 0, B, 1.0);        // K2(10:12,1:3) = R2R1 x1b
        mat3_mul(x2b, B, T);
        put(6, 0, T, 1.0);        // K2(7:9,1:3) = x2b * K2(10:12,1:3)
        // column block 4:6
        put(3, 3, I3, 1.0);
        put(0, 3, x1b, 1.0);
        put(9, 3, R2R1, -1.0);
        mat3_mul(x2b, R2R1, T);
        put(6, 3, T, -1.0);
        // column block 7:9
        mat3_mul(R1R2, x2b, B);
        put(3, 6, B, 1.0);
        mat3_mul(x1b, B, T);
        put(0, 6, T, 1.0);
        for (int i = 0; i < 3; ++i) d3[i] = E[1].p[i] - xw[0][i];
        mat3T_vec(R2, d3, a3);
        brac3(a3, B);
        put(9, 6, B, 1.0);
        mat3_mul(x2b, B, T);
        put(6, 6, T, 1.0);
        // column block 10:12
        put(3, 9, R1R2, -1.0);
        mat3_mul(x1b, R1R2, T);
        put(0, 9, T, -1.0);
        put(9, 9, I3, 1.0);
        put(6, 9, x2b, 1.0);
    }
    // D = -fx * [d_w'R1 G1, -d_w'R2 G2],  d_w = dfsdldot dx / l^2  ->  row vector dvec = -(d_w' A) (A = [-R1G1, R2G2])
    double dvec[12];
    {
        const double sdw = dfsdldot / l2;
#pragma unroll
        for (int cI = 0; cI < 12; ++cI) dvec[cI] = -sdw * (dx[0] * A[cI] + dx[1] * A[12 + cI] + dx[2] * A[24 + cI]);
    }
    const int r0 = 6 * sd, co = 6 * sd, cx = 6 * (1 - sd);
    const double fl = fs / l;
    for (int r = 0; r < 6; ++r)
        for (int cI = 0; cI < 6; ++cI) {
            const double fr = fx[r0 + r];
            Kown[6 * r + cI] = fr * (dfsdq[co + cI] / l - fs / l2 * dldq[co + cI]) - fl * K2[12 * r + co + cI];
            Kab[6 * r + cI] = fr * (dfsdq[cx + cI] / l - fs / l2 * dldq[cx + cI]) - fl * K2[12 * r + cx + cI];
            Down[6 * r + cI] = -fr * dvec[co + cI];
            Dab[6 * r + cI] = -fr * dvec[cx + cI];
        }
}

// What the out-of-line force routines need of the kernel's context, handed BY VALUE together with the body frame: passing
// `Ctx&` (or pointers to the caller's Rb / pb / phi) to a function that is not inlined makes those objects escape, the compiler
// then keeps the whole context in local memory and addresses shared memory generically everywhere in the kernel -- measured on
// the ground-contact variant as 1.3 GB of DRAM writes per launch and a 4.1-cycle long-scoreboard stall per issue.
struct PfCtx {
    const PointForce* pf;
    const int* pf_ep;
    double* pf_s;
    double cK;
};
struct BodyFrame {
    double R[9], p[3], phi[6];
};
struct Wrench6 {
    double f[6];
};

// One segment (point a -> point b) of a multi-point spring, ForceSpringMultiPointGeneric.m:58-81 and :103-170: length, length
// rate, the normalised generalised force fxn = [G1'R1'dx ; -G2'R2'dx] / |dx|, the gradients dldq, dldotdq (1 x 12), the
// damping row dqd = [-dxnor'R1G1, dxnor'R2G2], and -- for the end `side` (0: a, 1: b, < 0: none) -- the six rows of the
// normalised-vector stiffness K1 + K2/|dx| (6 x 12 row-major in Knr).
static __device__ __noinline__ void pf_segment(const PfEnd& Ea, const PfEnd& Eb, int side, bool deriv, double* dxlen_out, double* ldot_out,
                                        double* fxn, double* dldq, double* dldotdq, double* dqd, double* Knr) {
    const PfEnd* E[2] = {&Ea, &Eb};
    double xw[2][3], vl[2][3], vw[2][3];
#pragma unroll
    for (int e = 0; e < 2; ++e) pf_point(E[e]->R, E[e]->p, E[e]->phi, E[e]->x, xw[e], vl[e], vw[e]);
    double dx[3], dv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dx[i] = xw[1][i] - xw[0][i];
        dv[i] = vw[1][i] - vw[0][i];
    }
    const double l2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
    const double l = sqrt(l2);
    *dxlen_out = l;
    *ldot_out = (dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2]) / l;
    double fx[12];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double y[3], t3[3];
        mat3T_vec(E[e]->R, dx, y);
        cross3(E[e]->x, y, t3);
        const double sg = e ? -1.0 : 1.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            fx[6 * e + i] = sg * t3[i];
            fx[6 * e + 3 + i] = sg * y[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) fxn[i] = fx[i] / l;
    if (!deriv) return;
    double A[36];  // [-R1 G1, R2 G2]
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double X[9], RX[9];
        brac3(E[e]->x, X);
        mat3_mul(E[e]->R, X, RX);
        const double sg = e ? 1.0 : -1.0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cI = 0; cI < 3; ++cI) {
                A[12 * r + 6 * e + cI] = -sg * RX[3 * r + cI];
                A[12 * r + 6 * e + 3 + cI] = sg * E[e]->R[3 * r + cI];
            }
    }
    double w3[3];
    {
        const double dd = dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2];
        const double il3 = 1.0 / (l2 * l);
#pragma unroll
        for (int i = 0; i < 3; ++i) w3[i] = (l2 * dv[i] - dx[i] * dd) * il3;  // (I - n n')/|dx| dv
    }
    double dxA[12];  // dx' A
#pragma unroll
    for (int cI = 0; cI < 12; ++cI) {
        dxA[cI] = dx[0] * A[cI] + dx[1] * A[12 + cI] + dx[2] * A[24 + cI];
        dldq[cI] = dxA[cI] / l;
        dldotdq[cI] = w3[0] * A[cI] + w3[1] * A[12 + cI] + w3[2] * A[24 + cI];
        dqd[cI] = dxA[cI] / l;  // [-n'R1G1, n'R2G2] = n'A
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double y[3], t3[3];
        mat3T_vec(E[e]->R, dx, y);
        cross3(vl[e], y, t3);
        const double sg = e ? 1.0 : -1.0;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) dldotdq[6 * e + ax] += sg * t3[ax] / l;
    }
    if (side < 0) return;
    // K1 = fx [d'R1G1, -d'R2G2], d = -dx/|dx|^3  ->  K1 = fx (dx'A) / |dx|^3
    const double il3 = 1.0 / (l2 * l);
    for (int i = 0; i < 72; ++i) Knr[i] = 0.0;
    {
        const double* R1 = Ea.R;
        const double* R2 = Eb.R;
        double x1b[9], x2b[9], R2R1[9], R1R2[9], B[9], T[9], d3[3], a3[3];
        brac3(Ea.x, x1b);
        brac3(Eb.x, x2b);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cI = 0; cI < 3; ++cI) R2R1[3 * r + cI] = R2[r] * R1[cI] + R2[3 + r] * R1[3 + cI] + R2[6 + r] * R1[6 + cI];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cI = 0; cI < 3; ++cI) R1R2[3 * r + cI] = R2R1[3 * cI + r];
        auto put = [&](int r0, int c0, const double* M, double sg) {
            if (r0 / 6 != side) return;
            for (int r = 0; r < 3; ++r)
                for (int cI = 0; cI < 3; ++cI) Knr[12 * (r0 - 6 * side + r) + c0 + cI] = sg * M[3 * r + cI] / l;
        };
        const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int i = 0; i < 3; ++i) d3[i] = Ea.p[i] - xw[1][i];
        mat3T_vec(R1, d3, a3);
        brac3(a3, B);
        put(3, 0, B, 1.0);
        mat3_mul(x1b, B, T);
        put(0, 0, T, 1.0);
        mat3_mul(R2R1, x1b, B);
        put(9, 0, B, 1.0);
        mat3_mul(x2b, B, T);
        put(6, 0, T, 1.0);
        put(3, 3, I3, 1.0);
        put(0, 3, x1b, 1.0);
        put(9, 3, R2R1, -1.0);
        mat3_mul(x2b, R2R1, T);
        put(6, 3, T, -1.0);
        mat3_mul(R1R2, x2b, B);
        put(3, 6, B, 1.0);
        mat3_mul(x1b, B, T);
        put(0, 6, T, 1.0);
        for (int i = 0; i < 3; ++i) d3[i] = Eb.p[i] - xw[0][i];
        mat3T_vec(R2, d3, a3);
        brac3(a3, B);
        put(9, 6, B, 1.0);
        mat3_mul(x2b, B, T);
        put(6, 6, T, 1.0);
        put(3, 9, R1R2, -1.0);
        mat3_mul(x1b, R1R2, T);
        put(0, 9, T, -1.0);
        put(9, 9, I3, 1.0);
        put(6, 9, x2b, 1.0);
    }
    for (int r = 0; r < 6; ++r)
        for (int cI = 0; cI < 12; ++cI) Knr[12 * r + cI] += fx[6 * side + r] * dxA[cI] * il3;
}

// ForceCable through attachment `me` of force P (ForceSpringMultiPointGeneric.m:29-190, ForceCable.m:66-81): adds this body's
// wrench to fb and, if K != null, its diagonal blocks to K, D and writes its cross blocks K(me,k2), D(me,k2) (world frame) to
// shared memory.  K = fn dfsdq - fs Kn, D = fn dfsdqdot.
static __device__ __noinline__ void pf_cable(PfCtx c, const PointForce& P, int me, const double* Rb, const double* pb, double* fb, double* K,
                                      double* D) {
    const int np = P.npts;
    const double* recs = c.pf_s + P.rec_off;
    PfEnd E[PF_MAXPTS];
    for (int k = 0; k < np; ++k) {
        const bool world = P.body[k] < 0;
        const double* r = recs + (size_t)k * PF_REC;
        for (int i = 0; i < 9; ++i) E[k].R[i] = world ? ((i % 4 == 0) ? 1.0 : 0.0) : r[i];
        for (int i = 0; i < 3; ++i) {
            E[k].p[i] = world ? 0.0 : r[9 + i];
            E[k].x[i] = P.x[k][i];
        }
        for (int i = 0; i < 6; ++i) E[k].phi[i] = world ? 0.0 : r[12 + i];
    }
    const bool deriv = K != nullptr;
    double fn[6] = {0, 0, 0, 0, 0, 0};
    double dfl[6 * PF_MAXPTS], dfld[6 * PF_MAXPTS], dfd[6 * PF_MAXPTS];  // sums of dldq, dldotdq, dqd over the segments
    double Knrow[6 * 6 * PF_MAXPTS];                                       // rows of Kn of this attachment
    for (int i = 0; i < 6 * PF_MAXPTS; ++i) dfl[i] = dfld[i] = dfd[i] = 0.0;
    for (int i = 0; i < 36 * PF_MAXPTS; ++i) Knrow[i] = 0.0;
    double l = 0.0, ldot = 0.0;
    for (int sI = 0; sI + 1 < np; ++sI) {
        const int side = (me == sI) ? 0 : ((me == sI + 1) ? 1 : -1);
        double dxlen, ld1, fxn[12], dldq[12], dldotdq[12], dqd[12], Knr[72];
        pf_segment(E[sI], E[sI + 1], side, deriv, &dxlen, &ld1, fxn, dldq, dldotdq, dqd, Knr);
        l += dxlen;
        ldot += ld1;
        if (side >= 0)
            for (int i = 0; i < 6; ++i) fn[i] += fxn[6 * side + i];
        if (deriv) {
            for (int i = 0; i < 12; ++i) {
                dfl[6 * sI + i] += dldq[i];
                dfld[6 * sI + i] += dldotdq[i];
                dfd[6 * sI + i] += dqd[i];
            }
            if (side >= 0)
                for (int r = 0; r < 6; ++r)
                    for (int i = 0; i < 12; ++i) Knrow[(6 * PF_MAXPTS) * r + 6 * sI + i] += Knr[12 * r + i];
        }
    }
    // ForceCable.computeSpringForce: pulls only when stretched
    const double strain = (l - P.L) / P.L, dstrain = ldot / P.L;
    double fs = 0.0, dfsdl = 0.0, dfsdldot = 0.0;
    if (strain > 0) {
        fs = P.ks * strain + P.kd * dstrain;
        dfsdl = P.ks / P.L;
        dfsdldot = P.kd / P.L;
    }
    for (int i = 0; i < 6; ++i) fb[i] += fs * fn[i];
    if (!deriv) return;
    double* blks = c.pf_s + P.blk_off;
    for (int k2 = 0; k2 < np; ++k2) {
        double Kb[36], Db[36];
        for (int r = 0; r < 6; ++r)
            for (int i = 0; i < 6; ++i) {
                const int col = 6 * k2 + i;
                Kb[6 * r + i] = fn[r] * (dfsdl * dfl[col] + dfsdldot * dfld[col]) - fs * Knrow[(6 * PF_MAXPTS) * r + col];
                Db[6 * r + i] = fn[r] * (dfsdldot * dfd[col]);
            }
        if (k2 == me) {
            for (int i = 0; i < 36; ++i) {
                K[i] += Kb[i];
                D[i] += Db[i];
            }
        } else if (P.body[k2] >= 0) {
            double* blk = blks + (size_t)(me * np + k2) * PF_BLK;
            xtmy_store(blk, Rb, pb, E[k2].R, E[k2].p, Db, -c.cK);
            xtmy_store(blk + 36, Rb, pb, E[k2].R, E[k2].p, Kb, -c.cK);
        }
    }
}

// all point forces attached to body t: returns their body-frame wrench in f6 and, if deriv, adds the diagonal blocks
// -c X'DX, -c X'KX to the external-force fields aext / cext of joint t in the SoA block (after the caller has stored or zeroed
// them) and writes the cross blocks of its ordered pairs to shared memory.  Kept out of line and off the caller's registers:
// scenes without point forces must not pay for it.
static __device__ __noinline__ Wrench6 pf_body(PfCtx c, int pf_ptr, int pf_cnt, BodyFrame B, bool deriv, double* sa, int NS, int aext,
                                        int cext, int t) {
    const double* Rb = B.R;
    const double* pb = B.p;
    const double* phi = B.phi;
    double fb[6] = {0, 0, 0, 0, 0, 0};
    double Kacc[36], Dacc[36];
    for (int i = 0; i < 36; ++i) Kacc[i] = Dacc[i] = 0.0;
    double* K = deriv ? Kacc : nullptr;
    double* D = deriv ? Dacc : nullptr;
    for (int e = 0; e < pf_cnt; ++e) {
        const int code = __ldg(c.pf_ep + pf_ptr + e);
        const int f = code / PF_MAXPTS, sd = code % PF_MAXPTS;
        const PointForce& P = c.pf[f];
        const double* recs = c.pf_s + P.rec_off;  // attachment k at recs + k PF_REC
        double* blks = c.pf_s + P.blk_off;        // ordered pair (k, k2) at blks + (k npts + k2) PF_BLK
        if (P.kind == 2) {  // ForceCable
            pf_cable(c, P, sd, Rb, pb, fb, K, D);
            continue;
        }
        const int ob = P.body[1 - sd];
        const double ks = P.ks, kd = P.kd;
        double xl[3] = {P.x[sd][0], P.x[sd][1], P.x[sd][2]}, xo[3] = {P.x[1 - sd][0], P.x[1 - sd][1], P.x[1 - sd][2]};
        double xw[3], vl[3], vw[3], xwo[3], vlo[3] = {0, 0, 0}, vwo[3] = {0, 0, 0};
        pf_point(Rb, pb, phi, xl, xw, vl, vw);
        double Ro[9], po[3];
        if (ob >= 0) {
            const double* r = recs + (size_t)(1 - sd) * PF_REC;
#pragma unroll
            for (int i = 0; i < 9; ++i) Ro[i] = r[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) po[i] = r[9 + i];
            pf_point(Ro, po, r + 12, xo, xwo, vlo, vwo);
        } else {
            xwo[0] = xo[0]; xwo[1] = xo[1]; xwo[2] = xo[2];
        }
        if (P.kind == 1) {  // ForceSpringDamper
            PfEnd E[2];
            PfEnd& me = E[sd];
            PfEnd& ot = E[1 - sd];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                me.R[i] = Rb[i];
                ot.R[i] = (ob >= 0) ? Ro[i] : ((i % 4 == 0) ? 1.0 : 0.0);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                me.p[i] = pb[i];
                ot.p[i] = (ob >= 0) ? po[i] : 0.0;
                me.x[i] = xl[i];
                ot.x[i] = xo[i];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                me.phi[i] = phi[i];
                ot.phi[i] = (ob >= 0) ? recs[(size_t)(1 - sd) * PF_REC + 12 + i] : 0.0;
            }
            double f6[6], Ko[36], Do[36], Kab[36], Dab[36];
            pf_spring(E, sd, ks, kd, P.L, f6, Ko, Do, Kab, Dab, K != nullptr);
#pragma unroll
            for (int i = 0; i < 6; ++i) fb[i] += f6[i];
            if (K != nullptr) {
                for (int i = 0; i < 36; ++i) {
                    K[i] += Ko[i];
                    D[i] += Do[i];
                }
                if (ob >= 0) {
                    double* blk = blks + (size_t)(sd * 2 + (1 - sd)) * PF_BLK;
                    xtmy_store(blk, Rb, pb, ot.R, ot.p, Dab, -c.cK);
                    xtmy_store(blk + 36, Rb, pb, ot.R, ot.p, Kab, -c.cK);
                }
            }
            continue;
        }
        double fme[3], y[3], t3[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) fme[i] = ks * (xwo[i] - xw[i]) + kd * (vwo[i] - vw[i]);
        mat3T_vec(Rb, fme, y);
        cross3(xl, y, t3);
        fb[0] += t3[0]; fb[1] += t3[1]; fb[2] += t3[2];
        fb[3] += y[0]; fb[4] += y[1]; fb[5] += y[2];
        if (K == nullptr) continue;
        // diagonal blocks: Km_aa = G'[ks [R'(xw_o - p)] + kd [R' vw_o], -ks I],  Dm_aa = -kd G'G,  G = Gamma(xl) = [-[xl], I]
        {
            double d3[3] = {xwo[0] - pb[0], xwo[1] - pb[1], xwo[2] - pb[2]}, a3[3], b3[3], A[9], B[9], M36[18];
            mat3T_vec(Rb, d3, a3);
            mat3T_vec(Rb, vwo, b3);
            brac3(a3, A);
            brac3(b3, B);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cI = 0; cI < 3; ++cI) {
                    M36[6 * r + cI] = ks * A[3 * r + cI] + kd * B[3 * r + cI];
                    M36[6 * r + 3 + cI] = (r == cI) ? -ks : 0.0;
                }
            pf_gammaT(xl, M36, K, true);
            double X[9], G36[18];
            brac3(xl, X);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cI = 0; cI < 3; ++cI) {
                    G36[6 * r + cI] = kd * X[3 * r + cI];            // -kd * (-[xl])
                    G36[6 * r + 3 + cI] = (r == cI) ? -kd : 0.0;
                }
            pf_gammaT(xl, G36, D, true);
        }
        // off-diagonal blocks (both ends on bodies): Km_ab = G_a' R_a'R_o (ks [-[xo], I] - kd [[vl_o], 0]),  Dm_ab = kd G_a' R_a'R_o G_o
        if (ob >= 0) {
            double Rao[9], XO[9], VO[9], Kin[18], Din[18], Kab[36], Dab[36];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cI = 0; cI < 3; ++cI) Rao[3 * r + cI] = Rb[r] * Ro[cI] + Rb[3 + r] * Ro[3 + cI] + Rb[6 + r] * Ro[6 + cI];
            brac3(xo, XO);
            brac3(vlo, VO);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cI = 0; cI < 3; ++cI) {
                    double kl = 0.0, dl = 0.0;
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        kl += Rao[3 * r + m] * (-ks * XO[3 * m + cI] - kd * VO[3 * m + cI]);
                        dl += Rao[3 * r + m] * (-kd * XO[3 * m + cI]);
                    }
                    Kin[6 * r + cI] = kl;
                    Kin[6 * r + 3 + cI] = ks * Rao[3 * r + cI];
                    Din[6 * r + cI] = dl;
                    Din[6 * r + 3 + cI] = kd * Rao[3 * r + cI];
                }
            pf_gammaT(xl, Kin, Kab, false);
            pf_gammaT(xl, Din, Dab, false);
            double* blk = blks + (size_t)(sd * 2 + (1 - sd)) * PF_BLK;
            xtmy_store(blk, Rb, pb, Ro, po, Dab, -c.cK);
            xtmy_store(blk + 36, Rb, pb, Ro, po, Kab, -c.cK);
        }
    }
    if (deriv) {
        xtmx_add(sa, NS, aext, t, Rb, pb, Dacc, -c.cK);
        xtmx_add(sa, NS, cext, t, Rb, pb, Kacc, -c.cK);
    }
    Wrench6 w;
#pragma unroll
    for (int i = 0; i < 6; ++i) w.f[i] = fb[i];
    return w;
}

// Cross-term pass of the assembly: for every point force between two bodies and both orderings (a, b):
//   out[k][i] += scale * s_k . (Aext_ab c1_i + Cext_ab (sq s_i))     for joints k in anc*(a), i in anc*(b).
// Thread t owns column joint t; Wb: joint-major rows [L_k ; s_k] (stride NWD, s at offset NL).
__device__ __forceinline__ void pf_cross_pass(Ctx2& c, int t, int myidx, const double* c1, const double* sqs, double scale, double* out,
                                              int ld, const double* Wb, int NWD, int NL) {
    const int myend = (myidx >= 0) ? c.ie_s[t].y : 0;
    for (int f = 0; f < c.npf; ++f) {
        const PointForce& P = c.pf[f];
        const int np = P.npts;
        for (int pr = 0; pr < np * np; ++pr) {  // ordered pairs (k1, k2) of attachments on two bodies
            const int k1 = pr / np, k2 = pr - k1 * np;
            const int a = P.body[k1], b = P.body[k2];
            if (k1 == k2 || a < 0 || b < 0) continue;  // uniform
            const bool mine = myidx >= 0 && t <= b && b < myend;
            double y[6] = {0, 0, 0, 0, 0, 0};
            if (mine) {
                const double* A = c.pf_s + P.blk_off + (size_t)pr * PF_BLK;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < 6; ++m) acc += A[6 * r + m] * c1[m] + A[36 + 6 * r + m] * sqs[m];
                    y[r] = acc;
                }
            }
            for (int k = a; k >= 0; k = c.par_s[k]) {  // uniform walk up the ancestors of a
                const int ik = c.ie_s[k].x;
                if (ik < 0 || !mine) continue;
                const double* sk = Wb + (size_t)k * NWD + NL;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) acc += sk[m] * y[m];
                out[(size_t)myidx * ld + ik] += scale * acc;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// eval_base2: residual g at iterate c.q (and, if deriv, the composite blocks eval_columns2 needs).
// ---------------------------------------------------------------------------------------------
template <int NW, int GROUND, int KEEP>
__device__ void eval_base2(Ctx2& c, bool deriv) {
    typedef Fld<GROUND, KEEP> F;
    const int t = threadIdx.x;
    const int n = c.n, NS = (NW == 1) ? 33 : 65;
    const int NT = 32 * NW;
    // ---- joint-local transforms (the stage kinematics qdot, dqtmp of a dof are formed by its joint's thread below: a few
    //      exactly rounded operations, no vectors in shared memory) ------------------------------------------------------
    double Rj[9], pj[3];  // this joint's (partial) world frame, kept in registers through the scan
    int myidx = -1;
    if (t < n) {
        const JointConst& J = c.jc[t];
        myidx = J.idx;
        if (myidx >= 0 && !J.prismatic) {
            double Rq[9];
            aa_to_mat(J, c.q[myidx], Rq);
            mat3_mul(J.R0, Rq, Rj);
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) Rj[i] = J.R0[i];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) pj[i] = J.p0[i];
        if (myidx >= 0 && J.prismatic) {  // E_pj = E0_pj * trans(axis q)   (JointPrismatic.m:29-32)
            double aq[3] = {J.axis[0] * c.q[myidx], J.axis[1] * c.q[myidx], J.axis[2] * c.q[myidx]}, t3[3];
            mat3_vec(J.R0, aq, t3);
            pj[0] += t3[0]; pj[1] += t3[1]; pj[2] += t3[2];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) SA(F::RW, i, t) = Rj[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) SA(F::PW, i, t) = pj[i];
    }
    bsync<NW>();
    // ---- world frames: pointer-jumping prefix product  E_w,j = E_w,anc o E_(anc, j] -----------------------------
    RMX_SCAN_UNROLL
    for (int r = 0; r < 6; ++r) {
        if (r >= c.nrounds) break;  // uniform
        const int a = anc_round(c.anc_r, r);
        if (a >= 0) {
            double Ra[9], pa[3], Rn[9], pn[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) Ra[i] = SA(F::RW, i, a);
#pragma unroll
            for (int i = 0; i < 3; ++i) pa[i] = SA(F::PW, i, a);
            mat3_mul(Ra, Rj, Rn);
            mat3_vec(Ra, pj, pn);
#pragma unroll
            for (int i = 0; i < 9; ++i) Rj[i] = Rn[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) pj[i] = pn[i] + pa[i];
        }
        bsync<NW>();
        if (a >= 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) SA(F::RW, i, t) = Rj[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) SA(F::PW, i, t) = pj[i];
        }
        bsync<NW>();
    }
    // ---- screws, V = prefix sum of s qd ------------------------------------------------------------------------
    double s[6] = {0, 0, 0, 0, 0, 0}, Vj[6] = {0, 0, 0, 0, 0, 0}, Uj[6] = {0, 0, 0, 0, 0, 0};
    double qdk = 0.0, dqk = 0.0;
    if (t < n) {
        if (myidx >= 0) {
            const JointConst& J = c.jc[t];
            if (J.prismatic) {  // S = [0; a]  ->  world screw [0; R a]
                mat3_vec(Rj, J.axis, s + 3);
            } else {            // S = [a; 0]  ->  [R a; p x R a]
                mat3_vec(Rj, J.axis, s);
                cross3(pj, s, s + 3);
            }
            stage_kin(c.stage, c.h, c.q[myidx], c.hq0[myidx], c.hqd0[myidx], c.hq1[myidx], c.hqd1[myidx], qdk, dqk);
#pragma unroll
            for (int i = 0; i < 6; ++i) Vj[i] = s[i] * qdk;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            SA(F::S, i, t) = s[i];
            SA(F::V, i, t) = Vj[i];
        }
    }
    bsync<NW>();
    RMX_SCAN_UNROLL
    for (int r = 0; r < 6; ++r) {
        if (r >= c.nrounds) break;  // uniform
        const int a = anc_round(c.anc_r, r);
        if (a >= 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Vj[i] += SA(F::V, i, a);
        }
        bsync<NW>();
        if (a >= 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) SA(F::V, i, t) = Vj[i];
        }
        bsync<NW>();
    }
    // ---- U = prefix sum of (s dq + c qd ad(Vp) s) --------------------------------------------------------------
    if (t < n) {
        if (myidx >= 0) {
            const int par = c.par_s[t];
            double Vp[6] = {0, 0, 0, 0, 0, 0}, sd[6];
            if (par >= 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i) Vp[i] = SA(F::V, i, par);
            }
            ad_mv(Vp, s, sd);
            const double cq = c.c * qdk;
#pragma unroll
            for (int i = 0; i < 6; ++i) Uj[i] = s[i] * dqk + cq * sd[i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) SA(F::U, i, t) = Uj[i];
    }
    bsync<NW>();
    RMX_SCAN_UNROLL
    for (int r = 0; r < 6; ++r) {
        if (r >= c.nrounds) break;  // uniform
        const int a = anc_round(c.anc_r, r);
        if (a >= 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Uj[i] += SA(F::U, i, a);
        }
        bsync<NW>();
        if (a >= 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) SA(F::U, i, t) = Uj[i];
        }
        bsync<NW>();
    }
    // ---- point forces: every attached body publishes its frame and twist for the other end -------------------------
    if (GROUND == 2 && c.npf > 0) {  // uniform
        if (t < n && c.jc[t].pf_cnt > 0) {
            const JointConst& J = c.jc[t];
            double Rb[9], pb[3], phi[6];
            mat3_mul(Rj, J.Rji, Rb);
            mat3_vec(Rj, J.pji, pb);
            pb[0] += pj[0]; pb[1] += pj[1]; pb[2] += pj[2];
            xm_w2b(Rb, pb, Vj, phi);
            for (int e = 0; e < J.pf_cnt; ++e) {
                const int code = __ldg(c.pf_ep + J.pf_ptr + e);
                double* r = c.pf_s + c.pf[code / PF_MAXPTS].rec_off + (code % PF_MAXPTS) * PF_REC;
#pragma unroll
                for (int i = 0; i < 9; ++i) r[i] = Rb[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) r[9 + i] = pb[i];
#pragma unroll
                for (int i = 0; i < 6; ++i) r[12 + i] = phi[i];
            }
        }
        bsync<NW>();
    }
    // ---- per body: frame, twist, wrench, and (deriv) the world-frame blocks ------------------------------------
    bool myext = false;  // this body has ground contact or an attached force in this evaluation
    if (t < n) {
        const JointConst& J = c.jc[t];
        double Rb[9], pb[3];
        mat3_mul(Rj, J.Rji, Rb);
        mat3_vec(Rj, J.pji, pb);
        pb[0] += pj[0]; pb[1] += pj[1]; pb[2] += pj[2];
        double phi[6], u[6];
        xm_w2b(Rb, pb, Vj, phi);
        xm_w2b(Rb, pb, Uj, u);
        const double m = J.I[3];
        double Iw[3] = {J.I[0] * phi[0], J.I[1] * phi[1], J.I[2] * phi[2]};
        double mv[3] = {m * phi[3], m * phi[4], m * phi[5]};
        double fb[6];  // fcor + fgrav + fext (Body.m:98-107)
        cross3(Iw, phi, fb);
        cross3(mv, phi, fb + 3);
        double gw[3] = {c.gx, c.gy, c.gz}, gb[3];
        mat3T_vec(Rb, gw, gb);
        fb[3] += m * gb[0]; fb[4] += m * gb[1]; fb[5] += m * gb[2];
        if (GROUND) {
            // a cuboid whose lowest corner is clearly above the plane takes no part (ForceGroundCuboid.m:89-93 skips every
            // corner with d > 0): lowest corner depth = n.(p - xg) - sum_i |n_b,i| hs_i
            bool contact = false;
            if (J.has_ground) {
                double nb[3];
                mat3T_vec(Rb, J.gng, nb);
                const double dp = J.gng[0] * (pb[0] - J.gxg[0]) + J.gng[1] * (pb[1] - J.gxg[1]) + J.gng[2] * (pb[2] - J.gxg[2]);
                const double reach = fabs(nb[0]) * J.hs[0] + fabs(nb[1]) * J.hs[1] + fabs(nb[2]) * J.hs[2];
                contact = dp - reach <= 1e-12 * (fabs(dp) + reach);
            }
            myext = contact || (GROUND == 2 && c.npf > 0 && J.pf_cnt > 0);
            if (contact) {
                if (deriv) {
                    double K[36], D[36];
#pragma unroll
                    for (int i = 0; i < 36; ++i) K[i] = D[i] = 0;
                    ground_body<true>(J, Rb, pb, phi, fb, K, D);
#ifdef RMX_XTMX_TWO_COPIES
                    xtmx_store(c.sa, NS, F::AEXT, t, Rb, pb, D, -c.c);
                    xtmx_store(c.sa, NS, F::CEXT, t, Rb, pb, K, -c.c);
#else
                    // one copy of the congruence code for both blocks (code footprint: see anc_round above)
#pragma unroll 1
                    for (int pass = 0; pass < 2; ++pass)
                        xtmx_store(c.sa, NS, pass ? F::CEXT : F::AEXT, t, Rb, pb, pass ? K : D, -c.c);
#endif
                } else {
                    ground_body<false>(J, Rb, pb, phi, fb, nullptr, nullptr);
                }
            } else if (deriv) {
                for (int i = 0; i < 72; ++i) SA(F::AEXT, i, t) = 0.0;
            }
            if (GROUND == 2 && c.npf > 0 && J.pf_cnt > 0) {
                PfCtx pc;
                pc.pf = c.pf;
                pc.pf_ep = c.pf_ep;
                pc.pf_s = c.pf_s;
                pc.cK = c.c;
                BodyFrame B;
#pragma unroll
                for (int i = 0; i < 9; ++i) B.R[i] = Rb[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) B.p[i] = pb[i];
#pragma unroll
                for (int i = 0; i < 6; ++i) B.phi[i] = phi[i];
                const Wrench6 w = pf_body(pc, J.pf_ptr, J.pf_cnt, B, deriv, c.sa, NS, F::AEXT, F::CEXT, t);
#pragma unroll
                for (int i = 0; i < 6; ++i) fb[i] += w.f[i];
            }
        }
        double Fb[6], Fw[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) Fb[i] = J.I[i] * u[i] - c.c * fb[i];
        xf_b2w(Rb, pb, Fb, Fw);
        if (KEEP) {
#pragma unroll
            for (int i = 0; i < 9; ++i) SA(F::RB, i, t) = Rb[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) SA(F::PB, i, t) = pb[i];
            if (KEEP == 1) {
#pragma unroll
                for (int i = 0; i < 6; ++i) SA(F::PHI, i, t) = phi[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) SA(F::CF, i, t) = Fw[i];
        if (deriv) {
            // Jb = R I3 R' - m [p][p]   (symmetric: xx xy xz yy yz zz);  [p][p] = p p' - |p|^2 I
            const double pp = pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2];
            int e = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = a; b < 3; ++b) {
                    double v = Rb[3 * a] * J.I[0] * Rb[3 * b] + Rb[3 * a + 1] * J.I[1] * Rb[3 * b + 1] + Rb[3 * a + 2] * J.I[2] * Rb[3 * b + 2];
                    v -= m * (pb[a] * pb[b] - (a == b ? pp : 0.0));
                    SA(F::JB, e, t) = v;
                    ++e;
                }
#pragma unroll
            for (int i = 0; i < 3; ++i) SA(F::MP, i, t) = m * pb[i];
            SA(F::MS, 0, t) = m;
            // Ptl = I3[w] - [w]I3 + [I3 w]  (body frame);  Atl = -c (R Ptl R' + 2 m [p][vc]),  vc = R v_b
            double Ptl[9], W[9], IW[9], T1[9], T2[9], vc[3];
            brac3(phi, W);
            brac3(Iw, IW);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) Ptl[3 * a + b] = J.I[a] * W[3 * a + b] - W[3 * a + b] * J.I[b] + IW[3 * a + b];
            mat3_mul(Rb, Ptl, T1);
            // T2 = T1 * Rb'
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) T2[3 * a + b] = T1[3 * a] * Rb[3 * b] + T1[3 * a + 1] * Rb[3 * b + 1] + T1[3 * a + 2] * Rb[3 * b + 2];
            mat3_vec(Rb, phi + 3, vc);
            // [p][vc] = vc p' - (p.vc) I
            const double pv = pb[0] * vc[0] + pb[1] * vc[1] + pb[2] * vc[2];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    SA(F::ATL, 3 * a + b, t) = -c.c * (T2[3 * a + b] + 2.0 * m * (vc[a] * pb[b] - (a == b ? pv : 0.0)));
#pragma unroll
            for (int i = 0; i < 3; ++i) SA(F::MV, i, t) = m * vc[i];
        }
    }
    if (GROUND) {  // free flight: the 72 external-block components are all zero and are neither summed nor used
        if (NW == 1)
            c.anyext = __any_sync(0xffffffffu, myext);
        else
            c.anyext = __syncthreads_or(myext);
    }
    bsync<NW>();
    // ---- composite sums, leaves -> root (one thread per component) ---------------------------------------------
    {
        const int ncomp = deriv ? ((GROUND && !c.anyext) ? 28 : F::NCOMP) : 6;
        for (int comp = t; comp < ncomp; comp += NT) {
            double* col = c.sa + (size_t)(F::CF + comp) * NS;
            if (c.is_chain) {
                double acc = 0.0;
                for (int j = n - 1; j >= 0; --j) {
                    acc += col[j];
                    col[j] = acc;
                }
            } else {
                for (int j = n - 1; j > 0; --j) {
                    const int par = c.par_s[j];
                    if (par >= 0) col[par] += col[j];
                }
            }
        }
    }
    bsync<NW>();
    // ---- reduced residual ---------------------------------------------------------------------------------------
    if (t < n && myidx >= 0) {
        const JointConst& J = c.jc[t];
        const int r = myidx;
        const double qk = c.q[r];
        double fr = c.tau[r] + J.stiff * (J.qRest - qk) - J.damp * qdk;  // Joint.computeForce (Joint.m:448-454, 470-481)
        double dK = -J.stiff, dD = -J.damp;
        if (qk < J.qLimL) {
            fr += J.qLimK * (J.qLimL - qk) - J.qLimD * qdk;
            dK -= J.qLimK;
            dD -= J.qLimD;
        }
        if (qk > J.qLimU) {
            fr += J.qLimK * (J.qLimU - qk) - J.qLimD * qdk;
            dK -= J.qLimK;
            dD -= J.qLimD;
        }
        double Fc[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) Fc[i] = SA(F::CF, i, t);
        c.g[r] = dot6(s, Fc) - c.c * fr;
        c.sp2[r] = dK;
        c.sp1[r] = dD;
    }
    bsync<NW>();
}

// Per-joint part of the Newton-matrix assembly: from joint t's screw, its parent's V and U and its composite blocks, the
// row vector L_t (so that H[t][i] = L_t . Rt_i for t in sub(i)), and the column vectors Rt_t = [c2 ; c1 ; sq s] and Z_t
// (H[k][t] = s_k . Z_t for proper ancestors k).  Reads shared memory only; results stay in registers.
template <int NW, int GROUND, int KEEP>
__device__ __forceinline__ void columns_joint(Ctx2& c, int t, int myidx, double sq, double sqd, double sd, double* L, double* s,
                                              double* Rt, double* Z) {
    typedef Fld<GROUND, KEEP> F;
    const int NS = (NW == 1) ? 33 : 65;
    const double cc = c.c;
    if (myidx >= 0) {
        double Vp[6] = {0, 0, 0, 0, 0, 0}, Up[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = SA(F::S, i, t);
        const int par = c.par_s[t];
        if (par >= 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                Vp[i] = SA(F::V, i, par);
                Up[i] = SA(F::U, i, par);
            }
        }
        double a1[6], a2[6], a3[6], a4[6], c1[6], c2[6];
        ad_mv(s, Vp, a1);
        ad_mv(s, Up, a2);
        ad_mv(Vp, s, a3);
#pragma unroll
        for (int i = 0; i < 6; ++i) c1[i] = sqd * s[i] - sq * a1[i];
        ad_mv(c1, Vp, a4);
#pragma unroll
        for (int i = 0; i < 6; ++i) c2[i] = sd * s[i] - sq * a2[i] + cc * (sqd * a3[i] - a4[i]);
        // composite blocks of this joint
        double Jb[6], mp[3], Atl[9], mv[3], Fc[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            Jb[i] = SA(F::JB, i, t);
            Fc[i] = SA(F::CF, i, t);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            mp[i] = SA(F::MP, i, t);
            mv[i] = SA(F::MV, i, t);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) Atl[i] = SA(F::ATL, i, t);
        const double Ms = SA(F::MS, 0, t);
        const double gw[3] = {c.gx, c.gy, c.gz};
        // a = B^C s ;  B^C x = [Jb xw + mp x xv ; -mp x xw + Ms xv]
        {
            double t1[3], t2[3];
            cross3(mp, s + 3, t1);
            cross3(mp, s, t2);
            L[0] = Jb[0] * s[0] + Jb[1] * s[1] + Jb[2] * s[2] + t1[0];
            L[1] = Jb[1] * s[0] + Jb[3] * s[1] + Jb[4] * s[2] + t1[1];
            L[2] = Jb[2] * s[0] + Jb[4] * s[1] + Jb[5] * s[2] + t1[2];
            L[3] = Ms * s[3] - t2[0];
            L[4] = Ms * s[4] - t2[1];
            L[5] = Ms * s[5] - t2[2];
            // b = A^C' s : first three = Atl' sw + 2c (mv x sv)
            double t3[3], bw[3];
            cross3(mv, s + 3, t3);
            mat3T_vec(Atl, s, bw);
            L[6] = bw[0] + 2.0 * cc * t3[0];
            L[7] = bw[1] + 2.0 * cc * t3[1];
            L[8] = bw[2] + 2.0 * cc * t3[2];
            // e = C^C' s : first three = -c g x (mp x sw - Ms sv)
            double y[3] = {t2[0] - Ms * s[3], t2[1] - Ms * s[4], t2[2] - Ms * s[5]}, ew[3];
            cross3(gw, y, ew);
            const int EO = GROUND ? 12 : 9;
            L[EO] = -cc * ew[0];
            L[EO + 1] = -cc * ew[1];
            L[EO + 2] = -cc * ew[2];
            if (GROUND) {
                L[9] = L[10] = L[11] = 0.0;
                L[15] = L[16] = L[17] = 0.0;
            }
        }
        // Z = B^C c2 + A^C c1 + sq C^C s + sq ad*(s) F^C
        {
            double t1[3], t2[3], t3[3], t4[3], t5[3];
            cross3(mp, c2 + 3, t1);
            cross3(mp, c2, t2);
            Z[0] = Jb[0] * c2[0] + Jb[1] * c2[1] + Jb[2] * c2[2] + t1[0];
            Z[1] = Jb[1] * c2[0] + Jb[3] * c2[1] + Jb[4] * c2[2] + t1[1];
            Z[2] = Jb[2] * c2[0] + Jb[4] * c2[1] + Jb[5] * c2[2] + t1[2];
            Z[3] = Ms * c2[3] - t2[0];
            Z[4] = Ms * c2[4] - t2[1];
            Z[5] = Ms * c2[5] - t2[2];
            mat3_vec(Atl, c1, t3);
            cross3(mv, c1, t4);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                Z[i] += t3[i];
                Z[3 + i] -= 2.0 * cc * t4[i];
            }
            if (sq != 0.0) {
                double gs[3], az[6];
                cross3(gw, s, gs);  // g x sw
                cross3(mp, gs, t5);
                adstar_fv(s, Fc, az);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    Z[i] += sq * (az[i] - cc * t5[i]);
                    Z[3 + i] += sq * (az[3 + i] - cc * Ms * gs[i]);
                }
            }
        }
        if (GROUND && c.anyext) {
            // dense external blocks: b += Aext' s ; e += Cext' s ; Z += Aext c1 + sq Cext s
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double zr = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) {
                    const double av = SA(F::AEXT, 6 * r + m, t);
                    const double cv = SA(F::CEXT, 6 * r + m, t);
                    L[6 + m] += av * s[r];
                    L[12 + m] += cv * s[r];
                    zr += av * c1[m] + sq * cv * s[m];
                }
                Z[r] += zr;
            }
        }
        // Rt
#pragma unroll
        for (int i = 0; i < 6; ++i) Rt[i] = c2[i];
        if (GROUND) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                Rt[6 + i] = c1[i];
                Rt[12 + i] = sq * s[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                Rt[6 + i] = c1[i];
                Rt[9 + i] = sq * s[i];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// eval_columns2: out (nr x ld column-major) = scale * ( sq dg/dq + sqd dg/dqdot + sd dg/d(dqtmp) ), from the composite blocks.
// ---------------------------------------------------------------------------------------------
template <int NW, int GROUND, int KEEP>
__device__ void eval_columns2(Ctx2& c, double sq, double sqd, double sd, double scale, double* out) {
    typedef Fld<GROUND, KEEP> F;
    const int t = threadIdx.x;
    const int n = c.n, NS = (NW == 1) ? 33 : 65, ld = c.ld;
    const double cc = c.c;
    const int myidx = (t < n) ? c.ie_s[t].x : -1;
    double Rt[F::NL];  // [c2 (6) ; c1 (3 or 6) ; sq s (3 or 6)]
    double Z[6], L[F::NL], s[6];
    columns_joint<NW, GROUND, KEEP>(c, t, myidx, sq, sqd, sd, L, s, Rt, Z);
    if (myidx >= 0) {
        // W_t = [L_t ; s_t], joint-major so that the rows below are fetched with 128-bit broadcast loads
        {
            double* W = c.sa + (size_t)t * F::NW_;
#pragma unroll
            for (int i = 0; i < F::NL; ++i) W[i] = L[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) W[F::NL + i] = s[i];
        }
    }
    bsync<NW>();
    // ---- entries of column idx[t]: row k is L_k.Rt_i for k in sub(i), s_k.Z_i for k a proper ancestor of i ------------
    if (myidx >= 0) {
        const int i = t, iend = c.ie_s[t].y;
        double* col = out + (size_t)myidx * ld;
        const double dg = -cc * (sq * c.sp2[myidx] + sqd * c.sp1[myidx]);  // Kr, Dr of Joint.m:470-481
        for (int k = 0; k < n; ++k) {
            const int2 ie = c.ie_s[k];
            if (ie.x < 0) continue;
            const double2* W = reinterpret_cast<const double2*>(c.sa + (size_t)k * F::NW_);
            double v = 0.0, v2 = 0.0;
            if (k >= i && k < iend) {
#pragma unroll
                for (int e = 0; e < F::NL / 2; ++e) {
                    const double2 w = W[e];
                    v = fma(w.x, Rt[2 * e], v);
                    v2 = fma(w.y, Rt[2 * e + 1], v2);
                }
                if (k == i) v += dg;
            } else if (k < i && i < ie.y) {
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const double2 w = W[F::NL / 2 + e];
                    v = fma(w.x, Z[2 * e], v);
                    v2 = fma(w.y, Z[2 * e + 1], v2);
                }
            }
            col[ie.x] = scale * (v + v2);
        }
    }
    if (GROUND == 2 && c.npf > 0)  // off-diagonal blocks of the point forces (same thread owns the column: no barrier needed before)
        pf_cross_pass(c, t, myidx, Rt + 6, Rt + 12, scale, out, ld, c.sa, F::NW_, F::NL);
    bsync<NW>();
}

// ---------------------------------------------------------------------------------------------
// Warp-register LU with partial pivoting + solve for one warp (nr <= 32).  Lane r holds row r of H; rows never move:
// the pivot row of step k is broadcast with shuffles, `pos` tracks each row's LAPACK position so that ties are broken
// exactly like dgetf2's idamax (first maximum in the current row order).  Solves H x = scale * rhs, x -> dx (shared).
// If write_back, the factored image (unit-lower L below the diagonal, rows in pivot order) goes back to H and
// perm[k] = original row in position k (== Hp of lu(H,'vector')).
// ---------------------------------------------------------------------------------------------
template <int NR>
__device__ __noinline__ void lu_solve_warp_t(int nr, int ld, double* H, int* perm, const double* rhs, double scale, double* dx,
                                                bool write_back, double2* /*rowbuf*/) {
    // The matrix is padded to NR x NR with an identity block (lanes/columns >= nr), so no loop below needs a runtime
    // bound: padding rows are never chosen before the real ones (their column entries are exactly 0) and eliminate nothing.
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double a[NR];
#pragma unroll
    for (int cI = 0; cI < NR; ++cI) a[cI] = (cI < nr && lane < nr) ? H[(size_t)cI * ld + lane] : ((cI == lane) ? 1.0 : 0.0);
    double b = (lane < nr) ? scale * rhs[lane] : 0.0;
    bool done = lane >= NR;
    int pos = lane, mypos = -1;
    double rdiag = 1.0;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        // argmax |a[k]| over the rows that are not pivots yet; ties -> smallest LAPACK position.  |v| >= 0, so the IEEE bit
        // pattern orders like the value: one warp reduction on the high words decides unless two rows agree in their
        // top 32 bits, and only then the low words and the positions are consulted (warp-uniform branch).
        const double v = fabs(a[k]);
        const unsigned hi = done ? 0u : (unsigned)__double2hiint(v);
        const unsigned mh = __reduce_max_sync(FULL, hi);
        const bool c1 = !done && hi == mh;
        unsigned cand = __ballot_sync(FULL, c1);
        if (cand & (cand - 1)) {
            const unsigned lo = c1 ? (unsigned)__double2loint(v) : 0u;
            const unsigned ml = __reduce_max_sync(FULL, lo);
            const bool c2 = c1 && lo == ml;
            const unsigned pm = __reduce_min_sync(FULL, c2 ? (unsigned)pos : 0xffffu);
            cand = __ballot_sync(FULL, c2 && (unsigned)pos == pm);
        }
        const int src = __ffs(cand) - 1;
        const int kl = __ffs(__ballot_sync(FULL, !done && pos == k)) - 1;
        const int pos_src = __shfl_sync(FULL, pos, src);
        if (lane == kl) pos = pos_src;
        if (lane == src) {
            pos = k;
            done = true;
            mypos = k;
        }
        if (lane == k) perm[k] = src;
        const double piv = __shfl_sync(FULL, a[k], src);
        const double rp = __drcp_rn(piv);  // == 1.0 / piv, correctly rounded
        rdiag = (lane == src) ? rp : rdiag;
        const double l = done ? 0.0 : a[k] * rp;
        a[k] = done ? a[k] : l;
#pragma unroll
        for (int cI = k + 1; cI < NR; ++cI) {
            const double u = __shfl_sync(FULL, a[cI], src);
            a[cI] = fma(-l, u, a[cI]);  // l == 0 for rows that are already pivots
        }
        const double ub = __shfl_sync(FULL, b, src);
        b = fma(-l, ub, b);
    }
    __syncwarp();
    if (write_back && lane < nr) {
#pragma unroll
        for (int cI = 0; cI < NR; ++cI)
            if (cI < nr) H[(size_t)cI * ld + mypos] = a[cI];
    }
    // back substitution U x = y : row `mypos` of U lives in this lane, y_mypos = b
#pragma unroll
    for (int k = NR - 1; k >= 0; --k) {
        const int src = perm[k];
        const double xk = __shfl_sync(FULL, b * rdiag, src);
        b = (mypos < k) ? fma(-a[k], xk, b) : b;
        if (lane == k && k < nr) dx[k] = xk;
    }
    __syncwarp();
}

// (A variant that broadcast the pivot row through shared memory instead of shuffles ran at the same speed on B200 --
// profiles/r01_ab_lu_occupancy.log -- and was removed in round 2; the tensor-core kernels have their own blocked LU, rmx_tc.cuh.)
__device__ __forceinline__ void lu_solve_warp(int nr, int ld, double* H, int* perm, const double* rhs, double scale, double* dx,
                                              bool write_back, double2* rowbuf) {
    if (nr <= 8)
        lu_solve_warp_t<8>(nr, ld, H, perm, rhs, scale, dx, write_back, rowbuf);
    else if (nr <= 16)
        lu_solve_warp_t<16>(nr, ld, H, perm, rhs, scale, dx, write_back, rowbuf);
    else if (nr <= 24)
        lu_solve_warp_t<24>(nr, ld, H, perm, rhs, scale, dx, write_back, rowbuf);
    else
        lu_solve_warp_t<32>(nr, ld, H, perm, rhs, scale, dx, write_back, rowbuf);
}

}  // namespace rmx
