// rmx_device.cuh -- device code of the batched RedMax stepper (sm_100a, FP64).
//
// One thread block integrates one rollout for all its time steps (persistent over the whole time loop):
// the joint tree lives in shared memory, thread t owns joint/body t in the O(n) phases and column idx[t] of the
// Newton matrix H in the O(n^2) phase.  Nothing but q(t), qdot(t) (and the adjoint tape) goes to HBM.
//
// What it replaces in the reference (matlab-diff/):
//   Joint.update + JointRevolute.update_ + Body.update      (Joint.m:382, JointRevolute.m:29, Body.m:70)  -> fk_*
//   Joint.computeJacobian (dense J,Jdot,dJdq,dJdotdq)       (Joint.m:490-613)   -> never formed: world-frame
//       joint screws s_k give J*x, Jdot*qdot and J'*y as O(n) sweeps, and column i of H as a sweep over sub(i)
//   Body.computeMassGrav, Joint.computeForce, ForceGroundCuboid.computeValues_
//                                                            (Body.m:83, Joint.m:437, ForceGroundCuboid.m:54)
//   computeValues + evalBDF1/evalSDIRK2a/evalSDIRK2b/evalBDF2 (driverRedMaxBDF1.m:160-243, driverRedMaxBDF2.m:194-349)
//   newton (forward, with line search)                       (driverRedMaxBDF1.m:94-157)
//   newton (adjoint driver, LU tape, no line search)         (driverRedMaxAdjointBDF1.m:105-146)
//   simLoop                                                  (driverRedMaxBDF1.m:57-91, driverRedMaxBDF2.m:57-125)
//
// Maths (derivation in DESIGN.md): with world-frame screws s_k, V_j = sum_{k in anc(j)} s_k qd_k,
// U_j = sum_k (s_k d_k + c ad(V_p(k)) s_k qd_k), body wrench F_j = X_j^{-T}[ M u - c(fcor + fgrav + fext) ],
//   g_k = s_k' * sum_{j in sub(k)} F_j - c fr_k                      ( == M*dqtmp - c*f of the reference )
// and column i of H = dg/dq_i is  s_k' * sum_{j in sub(k) ^ sub(i)} T^i_j  (+ s_k' ad*(s_i) Fsub_i for ancestors k of i),
// where T^i_j needs only two per-column constants c1_i, c2_i transported into body j's frame.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rmx {

// ---------------------------------------------------------------------------------------------
// Scene constants (device global memory, read through the read-only path; identical for all rollouts)
// ---------------------------------------------------------------------------------------------
struct JointConst {
    double R0[9];   // E0_pj rotation (row-major)
    double p0[3];   // E0_pj translation
    double Rji[9];  // E0_ji rotation
    double pji[3];  // E0_ji translation
    double axis[3]; // S = [axis; 0]                          (JointRevolute.m:36)
    double axn[3];  // axis re-normalised as se3.aaToMat does (se3.m:119-123)
    double I[6];    // diag inertia
    double hs[3];   // half side lengths
    double stiff, damp, qRest, qLimL, qLimU, qLimK, qLimD;
    double gxg[3], gng[3], gkn, gkt, gkd, gmu;  // ForceGroundCuboid
    int parent;     // internal index of parent joint or -1
    int idx;        // reduced index (reference numbering) or -1 for a fixed joint
    int end;        // j + subtree size (DFS preorder => sub(j) = [j, end))
    int axtype;     // 0 general, 1 X, 2 Y, 3 Z (se3.aaToMat special cases); sign in axsign
    int axsign;     // +1 / -1
    int has_ground;
    int prismatic;  // 1: Q = trans(axis q), S = [0; axis] (JointPrismatic.m:28-34); 0: revolute (or fixed if idx < 0)
    int pf_ptr;     // CSR into DevScene.pf_ep: force attachments on this body (entry = PF_MAXPTS*force + attachment)
    int pf_cnt;
    int ends_ptr;   // CSR into DevScene.ends_list: joints k whose subtree ends exactly at this index
    int ends_cnt;
    int chart_mid;  // this one-DOF virtual joint is the middle Euler angle of a spherical / Free3D joint: 1 det T = cos q, 2 sin q
};

// Forces between body points (matlab-diff/+redmax): ForcePointPoint.m (linear, zero rest length), ForceSpringDamper.m
// (ForceSpringGeneric.m, along the line, rest length L) and ForceCable.m (ForceSpringMultiPointGeneric.m, routed through
// up to PF_MAXPTS points, pulls only when stretched)
constexpr int PF_MAXPTS = 4;
struct PointForce {
    int kind;                  // 0 point-point, 1 spring-damper, 2 cable
    int npts;                  // 2, or the number of cable points
    int body[PF_MAXPTS];       // internal joint index of each attachment's body, -1 = world
    double x[PF_MAXPTS][3];    // application points in body (or world) coordinates
    double ks, kd;
    double L;                  // rest length (kinds 1, 2)
    int rec_off;               // offset (doubles) into the shared-memory scratch of this force's attachment records [npts][PF_REC]
    int blk_off;               // ... and of its cross blocks: ordered pair (k, k2) at blk_off + (k npts + k2) PF_BLK
};
constexpr int PF_MAX = 8;        // forces per scene
constexpr int PF_REC = 18;       // published per attachment: R[9] p[3] phi[6]
constexpr int PF_BLK = 72;       // per ordered pair: Aext_ab[36], Cext_ab[36] (world frame, row-major)

struct DevScene {
    int n, nr;
    int is_chain;     // every joint's parent is j-1 (no subtree fix-ups needed)
    int has_ground;
    double grav[3];
    const JointConst* jc;
    const int* ends_list;
    const int* anc;   // [nrounds][n] 2^r-th ancestor of joint j or -1 (pointer-jumping scans of the fast path)
    int nrounds;
    const PointForce* pf;  // [npf]
    const int* pf_ep;      // endpoint lists (see JointConst::pf_ptr)
    int npf;
    int has_chart;         // some joint has chart_mid set
};

struct StepOpts {
    int scheme, nsteps, iterMax, iterLsMax, tau_mode, adjoint_newton;
    int shortcuts;   // forward Newton: skip work of stalled solves whose outcome is known bit for bit (newton_forward)
    double h, tol, dxMax;
    double lin_tol;  // Krylov linear solve: relative residual tolerance (c++/PCG Solver.h:43: 1e-6)
    int lin_maxit;
};

// stages of the implicit step
enum { ST_BDF1 = 0, ST_SDIRK_A = 1, ST_SDIRK_B = 2, ST_BDF2 = 3, ST_DIRECT = 4 };

constexpr int REC1 = 34;  // Rb[9] pb[3] phi[6] s[6] I[6] gb[3] pad
constexpr int REC2 = 30;  // Rw[9] pw[3] V[6] U[6] F[6]
constexpr int NVEC = 14;  // q qd dq g dx x0 tau hq0 hqd0 hq1 hqd1 + 3 spare

__host__ __device__ inline int h_ld(int nr) { return nr | 1; }  // odd leading dimension: conflict-free columns

__host__ __device__ inline size_t smem_doubles(int n, int nr, bool ground) {
    size_t d = (size_t)n * (REC1 + REC2) + (size_t)NVEC * nr + (size_t)nr * h_ld(nr) + 2 * 8 + 8;
    if (ground) d += (size_t)n * 72;
    return (d + 1) & ~(size_t)1;
}

// ---------------------------------------------------------------------------------------------
// small dense helpers (row-major 3x3; everything inlines into registers)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void cross3_acc(const double* a, const double* b, double* c) {
    c[0] += a[1] * b[2] - a[2] * b[1];
    c[1] += a[2] * b[0] - a[0] * b[2];
    c[2] += a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void mat3_vec(const double* A, const double* x, double* y) {
    y[0] = A[0] * x[0] + A[1] * x[1] + A[2] * x[2];
    y[1] = A[3] * x[0] + A[4] * x[1] + A[5] * x[2];
    y[2] = A[6] * x[0] + A[7] * x[1] + A[8] * x[2];
}
__device__ __forceinline__ void mat3T_vec(const double* A, const double* x, double* y) {
    y[0] = A[0] * x[0] + A[3] * x[1] + A[6] * x[2];
    y[1] = A[1] * x[0] + A[4] * x[1] + A[7] * x[2];
    y[2] = A[2] * x[0] + A[5] * x[1] + A[8] * x[2];
}
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// motion vector world -> body:  y = Ad(E^-1) x,  E = (R,p)
__device__ __forceinline__ void xm_w2b(const double* R, const double* p, const double* x, double* y) {
    double t[3];
    cross3(x, p, t);
    t[0] += x[3];
    t[1] += x[4];
    t[2] += x[5];
    mat3T_vec(R, x, y);
    mat3T_vec(R, t, y + 3);
}
// force vector body -> world:  Y = Ad(E^-1)^T F
__device__ __forceinline__ void xf_b2w(const double* R, const double* p, const double* F, double* Y) {
    mat3_vec(R, F + 3, Y + 3);
    mat3_vec(R, F, Y);
    cross3_acc(p, Y + 3, Y);
}
// ad(a) b for motion vectors [w; v]
__device__ __forceinline__ void ad_mv(const double* a, const double* b, double* o) {
    cross3(a, b, o);
    cross3(a, b + 3, o + 3);
    cross3_acc(a + 3, b, o + 3);
}
// -ad(s)^T F for force vectors [tau; f]
__device__ __forceinline__ void adstar_fv(const double* s, const double* F, double* o) {
    cross3(s, F, o);
    cross3_acc(s + 3, F + 3, o);
    cross3(s, F + 3, o + 3);
}
__device__ __forceinline__ double dot6(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
__device__ __forceinline__ void ld_vec(const double* __restrict__ src, double* dst, int n2) {
    // n2 = number of double2 (src 16B aligned)
    const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int i = 0; i < n2; ++i) {
        double2 v = s2[i];
        dst[2 * i] = v.x;
        dst[2 * i + 1] = v.y;
    }
}

// ---------------------------------------------------------------------------------------------
// block-level primitives (NW warps per rollout; all decisions are block-uniform)
// ---------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void bsync() {
    if (NW == 1)
        __syncwarp();
    else
        __syncthreads();
}

// Lockstep groups (one-warp forward kernels launched with blockDim.y = G > 1): G independent rollouts, one per warp, meet at the
// top of every evaluation pass, so that the G warps of a block walk through the same code at the same time and share its
// instruction-cache lines -- the kernels with external forces are bound by instruction fetch (SM instruction-cache hit rate
// 62 %, GPC-level cache requests at 90 % of peak: profiles/r02_residency_ab.log).  Named barrier 1 with an AND vote: a warp
// that has finished all of its work keeps arriving with done = true until every warp of the block has.  Every blocking
// wait of a grouped warp must keep calling this, or its partners stall.
__device__ __forceinline__ bool group_barrier(int G, bool done) {
    if (G <= 1) return done;
    int r;
    __syncwarp();  // the barrier instruction is .aligned: the whole warp has to arrive converged (compute-sanitizer synccheck)
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.s32 q, %1, 0;\n\t"
        "bar.red.and.pred p, 1, %2, q;\n\t"
        "selp.s32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(r)
        : "r"((int)done), "r"(G * 32)
        : "memory");
    return r != 0;
}

template <int NW>
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (NW == 1) return v;
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) s += red[i];
    return s;
}

// argmax of (|v|, first index on ties) over the block -- LAPACK idamax semantics for the LU pivot
template <int NW>
__device__ __forceinline__ int block_argmax(double v, int idx, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
    if (NW == 1) return idx;
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        red[w] = v;
        red[8 + w] = (double)idx;
    }
    __syncthreads();
    double bv = red[0];
    int bi = (int)red[8];
#pragma unroll
    for (int i = 1; i < NW; ++i) {
        double ov = red[i];
        int oi = (int)red[8 + i];
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    return bi;
}

// ---------------------------------------------------------------------------------------------
// per-rollout context: shared-memory carve-up
// ---------------------------------------------------------------------------------------------
struct Ctx {
    int n, nr, ld;
    int group;  // one-warp forward kernels: rollouts (warps) per block that run their evaluation passes in lockstep (1: none)
    const JointConst* __restrict__ jc;
    const int* __restrict__ ends_list;
    double gx, gy, gz;
    int is_chain;
    // shared memory
    double* rec1;
    double* rec2;
    double* KD;   // [n][72] body-frame Kext, Dext (ground only)
    double* q;    // current Newton iterate x  (reference numbering)
    double* qd;
    double* dq;   // dqtmp
    double* g;
    double* dx;
    double* x0;
    double* tau;
    double* hq0;  // history: q0, qdot0, q1, qdot1 of the reference's joint.q0/qdot0/q1/qdot1
    double* hqd0;
    double* hq1;
    double* hqd1;
    double* sp0;  // spare vectors (adjoint: dPdq, perm ...)
    double* sp1;
    double* sp2;
    double* H;
    double* red;  // reduction scratch [16]
    // point forces (composite kernels only)
    const PointForce* __restrict__ pf;
    const int* __restrict__ pf_ep;
    int npf;
    double* pf_s;  // attachment records and cross blocks (PointForce::rec_off / blk_off)
    // stage coefficients
    int stage;
    double h;
    double c;     // cK: h^2, (ah)^2, (ah)^2, (4/9)h^2
    double beta;  // d(qdot)/dq: 1/h, 1/(ah), 1/(ah), 3/(2h)
};

__device__ __forceinline__ void ctx_carve(Ctx& c, double* sm, int n, int nr, bool ground) {
    c.n = n;
    c.nr = nr;
    c.ld = h_ld(nr);
    double* p = sm;
    c.rec1 = p;
    p += (size_t)n * REC1;
    c.rec2 = p;
    p += (size_t)n * REC2;
    c.KD = p;
    if (ground) p += (size_t)n * 72;
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
    p = (double*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    c.H = p;
}

#define SDIRK_A_CONST 0.29289321881345254  // (2 - sqrt(2))/2 rounded to nearest double (driverRedMaxBDF2.m:75)

// Kinematic part of evalBDF1 / evalSDIRK2a / evalSDIRK2b / evalBDF2: qdot and dqtmp from the iterate x, written
// with explicit roundings in the reference's operation order (driverRedMaxBDF1.m:167-170, driverRedMaxBDF2.m:203-206,
// 237-241, 271-275) so the state fed to the dynamics is bit-identical to the reference's.
__device__ __forceinline__ void stage_kin(int stage, double h, double x, double q0, double qd0, double q1, double qd1,
                                          double& qd, double& dq) {
    const double a = SDIRK_A_CONST;
    if (stage == ST_BDF1) {
        double d = __dsub_rn(x, q0);
        dq = __dsub_rn(d, __dmul_rn(h, qd0));
        qd = __ddiv_rn(d, h);
    } else if (stage == ST_SDIRK_A) {
        double ah = __dmul_rn(a, h);
        double d = __dsub_rn(x, q0);
        dq = __dsub_rn(d, __dmul_rn(ah, qd0));
        qd = __ddiv_rn(d, ah);
    } else if (stage == ST_SDIRK_B) {
        // q1/qd1 slots hold (qa, qdota) here (driverRedMaxBDF2.m:81)
        double ah = __dmul_rn(a, h);
        double d = __dsub_rn(x, q0);
        double t1 = __dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(2.0, a), 1.0), h), qd0);
        double t2 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, __dsub_rn(1.0, a)), h), qd1);
        dq = __dsub_rn(__dsub_rn(d, t1), t2);
        double t3 = __dmul_rn(__dmul_rn(__dsub_rn(1.0, a), h), qd1);
        qd = __ddiv_rn(__dsub_rn(d, t3), ah);
    } else if (stage == ST_DIRECT) {
        qd = qd0;  // test hook: state given directly
        dq = q1;
    } else {
        double e = __dadd_rn(__dsub_rn(x, __dmul_rn(4.0 / 3.0, q1)), __dmul_rn(1.0 / 3.0, q0));
        dq = __dadd_rn(__dsub_rn(e, __dmul_rn(__dmul_rn(8.0 / 9.0, h), qd1)), __dmul_rn(__dmul_rn(2.0 / 9.0, h), qd0));
        qd = __dmul_rn(__ddiv_rn(3.0, __dmul_rn(2.0, h)), e);
    }
}

__device__ __forceinline__ void stage_coef(Ctx& c, int stage, double h) {
    const double a = SDIRK_A_CONST;
    c.stage = stage;
    c.h = h;
    if (stage == ST_BDF1) {
        c.c = h * h;
        c.beta = 1.0 / h;
    } else if (stage == ST_SDIRK_A || stage == ST_SDIRK_B) {
        double ah = a * h;
        c.c = ah * ah;
        c.beta = 1.0 / ah;
    } else {
        c.c = (4.0 / 9.0) * (h * h);
        c.beta = 3.0 / (2.0 * h);
    }
}

// se3.aaToMat (se3.m:111-176) with its axis-aligned special cases
__device__ __forceinline__ void aa_to_mat(const JointConst& J, double angle, double* R) {
    double sn, cs;
    if (J.axtype != 0 && J.axsign < 0) angle = -angle;
    sincos(angle, &sn, &cs);
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    if (J.axtype == 3) {
        R[0] = cs; R[1] = -sn; R[3] = sn; R[4] = cs;
    } else if (J.axtype == 1) {
        R[4] = cs; R[5] = -sn; R[7] = sn; R[8] = cs;
    } else if (J.axtype == 2) {
        R[0] = cs; R[2] = sn; R[6] = -sn; R[8] = cs;
    } else {
        const double ax = J.axn[0], ay = J.axn[1], az = J.axn[2];
        const double t = 1.0 - cs;
        const double xz = ax * az, xy = ax * ay, yz = ay * az;
        R[0] = t * ax * ax + cs; R[1] = t * xy - sn * az; R[2] = t * xz + sn * ay;
        R[3] = t * xy + sn * az; R[4] = t * ay * ay + cs; R[5] = t * yz - sn * ax;
        R[6] = t * xz - sn * ay; R[7] = t * yz + sn * ax; R[8] = t * az * az + cs;
    }
}

// ---------------------------------------------------------------------------------------------
// ForceGroundCuboid.computeValues_ (ForceGroundCuboid.m:54-153) for one body, in body coordinates.
// nb = R' * ng.  Adds the wrench to fext[6]; if DERIV adds the 6x6 Km, Dm (row-major) to K, D.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void outer_acc6(double* M, double a, const double* u, const double* v) {
    // M(6x6) += a * u * v'
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double au = a * u[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) M[6 * i + j] += au * v[j];
    }
}
// M(6x6) += a * G' * [A  B]   with G = [brac(xl)' I] (3x6), A, B 3x3 row-major (B may be null => zeros)
__device__ __forceinline__ void gt_acc(double* M, double a, const double* xl, const double* A, const double* B) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double col[3] = {A[j], A[3 + j], A[6 + j]};
        double top[3];
        cross3(xl, col, top);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            M[6 * i + j] += a * top[i];
            M[6 * (3 + i) + j] += a * col[i];
        }
    }
    if (B) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double col[3] = {B[j], B[3 + j], B[6 + j]};
            double top[3];
            cross3(xl, col, top);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                M[6 * i + 3 + j] += a * top[i];
                M[6 * (3 + i) + 3 + j] += a * col[i];
            }
        }
    }
}
__device__ __forceinline__ void brac3(const double* x, double* S) {
    S[0] = 0; S[1] = -x[2]; S[2] = x[1];
    S[3] = x[2]; S[4] = 0; S[5] = -x[0];
    S[6] = -x[1]; S[7] = x[0]; S[8] = 0;
}

template <bool DERIV>
__device__ void ground_body(const JointConst& J, const double* R, const double* p, const double* phi, double* fext,
                            double* K, double* D) {
    const double kn = J.gkn, kt = J.gkt, kd = J.gkd, mu = J.gmu;
    double nb[3];
    mat3T_vec(R, J.gng, nb);
    const double dp = J.gng[0] * (p[0] - J.gxg[0]) + J.gng[1] * (p[1] - J.gxg[1]) + J.gng[2] * (p[2] - J.gxg[2]);
    // (kept a loop on purpose: fully unrolled, the kernel grows from 9.9k to 15.9k instructions and runs 1.64x slower --
    // profiles/r02_ground_unroll_ab.log -- the external-force kernels are bound by their executed code footprint)
#pragma unroll 1
    for (int ci = 0; ci < 8; ++ci) {
        // corner order of ForceGroundCuboid.m:73-82: x sign is the slowest bit
        double xl[3] = {(ci & 4) ? J.hs[0] : -J.hs[0], (ci & 2) ? J.hs[1] : -J.hs[1], (ci & 1) ? J.hs[2] : -J.hs[2]};
        const double d = nb[0] * xl[0] + nb[1] * xl[1] + nb[2] * xl[2] + dp;
        if (d > 0) continue;
        // point velocity in body coordinates: G*phi = w x xl + v
        double gp[3];
        cross3(phi, xl, gp);
        gp[0] += phi[3]; gp[1] += phi[4]; gp[2] += phi[5];
        const double nbg = nb[0] * gp[0] + nb[1] * gp[1] + nb[2] * gp[2];
        const double fn = -(kn * d + kd * nbg);
        double fb[3] = {fn * nb[0], fn * nb[1], fn * nb[2]};
        cross3_acc(xl, fb, fext);
        fext[3] += fb[0]; fext[4] += fb[1]; fext[5] += fb[2];
        double gn[6];  // G' * nb
        if (DERIV) {
            cross3(xl, nb, gn);
            gn[3] = nb[0]; gn[4] = nb[1]; gn[5] = nb[2];
            // Km -= G' [kn*tmp1 + kd*tmp2, kn*RNR];  kn*tmp1 + kd*tmp2 = -fn*brac(nb) - nb * (nb x (kn xl + kd gp))'
            double y[3] = {kn * xl[0] + kd * gp[0], kn * xl[1] + kd * gp[1], kn * xl[2] + kd * gp[2]};
            double w[3];
            cross3(nb, y, w);
            double A1[9], A2[9], S[9];
            brac3(nb, S);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    A1[3 * i + j] = -fn * S[3 * i + j] - nb[i] * w[j];
                    A2[3 * i + j] = kn * nb[i] * nb[j];
                }
            gt_acc(K, -1.0, xl, A1, A2);
            outer_acc6(D, -kd, gn, gn);
        }
        if (mu == 0) continue;
        // tangential velocity in body coordinates: ab = (I - nb nb') gp
        double ab[3] = {gp[0] - nb[0] * nbg, gp[1] - nb[1] * nbg, gp[2] - nb[2] * nbg};
        const double anorm = sqrt(ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2]);
        if (mu * fabs(kn * d) > kt * anorm) {
            // static friction (ForceGroundCuboid.m:121-133)
            double fs[3] = {-kt * ab[0], -kt * ab[1], -kt * ab[2]};
            cross3_acc(xl, fs, fext);
            fext[3] += fs[0]; fext[4] += fs[1]; fext[5] += fs[2];
            if (DERIV) {
                // D += -kt G' B G,  B = I - nb nb';  B*G = [B*brac(xl)', B]
                double Bm[9], BX[9], Sx[9];
                brac3(xl, Sx);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) Bm[3 * i + j] = (i == j ? 1.0 : 0.0) - nb[i] * nb[j];
                // BX = B * Sx'
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        BX[3 * i + j] = Bm[3 * i] * Sx[3 * j] + Bm[3 * i + 1] * Sx[3 * j + 1] + Bm[3 * i + 2] * Sx[3 * j + 2];
                gt_acc(D, -kt, xl, BX, Bm);
                // K += -kt G' [brac(ab) - B*brac(gp), 0]
                double Sa[9], Sg[9], Ms[9];
                brac3(ab, Sa);
                brac3(gp, Sg);
                mat3_mul(Bm, Sg, Ms);
#pragma unroll
                for (int i = 0; i < 9; ++i) Ms[i] = Sa[i] - Ms[i];
                gt_acc(K, -kt, xl, Ms, nullptr);
            }
        } else {
            // dynamic friction (ForceGroundCuboid.m:134-150)
            const double mukn = mu * kn;
            const double ia = 1.0 / anorm;
            double tb[3] = {ab[0] * ia, ab[1] * ia, ab[2] * ia};
            double fd[3] = {-mukn * d * tb[0], -mukn * d * tb[1], -mukn * d * tb[2]};
            cross3_acc(xl, fd, fext);
            fext[3] += fd[0]; fext[4] += fd[1]; fext[5] += fd[2];
            if (DERIV) {
                // Ab = (|a|^2 I - a a')/|a|^3 ; AB = Ab * B
                const double a2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
                const double ia3 = 1.0 / (anorm * anorm * anorm);
                double Ab[9], Bm[9], AB[9];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        Ab[3 * i + j] = ((i == j ? a2 : 0.0) - ab[i] * ab[j]) * ia3;
                        Bm[3 * i + j] = (i == j ? 1.0 : 0.0) - nb[i] * nb[j];
                    }
                mat3_mul(Ab, Bm, AB);
                // D += -mukn*d * G' AB G ;  AB*G = [AB*brac(xl)', AB]
                double Sx[9], ABX[9];
                brac3(xl, Sx);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        ABX[3 * i + j] = AB[3 * i] * Sx[3 * j] + AB[3 * i + 1] * Sx[3 * j + 1] + AB[3 * i + 2] * Sx[3 * j + 2];
                gt_acc(D, -mukn * d, xl, ABX, AB);
                // K += -mukn G' (K1 + K2 + K3): K1 = d [brac(tb) 0], K2 = tb * gn', K3 = -d AB [brac(gp) 0]
                double St[9], Sg[9], K13[9];
                brac3(tb, St);
                brac3(gp, Sg);
                mat3_mul(AB, Sg, K13);
#pragma unroll
                for (int i = 0; i < 9; ++i) K13[i] = d * (St[i] - K13[i]);
                double A[9], Bq[9];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        A[3 * i + j] = K13[3 * i + j] + tb[i] * gn[j];
                        Bq[3 * i + j] = tb[i] * gn[3 + j];
                    }
                gt_acc(K, -mukn, xl, A, Bq);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One evaluation of the implicit-step residual g (and, if deriv, the Newton matrix H) at iterate c.q
//   == evalBDF1/evalSDIRK2a/evalSDIRK2b/evalBDF2 + jroot.update + computeValues of the reference.
// seeds (sq, sqd, sd) select what the column sweep differentiates: (1, beta, 1) -> H;  (0,0,1) -> M;  (0,1,0) -> -c*D.
// ---------------------------------------------------------------------------------------------
template <int NW, int GROUND>
__device__ void eval_base(Ctx& c, bool deriv) {
    const int t = threadIdx.x;
    const int n = c.n;
    // ---- Phase A: per-dof kinematics of the stage, per-joint local rotation --------------------------------
    if (t < c.nr) {
        double qd, dq;
        stage_kin(c.stage, c.h, c.q[t], c.hq0[t], c.hqd0[t], c.hq1[t], c.hqd1[t], qd, dq);
        c.qd[t] = qd;
        c.dq[t] = dq;
    }
    if (t < n) {
        const JointConst& J = c.jc[t];
        double* r2 = c.rec2 + (size_t)t * REC2;
        if (J.idx >= 0) {
            double Rq[9], Rl[9];
            aa_to_mat(J, c.q[J.idx], Rq);
            mat3_mul(J.R0, Rq, Rl);
#pragma unroll
            for (int i = 0; i < 9; ++i) r2[i] = Rl[i];
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) r2[i] = J.R0[i];
        }
    }
    bsync<NW>();
    // ---- Phase B: serial root->leaves sweep (warp 0, all lanes redundantly; lane 0 stores) ------------------
    if (t < 32) {
        double Rp[9], pp[3], Vp[6], Up[6];
        int prev = -2;
        for (int j = 0; j < n; ++j) {
            const JointConst& J = c.jc[j];
            double* r2 = c.rec2 + (size_t)j * REC2;
            double* r1 = c.rec1 + (size_t)j * REC1;
            const int par = J.parent;
            if (par != prev) {
                if (par < 0) {
                    Rp[0] = 1; Rp[1] = 0; Rp[2] = 0; Rp[3] = 0; Rp[4] = 1; Rp[5] = 0; Rp[6] = 0; Rp[7] = 0; Rp[8] = 1;
                    pp[0] = pp[1] = pp[2] = 0;
#pragma unroll
                    for (int i = 0; i < 6; ++i) Vp[i] = Up[i] = 0;
                } else {
                    const double* rp = c.rec2 + (size_t)par * REC2;
#pragma unroll
                    for (int i = 0; i < 9; ++i) Rp[i] = rp[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) pp[i] = rp[9 + i];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        Vp[i] = rp[12 + i];
                        Up[i] = rp[18 + i];
                    }
                }
            }
            double Rl[9], Rw[9], pw[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) Rl[i] = r2[i];
            mat3_mul(Rp, Rl, Rw);
            mat3_vec(Rp, J.p0, pw);
            pw[0] += pp[0]; pw[1] += pp[1]; pw[2] += pp[2];
            double s[6] = {0, 0, 0, 0, 0, 0};
            if (J.idx >= 0) {
                mat3_vec(Rw, J.axis, s);
                cross3(pw, s, s + 3);
                const double qd = c.qd[J.idx], dq = c.dq[J.idx];
                double sd[6];
                ad_mv(Vp, s, sd);
                const double cq = c.c * qd;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    Up[i] += s[i] * dq + cq * sd[i];
                    Vp[i] += s[i] * qd;
                }
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) Rp[i] = Rw[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) pp[i] = pw[i];
            __syncwarp();
            if (t == 0) {
#pragma unroll
                for (int i = 0; i < 9; ++i) r2[i] = Rw[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) r2[9 + i] = pw[i];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    r2[12 + i] = Vp[i];
                    r2[18 + i] = Up[i];
                    r1[18 + i] = s[i];
                }
            }
            __syncwarp();
            prev = j;
        }
    }
    bsync<NW>();
    // ---- Phase C: per-body frame, twist, wrench (thread t <-> body t) ---------------------------------------
    if (t < n) {
        const JointConst& J = c.jc[t];
        double* r2 = c.rec2 + (size_t)t * REC2;
        double* r1 = c.rec1 + (size_t)t * REC1;
        double Rw[9], pw[3], V[6], U[6];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rw[i] = r2[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) pw[i] = r2[9 + i];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            V[i] = r2[12 + i];
            U[i] = r2[18 + i];
        }
        double Rb[9], pb[3];
        mat3_mul(Rw, J.Rji, Rb);
        mat3_vec(Rw, J.pji, pb);
        pb[0] += pw[0]; pb[1] += pw[1]; pb[2] += pw[2];
        double phi[6], u[6];
        xm_w2b(Rb, pb, V, phi);
        xm_w2b(Rb, pb, U, u);
        const double m = J.I[3];
        double Iw[3] = {J.I[0] * phi[0], J.I[1] * phi[1], J.I[2] * phi[2]};
        double mv[3] = {m * phi[3], m * phi[4], m * phi[5]};
        double fb[6];  // fcor + fgrav + fext (Body.m:98-107)
        cross3(Iw, phi, fb);
        cross3(mv, phi, fb + 3);
        double gw[3] = {c.gx, c.gy, c.gz}, gb[3];
        mat3T_vec(Rb, gw, gb);
        gb[0] *= m; gb[1] *= m; gb[2] *= m;
        fb[3] += gb[0]; fb[4] += gb[1]; fb[5] += gb[2];
        if (GROUND) {
            if (J.has_ground) {
                double* KD = c.KD + (size_t)t * 72;
                if (deriv) {
                    double K[36], D[36];
#pragma unroll
                    for (int i = 0; i < 36; ++i) K[i] = D[i] = 0;
                    ground_body<true>(J, Rb, pb, phi, fb, K, D);
#pragma unroll
                    for (int i = 0; i < 36; ++i) {
                        KD[i] = K[i];
                        KD[36 + i] = D[i];
                    }
                } else {
                    ground_body<false>(J, Rb, pb, phi, fb, nullptr, nullptr);
                }
            } else if (deriv) {
                double* KD = c.KD + (size_t)t * 72;
                for (int i = 0; i < 72; ++i) KD[i] = 0;
            }
        }
        double Fb[6], Fw[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) Fb[i] = J.I[i] * u[i] - c.c * fb[i];
        xf_b2w(Rb, pb, Fb, Fw);
#pragma unroll
        for (int i = 0; i < 9; ++i) r1[i] = Rb[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) r1[9 + i] = pb[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            r1[12 + i] = phi[i];
            r1[24 + i] = J.I[i];
            r2[24 + i] = Fw[i];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) r1[30 + i] = gb[i];
    }
    bsync<NW>();
    // ---- Phase D: subtree wrench sums, leaves->root (6 lanes, one component each) ---------------------------
    if (t < 6) {
        for (int j = n - 1; j > 0; --j) {
            const int par = c.jc[j].parent;
            if (par >= 0) c.rec2[(size_t)par * REC2 + 24 + t] += c.rec2[(size_t)j * REC2 + 24 + t];
        }
    }
    bsync<NW>();
    // ---- Phase E: reduced residual (thread t <-> joint t) ---------------------------------------------------
    if (t < n) {
        const JointConst& J = c.jc[t];
        if (J.idx >= 0) {
            const int r = J.idx;
            const double* r1 = c.rec1 + (size_t)t * REC1;
            const double* r2 = c.rec2 + (size_t)t * REC2;
            const double qk = c.q[r], qdk = c.qd[r];
            // Joint.computeForce (Joint.m:448-454, 470-481)
            double fr = c.tau[r] + J.stiff * (J.qRest - qk) - J.damp * qdk;
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
            c.g[r] = dot6(r1 + 18, r2 + 24) - c.c * fr;
            c.sp2[r] = dK;  // kept for the column sweep (diagonal Kr, Dr)
            c.sp1[r] = dD;
        }
    }
    bsync<NW>();
}

// Column sweep: thread t (joint i = t with a dof) computes column idx[i] of
//    sq * dg/dq + sqd * dg/dqdot + sd * dg/d(dqtmp)   into out (nr x ld, column-major), scaled by `scale`.
template <int NW, int GROUND>
__device__ void eval_columns(Ctx& c, double sq, double sqd, double sd, double scale, double* out) {
    const int t = threadIdx.x;
    const int n = c.n;
    const int ld = c.ld;
    int i = -1, col = 0, iend = 0;
    double si[6], c1[6], c2[6], Rs[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) si[k] = c1[k] = c2[k] = Rs[k] = 0;
    if (t < n && c.jc[t].idx >= 0) {
        i = t;
        const JointConst& J = c.jc[t];
        col = J.idx;
        iend = J.end;
        for (int r = 0; r < c.nr; ++r) out[(size_t)col * ld + r] = 0.0;
        const double* r1 = c.rec1 + (size_t)t * REC1;
#pragma unroll
        for (int k = 0; k < 6; ++k) si[k] = r1[18 + k];
        double Vp[6] = {0, 0, 0, 0, 0, 0}, Up[6] = {0, 0, 0, 0, 0, 0};
        if (J.parent >= 0) {
            const double* rp = c.rec2 + (size_t)J.parent * REC2;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                Vp[k] = rp[12 + k];
                Up[k] = rp[18 + k];
            }
        }
        // c1 = sqd*s_i - sq*ad(s_i)Vp ;  c2 = sd*s_i - sq*ad(s_i)Up + c*sqd*ad(Vp)s_i - c*ad(c1)Vp
        double a1[6], a2[6], a3[6], a4[6];
        ad_mv(si, Vp, a1);
        ad_mv(si, Up, a2);
        ad_mv(Vp, si, a3);
#pragma unroll
        for (int k = 0; k < 6; ++k) c1[k] = sqd * si[k] - sq * a1[k];
        ad_mv(c1, Vp, a4);
#pragma unroll
        for (int k = 0; k < 6; ++k) c2[k] = sd * si[k] - sq * a2[k] + c.c * (sqd * a3[k] - a4[k]);
    }
    const double cc = c.c;
    // descending preorder sweep over bodies; thread active while j in sub(i) = [i, iend)
    for (int j = n - 1; j >= 0; --j) {
        const double* r1 = c.rec1 + (size_t)j * REC1;
        const bool act = (i >= 0) && (j >= i) && (j < iend);
        if (act) {
            double rec[REC1];
            ld_vec(r1, rec, REC1 / 2);
            const double* Rb = rec;
            const double* pb = rec + 9;
            const double* phi = rec + 12;
            const double* sj = rec + 18;
            const double* I = rec + 24;
            const double* gb = rec + 30;
            double xi[6] = {0, 0, 0, 0, 0, 0}, dphi[6], du[6], tmp[6];
            xm_w2b(Rb, pb, c1, dphi);
            xm_w2b(Rb, pb, c2, du);
            ad_mv(dphi, phi, tmp);
#pragma unroll
            for (int k = 0; k < 6; ++k) du[k] += cc * tmp[k];
            // d(fcor) = [ (I dw) x w + (I w) x dw ;  m (dv x w + v x dw) ]
            double Iw[3] = {I[0] * phi[0], I[1] * phi[1], I[2] * phi[2]};
            double dIw[3] = {I[0] * dphi[0], I[1] * dphi[1], I[2] * dphi[2]};
            double df[6];
            cross3(dIw, phi, df);
            cross3_acc(Iw, dphi, df);
            cross3(dphi + 3, phi, df + 3);
            cross3_acc(phi + 3, dphi, df + 3);
            df[3] *= I[3]; df[4] *= I[3]; df[5] *= I[3];
            if (sq != 0.0) {
                xm_w2b(Rb, pb, si, xi);
#pragma unroll
                for (int k = 0; k < 6; ++k) xi[k] *= sq;
                // Km_grav * xi: force part += fgrav x xi_w (Body.m:119)
                cross3_acc(gb, xi, df + 3);
            }
            if (GROUND) {
                const double* KD = c.KD + (size_t)j * 72;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    double acc = 0;
#pragma unroll
                    for (int b = 0; b < 6; ++b) acc += KD[36 + 6 * a + b] * dphi[b];
                    if (sq != 0.0) {
#pragma unroll
                        for (int b = 0; b < 6; ++b) acc += KD[6 * a + b] * xi[b];
                    }
                    df[a] += acc;
                }
            }
            double dFb[6], T[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) dFb[k] = I[k] * du[k] - cc * df[k];
            xf_b2w(Rb, pb, dFb, T);
#pragma unroll
            for (int k = 0; k < 6; ++k) Rs[k] += T[k];
            const int rj = c.jc[j].idx;
            if (rj >= 0) out[(size_t)col * ld + rj] += scale * dot6(sj, Rs);
        }
        if (!c.is_chain) {
            // subtree fix-ups: joints k whose subtree ends at index j (i.e. [k, j)) must not see T of j' >= j.
            // Rs currently holds sum_{j' >= j, j' in sub(i)} T.
            const JointConst& Jj = c.jc[j];
            for (int e = 0; e < Jj.ends_cnt; ++e) {
                const int k = c.ends_list[Jj.ends_ptr + e];
                if (i >= 0 && k >= i && k < iend && j < iend) {
                    const int rk = c.jc[k].idx;
                    if (rk >= 0) out[(size_t)col * ld + rk] -= scale * dot6(c.rec1 + (size_t)k * REC1 + 18, Rs);
                }
            }
        }
    }
    if (i >= 0) {
        // ancestors k of i: s_k' (Tsub_i + sq * ad*(s_i) Fsub_i)
        double Z[6];
        if (sq != 0.0) {
            adstar_fv(si, c.rec2 + (size_t)i * REC2 + 24, Z);
#pragma unroll
            for (int k = 0; k < 6; ++k) Z[k] = Rs[k] + sq * Z[k];
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) Z[k] = Rs[k];
        }
        for (int k = c.jc[i].parent; k >= 0; k = c.jc[k].parent) {
            const int rk = c.jc[k].idx;
            if (rk >= 0) out[(size_t)col * ld + rk] = scale * dot6(c.rec1 + (size_t)k * REC1 + 18, Z);
        }
        // joint-level diagonal: -c * (sq*dK + sqd*dD)   (Kr, Dr of Joint.m:470-481)
        out[(size_t)col * ld + col] += scale * (-cc) * (sq * c.sp2[col] + sqd * c.sp1[col]);
    }
    bsync<NW>();
}

// ---------------------------------------------------------------------------------------------
// In-block LU with partial pivoting (== MATLAB H\g / lu(H,'vector'): LAPACK dgetrf pivoting, first max on ties).
// H (nr x ld column-major, shared) is overwritten by L\U; perm[r] = original row now in position r.
// Solves H dx = -g.  Thread t owns column t.
// ---------------------------------------------------------------------------------------------
template <int NW>
__device__ void lu_factor(Ctx& c, double* H, int* perm) {
    const int t = threadIdx.x;
    const int nr = c.nr, ld = c.ld;
    if (t < nr) perm[t] = t;
    bsync<NW>();
    for (int k = 0; k < nr; ++k) {
        double v = (t >= k && t < nr) ? fabs(H[(size_t)k * ld + t]) : -1.0;
        const int p = block_argmax<NW>(v, t, c.red);
        if (t < nr && p != k) {
            double a = H[(size_t)t * ld + k];
            H[(size_t)t * ld + k] = H[(size_t)t * ld + p];
            H[(size_t)t * ld + p] = a;
        }
        if (t == 0 && p != k) {
            int a = perm[k];
            perm[k] = perm[p];
            perm[p] = a;
        }
        bsync<NW>();
        const double piv = H[(size_t)k * ld + k];
        if (t > k && t < nr) H[(size_t)k * ld + t] = H[(size_t)k * ld + t] / piv;  // multipliers l_{t,k}
        bsync<NW>();
        if (t > k && t < nr) {
            double* col = H + (size_t)t * ld;
            const double* lk = H + (size_t)k * ld;
            const double ukt = col[k];
            for (int r = k + 1; r < nr; ++r) col[r] -= lk[r] * ukt;
        }
        bsync<NW>();
    }
}

// x <- solution of (P'LU) x = rhs ; thread t holds entry t.  Uses c.red-free shared vector `w` (nr doubles).
template <int NW>
__device__ void lu_solve(Ctx& c, const double* H, const int* perm, const double* rhs, double* w, double scale) {
    const int t = threadIdx.x;
    const int nr = c.nr, ld = c.ld;
    double y = 0;
    if (t < nr) y = scale * rhs[perm[t]];
    // forward: L y = P rhs   (unit lower)
    for (int k = 0; k < nr; ++k) {
        if (t == k) w[k] = y;
        bsync<NW>();
        if (t > k && t < nr) y -= H[(size_t)k * ld + t] * w[k];
    }
    // backward: U x = y
    for (int k = nr - 1; k >= 0; --k) {
        if (t == k) {
            y = y / H[(size_t)k * ld + k];
            w[k] = y;
        }
        bsync<NW>();
        if (t < k) y -= H[(size_t)k * ld + t] * w[k];
    }
    bsync<NW>();
}

// x <- solution of (P'LU)' x = rhs   (adjoint recursion, TaskBDF1.m:77: zkk0(Hp) = Hl'\(Hu'\yk))
template <int NW>
__device__ void lu_solve_T(Ctx& c, const double* H, const int* perm, const double* rhs, double* w, double* out) {
    const int t = threadIdx.x;
    const int nr = c.nr, ld = c.ld;
    double y = 0;
    if (t < nr) y = rhs[t];
    // U' v = rhs : U' is lower triangular with U'(r,k) = U(k,r) = H[r*ld + k]
    for (int k = 0; k < nr; ++k) {
        if (t == k) {
            y = y / H[(size_t)k * ld + k];
            w[k] = y;
        }
        bsync<NW>();
        if (t > k && t < nr) y -= H[(size_t)t * ld + k] * w[k];
    }
    // L' u = v : L' upper unit with L'(r,k) = L(k,r) = H[r*ld + k], k > r
    for (int k = nr - 1; k >= 0; --k) {
        if (t == k) w[k] = y;
        bsync<NW>();
        if (t < k) y -= H[(size_t)t * ld + k] * w[k];
    }
    bsync<NW>();
    if (t < nr) out[perm[t]] = y;
    bsync<NW>();
}

}  // namespace rmx
