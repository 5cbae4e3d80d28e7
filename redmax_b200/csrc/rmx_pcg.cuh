// rmx_pcg.cuh -- alternative Newton linear solve: preconditioned Krylov iteration with the projected block-Jacobi
// preconditioner of the reference's c++/PCG solver.
//
// Template (reference, c++/PCG/src/): Solver::pcdSaad2003 (Solver.cpp:81-167, Saad Alg. 9.1, rel. tolerance 1e-6 on the
// residual, Solver.h:43) with the preconditioner of ConstraintJoint::preprocess_PCG_preconditioner (ConstraintJoint.cpp:1236:
// Mhat, Psi, Pi) applied by ConstraintJoint::computeMinv_x (:1455, notes.pdf Alg. 10): the exact O(n) inverse of
// J' blkdiag(Mhat_j) J + Pr obtained by an articulated-body recursion.
//
// What differs, and why:
//   * the reference integrates with linearly-implicit Euler, whose system matrix is SPD, and runs plain PCG on it.  The Newton
//     matrix of the fully implicit BDF steps, H = M - cD D - cK K + dM/dq dqtmp (driverRedMaxBDF1.m:182-184), is NOT symmetric,
//     so the Krylov method here is BiCGStab with the same preconditioner (SURVEY.md hard part H2);
//   * the recursion is written in the world frame (world-frame screws s_k and body inertias): no adjoint transforms between
//     frames appear, children simply add into their parent:
//        leaves->root : IA_j = I_j + sum_c (IA_c - U_c U_c'/d_c),   U_c = IA_c s_c,   d_c = s_c'U_c + Pr_c     (= Pi, Psi^-1)
//        leaves->root : u_j = x_j - s_j'p_j ;   p_parent += p_j + U_j u_j / d_j                                 (Bhat, beta)
//        root->leaves : y_j = (u_j - U_j'a_parent) / d_j ;   a_j = a_parent + s_j y_j                            (Vdot, y)
//     Mhat_j is the body inertia, Pr the joint-space diagonal -c (dfr/dq + beta dfr/dqdot) (joint stiffness, damping, limits);
//     tools/proto_precond.py checks the recursion against a dense solve.
//   * the operator is applied MATRIX-FREE, as the reference applies J x / LHS / J' y (ConstraintJoint.cpp:1090, 1137, 1188): the
//     Newton matrix is tree-semiseparable, H[k][i] = L_k . Rt_i for k in sub(i) and s_k . Z_i for k a proper ancestor of i
//     (rmx_fast.cuh), so with the per-joint vectors (O(n) to form, no n x n matrix is ever assembled)
//        (H x)_k = L_k . P_k + s_k . Q_k + dg_k x_k ,   P_k = sum_{i in anc*(k)} Rt_i x_i   (root -> leaves sweep)
//                                                      Q_k = sum_{i in sub(k), i != k} Z_i x_i  (leaves -> root sweep)
//     the first sweep a pointer-jumping prefix sum, the second the subtree accumulation the evaluation uses for its wrenches.
//     Scenes with forces between body points (off-tree couplings) keep the dense matrix and a dense product.
// Measured: for nr <= 64 the in-block LU is faster than this solve; LU stays the default and the parity path.
#pragma once
#include "rmx_fast.cuh"

namespace rmx {

struct PcgMem {
    double* IA;    // [n][21] articulated inertias (symmetric 6x6, upper triangle row-major), joint-major
    double* U;     // [n][6]
    double* P;     // [n][6]  bias forces / accelerations (reused)
    double* dinv;  // [n]
    double* u;     // [n]
    double* vec;   // 8 nr-vectors: r, r0, p, v, s, t, phat, shat
    double* WR;    // [n][2 * 24]  per joint [L_k ; s_k | Rt_k ; Z_k]  (matrix-free operator)
    double* PS;    // [n][18]      prefix sums of Rt_i x_i along the root paths
    double* QS;    // [n][6]       subtree sums of Z_i x_i
    double* dgv;   // [nr]         joint-level diagonal of H
};
constexpr int PCG_WR = 48, PCG_PS = 18;

__host__ __device__ inline size_t pcg_doubles(int n, int nr) {
    return (size_t)n * (21 + 6 + 6 + 2 + PCG_WR + PCG_PS + 6) + 9 * (size_t)nr + 2;
}

__device__ __forceinline__ void pcg_carve(PcgMem& m, double* p, int n, int nr) {
    m.IA = p; p += (size_t)n * 21;
    m.U = p; p += (size_t)n * 6;
    m.P = p; p += (size_t)n * 6;
    m.dinv = p; p += n;
    m.u = p; p += n;
    m.vec = p; p += 8 * (size_t)nr;
    m.WR = p; p += (size_t)n * PCG_WR;
    m.PS = p; p += (size_t)n * PCG_PS;
    m.QS = p; p += (size_t)n * 6;
    m.dgv = p;
}

__device__ __forceinline__ int sym_idx(int a, int b) {  // a <= b
    return a * 6 - (a * (a - 1)) / 2 + (b - a);
}

// Preconditioner set-up for the current evaluation point (needs eval_base2's S, RB, PB fields: KEEP layout).
template <int NW, int GROUND>
__device__ void precond_setup(Ctx2& c, PcgMem& m) {
    typedef Fld<GROUND, true> F;
    const int t = threadIdx.x;
    const int n = c.n, NS = c.NS;
    if (t < n) {
        const JointConst& J = c.jc[t];
        double R[9], p[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = SA(F::RB, i, t);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = SA(F::PB, i, t);
        const double ms = J.I[3];
        double* ia = m.IA + (size_t)t * 21;
        const double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        // [[R I3 R' - m[p][p], m[p]], [-m[p], m 1]]
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = a; b < 3; ++b) {
                double v = R[3 * a] * J.I[0] * R[3 * b] + R[3 * a + 1] * J.I[1] * R[3 * b + 1] + R[3 * a + 2] * J.I[2] * R[3 * b + 2];
                v -= ms * (p[a] * p[b] - (a == b ? pp : 0.0));
                ia[sym_idx(a, b)] = v;
            }
        // m[p] : rows 0..2, cols 3..5
        ia[sym_idx(0, 3)] = 0.0;          ia[sym_idx(0, 4)] = -ms * p[2];   ia[sym_idx(0, 5)] = ms * p[1];
        ia[sym_idx(1, 3)] = ms * p[2];    ia[sym_idx(1, 4)] = 0.0;          ia[sym_idx(1, 5)] = -ms * p[0];
        ia[sym_idx(2, 3)] = -ms * p[1];   ia[sym_idx(2, 4)] = ms * p[0];    ia[sym_idx(2, 5)] = 0.0;
        ia[sym_idx(3, 3)] = ms; ia[sym_idx(3, 4)] = 0.0; ia[sym_idx(3, 5)] = 0.0;
        ia[sym_idx(4, 4)] = ms; ia[sym_idx(4, 5)] = 0.0;
        ia[sym_idx(5, 5)] = ms;
    }
    bsync<NW>();
    if (t < 32) {
        const int e = t;
        int ea = 0, eb = 0;  // (row, col) of symmetric entry e
        {
            int k = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) {
                    if (k == e) {
                        ea = a;
                        eb = b;
                    }
                    ++k;
                }
        }
        for (int j = n - 1; j >= 0; --j) {
            const int idx = c.ie_s[j].x, par = c.par_s[j];
            double* ia = m.IA + (size_t)j * 21;
            double val = (e < 21) ? ia[e] : 0.0;
            if (idx >= 0) {
                double s[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) s[i] = SA(F::S, i, j);
                if (e < 6) {
                    double acc = 0.0;
#pragma unroll
                    for (int b = 0; b < 6; ++b) acc += ia[e <= b ? sym_idx(e, b) : sym_idx(b, e)] * s[b];
                    m.U[(size_t)j * 6 + e] = acc;
                }
                __syncwarp();
                double d = -c.c * (c.sp2[idx] + c.beta * c.sp1[idx]);  // Pr
#pragma unroll
                for (int i = 0; i < 6; ++i) d += s[i] * m.U[(size_t)j * 6 + i];
                const double di = 1.0 / d;
                if (e == 0) m.dinv[j] = di;
                if (e < 21) val -= m.U[(size_t)j * 6 + ea] * m.U[(size_t)j * 6 + eb] * di;
            }
            if (par >= 0 && e < 21) m.IA[(size_t)par * 21 + e] += val;
            __syncwarp();
        }
    }
    bsync<NW>();
}

// y = (J' blkdiag(M_j) J + Pr)^-1 x    (x, y indexed by reduced index)
template <int NW, int GROUND>
__device__ void precond_apply(Ctx2& c, PcgMem& m, const double* x, double* y) {
    typedef Fld<GROUND, true> F;
    const int t = threadIdx.x;
    const int n = c.n, NS = c.NS;
    if (t < 32) {
        const int e = t < 6 ? t : 0;
        for (int j = t; j < 6 * n; j += 32) m.P[j] = 0.0;
        __syncwarp();
        for (int j = n - 1; j >= 0; --j) {
            const int idx = c.ie_s[j].x, par = c.par_s[j];
            double pj = m.P[(size_t)j * 6 + e];
            if (idx >= 0) {
                double uj = x[idx];
#pragma unroll
                for (int i = 0; i < 6; ++i) uj -= SA(F::S, i, j) * m.P[(size_t)j * 6 + i];
                if (t == 0) m.u[j] = uj;
                pj += m.U[(size_t)j * 6 + e] * (uj * m.dinv[j]);
            }
            if (par >= 0 && t < 6) m.P[(size_t)par * 6 + e] += pj;
            __syncwarp();
        }
        // accelerations reuse P (the bias forces are no longer needed once u is known)
        for (int j = 0; j < n; ++j) {
            const int idx = c.ie_s[j].x, par = c.par_s[j];
            double ap[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) ap[i] = (par >= 0) ? m.P[(size_t)par * 6 + i] : 0.0;
            double aj = ap[e];
            if (idx >= 0) {
                double yj = m.u[j];
#pragma unroll
                for (int i = 0; i < 6; ++i) yj -= m.U[(size_t)j * 6 + i] * ap[i];
                yj *= m.dinv[j];
                if (t == 0) y[idx] = yj;
                aj += SA(F::S, e, j) * yj;
            }
            __syncwarp();
            if (t < 6) m.P[(size_t)j * 6 + e] = aj;
            __syncwarp();
        }
    }
    bsync<NW>();
}

// Matrix-free form of sq dg/dq + sqd dg/dqdot + sd dg/d(dqtmp): the per-joint vectors only (columns_joint), kept in shared memory.
template <int NW, int GROUND>
__device__ void eval_columns_mf(Ctx2& c, PcgMem& m, double sq, double sqd, double sd) {
    typedef Fld<GROUND, 1> F;
    constexpr int NL = F::NL;
    const int t = threadIdx.x;
    const int n = c.n;
    const int myidx = (t < n) ? c.ie_s[t].x : -1;
    double Rt[NL], Z[6], L[NL], s[6];
    columns_joint<NW, GROUND, 1>(c, t, myidx, sq, sqd, sd, L, s, Rt, Z);
    if (t < n) {
        double* w = m.WR + (size_t)t * PCG_WR;
        const bool dof = myidx >= 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            w[i] = dof ? L[i] : 0.0;
            w[24 + i] = dof ? Rt[i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            w[NL + i] = dof ? s[i] : 0.0;
            w[24 + NL + i] = dof ? Z[i] : 0.0;
        }
        if (dof) m.dgv[myidx] = -c.c * (sq * c.sp2[myidx] + sqd * c.sp1[myidx]);  // Kr, Dr of Joint.m:470-481
    }
    bsync<NW>();
}

// y = H x through the two tree sweeps (x, y indexed by reduced index).
template <int NW, int GROUND>
__device__ void hx_apply(Ctx2& c, PcgMem& m, const double* x, double* y) {
    typedef Fld<GROUND, 1> F;
    constexpr int NL = F::NL;
    const int t = threadIdx.x;
    const int n = c.n;
    const int myidx = (t < n) ? c.ie_s[t].x : -1;
    const double* w = m.WR + (size_t)(t < n ? t : 0) * PCG_WR;
    const double xi = (myidx >= 0) ? x[myidx] : 0.0;
    double P[NL];
    if (t < n) {
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            P[i] = w[24 + i] * xi;
            m.PS[(size_t)t * PCG_PS + i] = P[i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) m.QS[(size_t)t * 6 + i] = w[24 + NL + i] * xi;
    }
    bsync<NW>();
    // root -> leaves (computeJ_x's direction): P_k = sum over anc*(k), pointer jumping
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        if (r >= c.nrounds) break;  // uniform
        const int a = (t < n) ? c.anc_r[r] : -1;
        if (a >= 0) {
#pragma unroll
            for (int i = 0; i < NL; ++i) P[i] += m.PS[(size_t)a * PCG_PS + i];
        }
        bsync<NW>();
        if (a >= 0) {
#pragma unroll
            for (int i = 0; i < NL; ++i) m.PS[(size_t)t * PCG_PS + i] = P[i];
        }
        bsync<NW>();
    }
    // leaves -> root (computeJT_x's direction): subtree sums, one thread per component
    if (t < 6) {
        if (c.is_chain) {
            double acc = 0.0;
            for (int j = n - 1; j >= 0; --j) {
                acc += m.QS[(size_t)j * 6 + t];
                m.QS[(size_t)j * 6 + t] = acc;
            }
        } else {
            for (int j = n - 1; j > 0; --j) {
                const int par = c.par_s[j];
                if (par >= 0) m.QS[(size_t)par * 6 + t] += m.QS[(size_t)j * 6 + t];
            }
        }
    }
    bsync<NW>();
    if (myidx >= 0) {
        double acc = m.dgv[myidx] * xi;
#pragma unroll
        for (int i = 0; i < NL; ++i) acc = fma(w[i], P[i], acc);
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fma(w[NL + i], m.QS[(size_t)t * 6 + i] - w[24 + NL + i] * xi, acc);
        y[myidx] = acc;
    }
    bsync<NW>();
}

// Solves H x = scale * rhs with preconditioned BiCGStab; x -> c.dx.  Returns the number of iterations.
// MF: the operator is hx_apply (per-joint vectors from eval_columns_mf); otherwise the dense matrix H in shared memory.
template <int NW, int GROUND, bool MF>
__device__ int krylov_solve(Ctx2& c, PcgMem& m, const double* H, const double* rhs, double scale, double tol, int maxit) {
    const int t = threadIdx.x;
    const int nr = c.nr, ld = c.ld;
    double* r = m.vec;
    double* r0 = r + nr;
    double* p = r0 + nr;
    double* v = p + nr;
    double* s = v + nr;
    double* tt = s + nr;
    double* ph = tt + nr;
    double* sh = ph + nr;
    precond_setup<NW, GROUND>(c, m);
    double xt = 0.0, rt = 0.0, r0t = 0.0, pt = 0.0, vt = 0.0;
    if (t < nr) {
        rt = scale * rhs[t];
        r0t = rt;
    }
    const double bnorm2 = block_sum<NW>(rt * rt, c.red);
    const double thr2 = tol * tol * bnorm2;  // stop when ||r||^2 < tol^2 ||r0||^2 (Solver.cpp:137)
    double rho = 1.0, alpha = 1.0, omega = 1.0;
    int it = 0;
    if (bnorm2 > 0.0) {
        while (it < maxit) {
            ++it;
            const double rho_new = block_sum<NW>(r0t * rt, c.red);
            if (rho_new == 0.0) break;
            const double beta = (rho_new / rho) * (alpha / omega);
            pt = rt + beta * (pt - omega * vt);
            if (t < nr) p[t] = pt;
            bsync<NW>();
            precond_apply<NW, GROUND>(c, m, p, ph);
            vt = 0.0;
            if (MF) {
                hx_apply<NW, GROUND>(c, m, ph, v);
                if (t < nr) vt = v[t];
            } else if (t < nr) {
                for (int k = 0; k < nr; ++k) vt = fma(H[(size_t)k * ld + t], ph[k], vt);
            }
            alpha = rho_new / block_sum<NW>(r0t * vt, c.red);
            const double st = rt - alpha * vt;
            const double pht = (t < nr) ? ph[t] : 0.0;
            const double sn2 = block_sum<NW>(st * st, c.red);
            if (sn2 < thr2) {
                xt += alpha * pht;
                break;
            }
            if (t < nr) s[t] = st;
            bsync<NW>();
            precond_apply<NW, GROUND>(c, m, s, sh);
            double tv = 0.0;
            if (MF) {
                hx_apply<NW, GROUND>(c, m, sh, tt);
                if (t < nr) tv = tt[t];
            } else if (t < nr) {
                for (int k = 0; k < nr; ++k) tv = fma(H[(size_t)k * ld + t], sh[k], tv);
            }
            const double ts = block_sum<NW>(tv * st, c.red);
            const double t2 = block_sum<NW>(tv * tv, c.red);
            omega = (t2 > 0.0) ? ts / t2 : 0.0;
            const double sht = (t < nr) ? sh[t] : 0.0;
            xt += alpha * pht + omega * sht;
            rt = st - omega * tv;
            const double rn2 = block_sum<NW>(rt * rt, c.red);
            if (rn2 < thr2 || omega == 0.0) break;
            rho = rho_new;
        }
    }
    (void)r;
    (void)r0;
    (void)ld;
    if (t < nr) c.dx[t] = xt;
    bsync<NW>();
    return it;
}

}  // namespace rmx
