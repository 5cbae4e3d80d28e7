// rmx_rollout.cuh -- Newton iteration, time loop and the __global__ entry points.
#pragma once
#include "rmx_device.cuh"

namespace rmx {

struct TapeArgs {
    // adjoint tape, per rollout b and step k (internal layout, consumed only by adjoint_bwd_kernel):
    //   LU   : nr*nr   column-major LU factors of H (unit-lower L below the diagonal)      [b][k][nr*nr]
    //   perm : nr      int32 (Hp of lu(H,'vector'), 0-based)                               [b][k][nr]
    //   M, D : nr*nr   row-major (i.e. transposed) so that M'z is a coalesced column sweep [b][k][nr*nr]
    //   dPdq : nr                                                                          [b][k][nr]
    double* LU;
    double* M;
    double* D;
    double* dPdq;
    int* perm;
};

struct TaskArgs {
    int body;  // internal joint index of the task body
    double xlocal[3];
    double t_target, pscale, wreg, wpos;
    const double* p;        // nr x B
    const double* xtarget;  // 3 x B
    double* P;              // B
    double* dPdp;           // nr x B
};

struct RolloutArgs {
    DevScene sc;
    StepOpts op;
    long long B;
    const double* q0;
    const double* qd0;
    const double* tau;
    double* q_out;
    double* qd_out;
    int* status;
    int* iters;
    TapeArgs tape;
    TaskArgs task;
};

// ---------------------------------------------------------------------------------------------
// newton() of driverRedMaxBDF1.m:94-157 (forward drivers: damped Newton + backtracking line search).
// Written as a two-state machine (FULL evaluation with H / residual-only line-search trial) so that the
// evaluation code has a single call site.
// ---------------------------------------------------------------------------------------------
template <int NW, bool GROUND>
__device__ __forceinline__ int newton_forward(Ctx& c, const StepOpts& op, int* perm, int& n_iter, int& n_ls) {
    const int t = threadIdx.x;
    const int nr = c.nr;
    int status = 0;
    int iter = 1;
    bool full = true;
    double f0 = 0.0, x0t = 0.0, dxt = 0.0, alpha = 1.0;
    int iterLs = 1;
    while (true) {
        eval_base<NW, GROUND>(c, full);
        const double gt = (t < nr) ? c.g[t] : 0.0;
        const double gsum = block_sum<NW>(gt * gt, c.red);
        if (full) {
            eval_columns<NW, GROUND>(c, 1.0, c.beta, 1.0, 1.0, c.H);
            f0 = 0.5 * gsum;
            // dx = -H\g
            lu_factor<NW>(c, c.H, perm);
            lu_solve<NW>(c, c.H, perm, c.g, c.dx, -1.0);
            dxt = (t < nr) ? c.dx[t] : 0.0;
            const double dxn = sqrt(block_sum<NW>(dxt * dxt, c.red));
            ++n_iter;
            if (dxn > op.dxMax) {
                status |= 1;  // 'Newton diverged': x stays at the evaluation point (driverRedMaxBDF1.m:118-121)
                break;
            }
            if (t < nr) x0t = c.q[t];
            alpha = 1.0;
            iterLs = 1;
            full = false;
        } else {
            ++n_ls;
            const double f = 0.5 * gsum;
            bool accept = f < f0;
            if (!accept && iterLs >= op.iterLsMax) {
                status |= 4;  // line search exhausted: keep the last trial (driverRedMaxBDF1.m:135-138)
                accept = true;
            }
            if (accept) {
                if (sqrt(gsum) < op.tol) break;
                if (iter >= op.iterMax) {
                    status |= 2;
                    break;
                }
                ++iter;
                full = true;
                continue;
            }
            alpha = 0.5 * alpha;
            ++iterLs;
        }
        if (t < nr) c.q[t] = __dadd_rn(x0t, __dmul_rn(alpha, dxt));
        bsync<NW>();
    }
    return status;
}

// ---------------------------------------------------------------------------------------------
// Forward rollout kernel: simLoop of driverRedMaxBDF1.m:57-91 / driverRedMaxBDF2.m:57-125, one block per rollout.
// ---------------------------------------------------------------------------------------------
template <int NW, bool GROUND>
__global__ void __launch_bounds__(32 * NW) rollout_fwd_kernel(RolloutArgs a) {
    extern __shared__ double2 smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    __shared__ int perm_s[32 * NW];
    const int t = threadIdx.x;
    const int n = a.sc.n, nr = a.sc.nr;
    Ctx c;
    ctx_carve(c, sm, n, nr, GROUND);
    c.jc = a.sc.jc;
    c.ends_list = a.sc.ends_list;
    c.gx = a.sc.grav[0];
    c.gy = a.sc.grav[1];
    c.gz = a.sc.grav[2];
    c.is_chain = a.sc.is_chain;
    const StepOpts op = a.op;
    const double h = op.h;
    const double ah = __dmul_rn(SDIRK_A_CONST, h);
    const double bh = __dmul_rn(__dsub_rn(1.0, SDIRK_A_CONST), h);

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        double qc = 0.0, qdc = 0.0;  // current state of dof t (joint.q / joint.qdot)
        double q0t = 0.0, qd0t = 0.0, q1t = 0.0, qdat = 0.0;
        if (t < nr) {
            qc = a.q0[b * nr + t];
            qdc = a.qd0[b * nr + t];
            c.tau[t] = (op.tau_mode == 1) ? a.tau[b * nr + t] : 0.0;
            c.hq1[t] = qc;
            c.hqd1[t] = qdc;
        }
        int status = 0, n_iter = 0, n_ls = 0;
        for (int k = 0; k < op.nsteps; ++k) {
            if (op.tau_mode == 2 && t < nr) c.tau[t] = a.tau[((size_t)b * op.nsteps + k) * nr + t];
            const int nsub = (op.scheme == 2 && k == 0) ? 2 : 1;
            for (int sub = 0; sub < nsub; ++sub) {
                const int stage = (op.scheme == 1) ? ST_BDF1 : (k == 0 ? (sub == 0 ? ST_SDIRK_A : ST_SDIRK_B) : ST_BDF2);
                // ---- save old state + initial guess -------------------------------------------------------
                if (t < nr) {
                    if (stage == ST_BDF1) {  // driverRedMaxBDF1.m:64-70
                        q0t = qc;
                        c.hq0[t] = qc;
                        c.hqd0[t] = qdc;
                        c.q[t] = __dadd_rn(qc, __dmul_rn(h, qdc));
                    } else if (stage == ST_SDIRK_A) {  // driverRedMaxBDF2.m:67-76
                        q0t = qc;
                        qd0t = qdc;
                        c.hq0[t] = qc;
                        c.hqd0[t] = qdc;
                        c.q[t] = __dadd_rn(qc, __dmul_rn(ah, qdc));
                    } else if (stage == ST_SDIRK_B) {  // :81-84 (qa, qdota live in the q1 slots)
                        c.hq1[t] = qc;
                        c.hqd1[t] = qdat;
                        c.q[t] = __dadd_rn(qc, __dmul_rn(bh, qdat));
                    } else {  // BDF2, :97-103: Q1->Q0, Q->Q1
                        q0t = c.hq1[t];
                        q1t = qc;
                        c.hq0[t] = q0t;
                        c.hqd0[t] = c.hqd1[t];
                        c.hq1[t] = qc;
                        c.hqd1[t] = qdc;
                        c.q[t] = __dadd_rn(qc, __dmul_rn(h, qdc));
                    }
                }
                stage_coef(c, stage, h);
                bsync<NW>();
                status |= newton_forward<NW, GROUND>(c, op, perm_s, n_iter, n_ls);
                // ---- new state ----------------------------------------------------------------------------
                if (t < nr) {
                    const double x = c.q[t];
                    if (stage == ST_BDF1) {  // :71
                        qdc = __ddiv_rn(__dsub_rn(x, q0t), h);
                    } else if (stage == ST_SDIRK_A) {  // :78 ; qc temporarily holds qa
                        qdat = __ddiv_rn(__dsub_rn(x, q0t), ah);
                    } else if (stage == ST_SDIRK_B) {  // :86-91
                        qdc = __ddiv_rn(__dsub_rn(__dsub_rn(x, q0t), __dmul_rn(bh, qdat)), ah);
                        c.hq1[t] = q0t;  // jroot.setQ1(q0,qdot0)
                        c.hqd1[t] = qd0t;
                    } else {  // :105
                        const double e = __dadd_rn(__dsub_rn(x, __dmul_rn(4.0 / 3.0, q1t)), __dmul_rn(1.0 / 3.0, q0t));
                        qdc = __dmul_rn(__ddiv_rn(3.0, __dmul_rn(2.0, h)), e);
                    }
                    qc = x;
                }
            }
            if (t < nr) {
                const size_t o = ((size_t)b * op.nsteps + k) * nr + t;
                a.q_out[o] = qc;
                if (a.qd_out) a.qd_out[o] = qdc;
            }
            bsync<NW>();
        }
        double bad = (t < nr && !(isfinite(qc) && isfinite(qdc))) ? 1.0 : 0.0;
        bad = block_sum<NW>(bad, c.red);
        if (bad > 0.0) status |= 8;
        if (t == 0) {
            a.status[b] = status;
            if (a.iters) {
                a.iters[2 * b] = n_iter;
                a.iters[2 * b + 1] = n_ls;
            }
        }
        bsync<NW>();
    }
}

// ---------------------------------------------------------------------------------------------
// Test hook: one evaluation (B = 1) -> g, H, M, D in global memory (nr x nr column-major, dense ld = nr)
// ---------------------------------------------------------------------------------------------
struct EvalArgs {
    DevScene sc;
    const double* q;
    const double* qd;
    const double* dq;
    const double* tau;
    double cK, beta;
    double* g;
    double* H;
    double* M;
    double* D;
};

template <int NW, bool GROUND>
__global__ void __launch_bounds__(32 * NW) eval_kernel(EvalArgs a) {
    extern __shared__ double2 smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const int t = threadIdx.x;
    const int n = a.sc.n, nr = a.sc.nr;
    Ctx c;
    ctx_carve(c, sm, n, nr, GROUND);
    c.jc = a.sc.jc;
    c.ends_list = a.sc.ends_list;
    c.gx = a.sc.grav[0];
    c.gy = a.sc.grav[1];
    c.gz = a.sc.grav[2];
    c.is_chain = a.sc.is_chain;
    c.stage = ST_DIRECT;  // direct: qd = hqd0, dq = hq1
    c.h = 1.0;
    c.c = a.cK;
    c.beta = a.beta;
    if (t < nr) {
        c.q[t] = a.q[t];
        c.hqd0[t] = a.qd[t];
        c.hq1[t] = a.dq[t];
        c.hq0[t] = 0;
        c.hqd1[t] = 0;
        c.tau[t] = a.tau ? a.tau[t] : 0.0;
    }
    bsync<NW>();
    eval_base<NW, GROUND>(c, true);
    if (t < nr && a.g) a.g[t] = c.g[t];
    const int ld = c.ld;
    for (int pass = 0; pass < 3; ++pass) {
        double* dst = pass == 0 ? a.H : (pass == 1 ? a.M : a.D);
        if (!dst) continue;
        if (pass == 0) eval_columns<NW, GROUND>(c, 1.0, c.beta, 1.0, 1.0, c.H);
        if (pass == 1) eval_columns<NW, GROUND>(c, 0.0, 0.0, 1.0, 1.0, c.H);
        if (pass == 2) eval_columns<NW, GROUND>(c, 0.0, 1.0, 0.0, -1.0 / c.c, c.H);
        for (int e = t; e < nr * nr; e += blockDim.x) {
            const int col = e / nr, row = e % nr;
            dst[e] = c.H[(size_t)col * ld + row];
        }
        bsync<NW>();
    }
}

}  // namespace rmx
