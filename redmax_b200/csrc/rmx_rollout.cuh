// rmx_rollout.cuh -- Newton iteration, time loop and the __global__ entry points.
#pragma once
#include "rmx_device.cuh"
#include "rmx_fast.cuh"
#include "rmx_tc.cuh"
#include "rmx_pcg.cuh"

namespace rmx {

struct TapeArgs {
    // adjoint tape, per rollout b and step k (internal layout, consumed only by adjoint_bwd_kernel).
    // What Scene.saveHistory keeps for Task*.calcFinal (Scene.m:141-148; only Hl,Hu,Hp,M,D are consumed):
    //   A[b][k][sza] : one 16B-aligned record  { LU[nr*ld] | dPdq[nr] | perm int32[nr] (+pad) }
    //                  LU = the shared-memory image of the factored H (column-major, ld = nr|1, unit-lower L
    //                  below the diagonal) so the backward kernel can bulk-copy (TMA) it straight into smem;
    //                  perm = Hp of lu(H,'vector') (0-based)
    //   M[b][k][nr*nr], D[b][k][nr*nr] : row-major, so thread i streams column i of M' with coalesced loads
    double* A;
    double* M;
    double* D;
    int sza;  // doubles per A record
};
__host__ __device__ inline int tape_sza(int nr) {
    int d = nr * h_ld(nr) + nr + (nr + 1) / 2;
    return (d + 1) & ~1;
}

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
    unsigned long long* kry_total;  // total Krylov iterations of the launch (LIN == 1) or null
    // Segment schedule (forward kernels, optional): block s runs segments seg[seg_off[s] .. seg_off[s+1]), each
    // {rollout, first step, end step, bit0: wait for flags[rollout] / bit1: set it when done}.  A rollout is cut at most once,
    // its first part at the head of one block's list, its second part at the tail of the previous block's list (McNaughton's
    // wrap-around rule), so all blocks finish together instead of leaving a partial last wave.  Null: one block per rollout.
    const int4* seg;
    const int* seg_off;
    int* flags;
    // cursor[s]: how many segments of list s have been claimed (zeroed before the launch).  A block works through its own list
    // and, once that is empty, takes unclaimed segments of other lists (claim_segment): the segments queued behind a stalled
    // rollout -- one that runs 10 nr iterations x 20 halvings in some step -- are integrated by blocks that are out of work
    // instead of waiting for it.  Who integrates a segment does not change its arithmetic: results stay bitwise the same.
    int* cursor;
    int nlists;  // number of segment lists (co-resident slots of a load-balanced launch; B for rmx_rollout_resume)
    // Optional mirrors of q_out / qd_out in mapped pinned HOST memory (device pointers of the caller's buffers): every step
    // is stored to both, so the trajectories cross PCIe while the rollout is still running and the host-pointer entry needs
    // no device-to-host copy afterwards.  The device copies stay the ones a cut rollout resumes from.  Null: not mirrored.
    double* q_host;
    double* qd_host;
    // Lockstep groups (one-warp forward kernels; group_barrier in rmx_device.cuh): `group` warps per block, each an independent
    // rollout slot with `group_stride` doubles of the block's dynamic shared memory.  group <= 1: one rollout per block.
    int group;
    int group_stride;
};
#ifndef RMX_MAX_GROUP
#define RMX_MAX_GROUP 8
#endif

// Shared-memory scratch of the point forces sits behind everything any kernel variant places after the Newton matrix
// (task Jacobian rows of the adjoint kernels, Krylov vectors), at the same offset for all of them.
__host__ __device__ inline size_t pf_offset_doubles(int n, int nr, bool ground, bool keep) {
    return smem_doubles2(n, nr, ground, keep) + 6 * (size_t)nr + pcg_doubles(n, nr);
}

__device__ __forceinline__ int krylov_count(const Ctx&) { return 0; }
__device__ __forceinline__ void krylov_reset(Ctx&) {}

// ---------------------------------------------------------------------------------------------
// Two implementations of the evaluation behind one interface:
//   IMPL 1 : rmx_device.cuh  -- serial tree sweeps + per-(column, body) tangent sweep (kept for n > 64 and as cross-check)
//   IMPL 2 : rmx_fast.cuh    -- scans + composite blocks + (one warp) register LU
// ---------------------------------------------------------------------------------------------
template <int IMPL, int NW, int GROUND, bool KEEP, int LIN, bool ATC = false>
struct Eval;

struct Ctx2L : Ctx2 {  // fast-path context + state of the Krylov linear solve (LIN == 1)
    PcgMem pm;
    double lin_tol;
    int lin_maxit;
    int kry_iters;
};

template <int NW, int GROUND, bool KEEP, int LIN, bool ATC>
struct Eval<1, NW, GROUND, KEEP, LIN, ATC> {
    typedef Ctx C;
    static constexpr bool TCA = false;
    static __device__ __forceinline__ double* jrows(const C& c) { return c.H + (size_t)c.nr * c.ld; }
    static __device__ __forceinline__ size_t extra_off(const C& c) { return (size_t)c.nr * c.ld; }
    static __device__ __forceinline__ void setup(C& c, double* sm, const DevScene& sc, const StepOpts&) {
        ctx_carve(c, sm, sc.n, sc.nr, GROUND);
        c.pf = nullptr;  // the sweep kernels take no point forces (rmx_scene_create rejects the combination)
        c.pf_ep = nullptr;
        c.npf = 0;
        c.pf_s = nullptr;
        c.jc = sc.jc;
        c.ends_list = sc.ends_list;
        c.gx = sc.grav[0];
        c.gy = sc.grav[1];
        c.gz = sc.grav[2];
        c.is_chain = sc.is_chain;
    }
    static __device__ __forceinline__ void base(C& c, bool deriv) { eval_base<NW, GROUND>(c, deriv); }
    static __device__ __forceinline__ void columns(C& c, double sq, double sqd, double sd, double scale, double* out) {
        eval_columns<NW, GROUND>(c, sq, sqd, sd, scale, out);
    }
    // dx = scale * H \ rhs ; H is overwritten by its LU image, perm = row permutation
    static __device__ __forceinline__ void factor_solve(C& c, int* perm, double scale, bool /*write_back*/) {
        lu_factor<NW>(c, c.H, perm);
        lu_solve<NW>(c, c.H, perm, c.g, c.dx, scale);
    }
    static __device__ __forceinline__ void body_frame(const C& c, int j, double* R, double* p) {
        const double* r1 = c.rec1 + (size_t)j * REC1;
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = r1[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = r1[9 + i];
    }
    static __device__ __forceinline__ void screw(const C& c, int j, double* s) {
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = c.rec1[(size_t)j * REC1 + 18 + i];
    }
    static __device__ __forceinline__ void body_phi(const C& c, int j, double* phi) {
#pragma unroll
        for (int i = 0; i < 6; ++i) phi[i] = c.rec1[(size_t)j * REC1 + 12 + i];
    }
};

template <int NW, int GROUND, bool KEEP, int LIN, bool ATC>
struct Eval<2, NW, GROUND, KEEP, LIN, ATC> {
    typedef Ctx2L C;
    // one- and two-warp forward kernels assemble and factor the Newton matrix on the FP64 tensor cores (rmx_tc.cuh); so does
    // the one-warp adjoint forward kernel of scenes without external forces (TCA, layout TcLayoutA)
    static constexpr bool TC = (NW <= 2 && !KEEP && LIN == 0);
    static constexpr bool TCA = (ATC && NW == 1 && !GROUND && KEEP && LIN == 0);
    static constexpr int KF = TCA ? 2 : (KEEP ? 1 : 0);  // which fields eval_base2 keeps (Fld<GROUND, KF>)
    // Krylov solve: matrix-free operator unless forces between body points couple bodies off the tree (rmx_pcg.cuh)
    static constexpr bool MF = (LIN == 1 && GROUND < 2);
    typedef Fld<GROUND, KF> F;
    static __device__ __forceinline__ double* jrows(const C& c) { return TCA ? c.jrows : c.H + (size_t)c.nr * c.ld; }
    static __device__ __forceinline__ void setup(C& c, double* sm, const DevScene& sc, const StepOpts& op) {
        if (TC)
            ctx2_carve_tc<GROUND, NW>(c, sm, sc.n, sc.nr);
        else if (TCA)
            ctx2_carve_tca(c, sm, sc.n, sc.nr);
        else
            ctx2_carve(c, sm, sc.n, sc.nr, GROUND, KEEP);
        c.lin_tol = op.lin_tol;
        c.lin_maxit = op.lin_maxit;
        c.kry_iters = 0;
        if (LIN == 1) pcg_carve(c.pm, c.H + (size_t)c.nr * c.ld, sc.n, sc.nr);
        c.jc = sc.jc;
        c.ends_list = sc.ends_list;
        c.gx = sc.grav[0];
        c.gy = sc.grav[1];
        c.gz = sc.grav[2];
        c.is_chain = sc.is_chain;
        c.anc = sc.anc;
        c.nrounds = sc.nrounds;
#pragma unroll
        for (int r = 0; r < 6; ++r) c.anc_r[r] = (r < sc.nrounds && (int)threadIdx.x < sc.n) ? __ldg(sc.anc + r * sc.n + threadIdx.x) : -1;
        c.pf = sc.pf;
        c.pf_ep = sc.pf_ep;
        c.npf = sc.npf;
        c.pf_s = sm + pf_offset_doubles(sc.n, sc.nr, GROUND, KEEP);
        for (int j = threadIdx.x; j < sc.n; j += 32 * NW) {
            c.ie_s[j] = make_int2(sc.jc[j].idx, sc.jc[j].end);
            c.par_s[j] = sc.jc[j].parent;
        }
        bsync<NW>();
        if (TC || TCA) {  // tree relation as bit masks (thread = joint k): the tile epilogue tests one bit per matrix entry
            typedef typename TcMask<NW>::type mask_t;
            const int k = threadIdx.x;
            mask_t sub = 0, anc = 0;
            int idx = -1;
            if (k < sc.n) {
                idx = c.ie_s[k].x;
                const int endk = c.ie_s[k].y;
                for (int i = 0; i < sc.n; ++i) {
                    if (i <= k && k < c.ie_s[i].y) sub |= (mask_t)1 << i;
                    if (k < i && i < endk) anc |= (mask_t)1 << i;
                }
            }
            c.tcidx_s[k] = idx;
            reinterpret_cast<mask_t*>(c.tcsub_s)[k] = sub;
            reinterpret_cast<mask_t*>(c.tcanc_s)[k] = anc;
            bsync<NW>();
        }
    }
    static __device__ __forceinline__ void base(C& c, bool deriv) { eval_base2<NW, GROUND, KF>(c, deriv); }
    static __device__ __forceinline__ void columns(C& c, double sq, double sqd, double sd, double scale, double* out) {
        if (TC)
            eval_columns_tc<NW, GROUND, TcLayout<GROUND, NW>, false>(c, sq, sqd, sd, scale, out);
        else if (TCA)
            eval_columns_tc<1, false, TcLayoutA, false>(c, sq, sqd, sd, scale, out);
        else if (MF)
            eval_columns_mf<NW, GROUND>(c, c.pm, sq, sqd, sd);  // (scale is 1 for the Newton matrix)
        else
            eval_columns2<NW, GROUND, KF>(c, sq, sqd, sd, scale, out);
    }
    // adjoint tape: scale * (...) as nr x nr row-major straight to global memory (TCA only)
    static __device__ __forceinline__ void columns_global(C& c, double sq, double sqd, double sd, double scale, double* out) {
        eval_columns_tc<1, false, TcLayoutA, true>(c, sq, sqd, sd, scale, out);
    }
    static __device__ __forceinline__ void factor_solve(C& c, int* perm, double scale, bool write_back) {
        if (TC || TCA) {
            lu_solve_tc<NW>(c.nr, c.H, perm, c.rem_s, c.tcrow_s, c.g, scale, c.dx);
        } else if (LIN == 1) {
            c.kry_iters += krylov_solve<NW, GROUND, MF>(c, c.pm, c.H, c.g, scale, c.lin_tol, c.lin_maxit);
        } else if (NW == 1) {
            lu_solve_warp(c.nr, c.ld, c.H, perm, c.g, scale, c.dx, write_back, c.lubuf);
        } else {
            lu_factor<NW>(c, c.H, perm);
            lu_solve<NW>(c, c.H, perm, c.g, c.dx, scale);
        }
    }
    static __device__ __forceinline__ void body_frame(const C& c, int j, double* R, double* p) {
        const int NS = c.NS;
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = SA(F::RB, i, j);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = SA(F::PB, i, j);
    }
    static __device__ __forceinline__ void screw(const C& c, int j, double* s) {
        const int NS = c.NS;
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = SA(F::S, i, j);
    }
    static __device__ __forceinline__ void body_phi(const C& c, int j, double* phi) {
        const int NS = c.NS;
#pragma unroll
        for (int i = 0; i < 6; ++i) phi[i] = SA(F::PHI, i, j);
    }
};

__device__ __forceinline__ int krylov_count(const Ctx2L& c) { return c.kry_iters; }
__device__ __forceinline__ void krylov_reset(Ctx2L& c) { c.kry_iters = 0; }

// ---------------------------------------------------------------------------------------------
// newton() of driverRedMaxBDF1.m:94-157 (forward drivers: damped Newton + backtracking line search).
// Written as a two-state machine (FULL evaluation with H / residual-only line-search trial) so that the
// evaluation code has a single call site.
// ---------------------------------------------------------------------------------------------
// block-wide AND of a per-thread predicate (all decisions of the Newton loop are block-uniform)
template <int NW>
__device__ __forceinline__ bool block_all(bool p) {
    if (NW == 1) return __all_sync(0xffffffffu, p);
    return __syncthreads_and(p) != 0;
}

template <class E, int NW>
__device__ __forceinline__ int newton_forward(typename E::C& c, const StepOpts& op, int* perm, int& n_iter, int& n_ls) {
    // The reference evaluates g at every line-search trial (nargout == 1) and then [g,H] again at the accepted point when
    // the next iteration starts.  Both evaluations see the same x, so the trial evaluation here also produces what the
    // Newton matrix needs; an accepted, not yet converged trial is reused as the next iteration's evaluation (bitwise the
    // same numbers, one forward-kinematics pass less per iteration).  n_iter / n_ls count what the reference would do.
    // Written as a state machine so that the evaluation, assembly and LU code each have a single call site.
    //
    // Stalled solves (||g|| cannot get below tol: the reference then runs 10 nr iterations x 20 halvings, silently) are the
    // stragglers that set the makespan of a batch.  Two shortcuts remove the part of that work whose outcome is already known,
    // bit for bit (the evaluation is a pure function of the iterate and the step's history):
    //   * a trial point x0 + alpha dx that rounds to x0 in every component evaluates to f == f0, which the strict `<` rejects,
    //     and so does every later trial (alpha only shrinks): one evaluation at x0 stands for all of them;
    //   * if the line search ends there (x == x0), the next iteration starts from the very state this one started from and
    //     repeats it exactly, until iterMax: counted, not executed.
    const int t = threadIdx.x;
    const int nr = c.nr;
    int status = 0;
    int iter = 1;
    bool trial = false;  // false: first evaluation of the solve; true: line-search trial at x0 + alpha dx
    double f0 = 0.0, x0t = 0.0, dxt = 0.0, alpha = 1.0;
    int iterLs = 1;
    bool at_x0 = false;  // the trial being evaluated is x0 itself (and it is the last of this line search)
    while (true) {
        if (NW == 1) group_barrier(c.group, false);  // lockstep groups: every evaluation pass starts together
        E::base(c, true);
        const double gt = (t < nr) ? c.g[t] : 0.0;
        const double gsum = block_sum<NW>(gt * gt, c.red);
        if (trial) {  // driverRedMaxBDF1.m:123-141
            ++n_ls;
            bool accept = 0.5 * gsum < f0;
            if (!accept && iterLs >= op.iterLsMax) {
                status |= 4;  // line search exhausted: keep the last trial (driverRedMaxBDF1.m:135-138)
                accept = true;
            }
            if (!accept) {
                alpha = 0.5 * alpha;
                ++iterLs;
                double xt = x0t;
                if (t < nr) {
                    xt = __dadd_rn(x0t, __dmul_rn(alpha, dxt));
                    c.q[t] = xt;
                }
                if (op.shortcuts && iterLs < op.iterLsMax && block_all<NW>(xt == x0t)) {
                    // this and every remaining trial are x0: evaluate it once, as the last one
                    n_ls += op.iterLsMax - iterLs;
                    iterLs = op.iterLsMax;
                    at_x0 = true;
                }
                bsync<NW>();
                continue;
            }
            if (sqrt(gsum) < op.tol) break;
            if (iter >= op.iterMax) {
                status |= 2;
                break;
            }
            if (op.shortcuts && (status & 4) && (at_x0 || block_all<NW>(t >= nr || c.q[t] == x0t))) {
                // exhausted line search that ended on x0: iterations iter+1 .. iterMax repeat this one exactly
                n_iter += op.iterMax - iter;
                n_ls += (op.iterMax - iter) * op.iterLsMax;
                status |= 2;
                break;
            }
            ++iter;
        }
        E::columns(c, 1.0, c.beta, 1.0, 1.0, c.H);
        f0 = 0.5 * gsum;
        // dx = -H\g
        E::factor_solve(c, perm, -1.0, false);
        dxt = (t < nr) ? c.dx[t] : 0.0;
        const double dxn = sqrt(block_sum<NW>(dxt * dxt, c.red));
        ++n_iter;
        if (dxn > op.dxMax) {
            status |= 1;  // 'Newton diverged': x stays at the evaluation point (driverRedMaxBDF1.m:118-121)
            break;
        }
        if (t < nr) {
            x0t = c.q[t];
            c.q[t] = __dadd_rn(x0t, dxt);  // alpha = 1
        }
        alpha = 1.0;
        iterLs = 1;
        at_x0 = false;
        trial = true;
        bsync<NW>();
    }
    return status;
}

// ---------------------------------------------------------------------------------------------
// newton() of driverRedMaxAdjointBDF1.m:105-146 (= driverRedMaxAdjointBDF2.m:139-180): LU with row permutation,
// no line search, convergence test on the PRE-update residual, so one more full step is always taken after
// ||g|| < tol and the returned tape (LU(H), Hp, M, D, J) belongs to the last evaluation point.
// If `save`, that tape is written to global memory; if `want_J`, the task body's rows of J (6 x nr, the only rows
// TaskBDF1PointPos.calcStep reads) are left in shared memory behind H.
// ---------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void store_rowmajor(const Ctx& c, const double* H, double* __restrict__ dst) {
    const int nr = c.nr, ld = c.ld;
    for (int e = threadIdx.x; e < nr * nr; e += 32 * NW) {
        const int r = e / nr, i = e - r * nr;
        dst[e] = H[(size_t)i * ld + r];
    }
}

template <class E, int NW>
__device__ __forceinline__ int newton_adjoint(typename E::C& c, const StepOpts& op, int* perm, int& n_iter, bool save,
                                              double* __restrict__ tA, double* __restrict__ tM,
                                              double* __restrict__ tD, bool want_J, int jb) {
    const int t = threadIdx.x;
    const int nr = c.nr, ld = c.ld;
    int status = 0;
    int iter = 1;
    while (true) {
        E::base(c, true);
        const double gt = (t < nr) ? c.g[t] : 0.0;
        const double gsum = block_sum<NW>(gt * gt, c.red);
        E::columns(c, 1.0, c.beta, 1.0, 1.0, c.H);
        E::factor_solve(c, perm, -1.0, save);
        const double dxt = (t < nr) ? c.dx[t] : 0.0;
        const double dxn = sqrt(block_sum<NW>(dxt * dxt, c.red));
        ++n_iter;
        const bool diverged = dxn > op.dxMax;
        const bool conv = sqrt(gsum) < op.tol;
        const bool last = diverged || conv || iter >= op.iterMax;
        if (last && save) {
            for (int e = t; e < nr * ld; e += 32 * NW) tA[e] = c.H[e];
            if (t < nr) reinterpret_cast<int*>(tA + (size_t)nr * ld + nr)[t] = perm[t];
            bsync<NW>();
            E::columns(c, 0.0, 0.0, 1.0, 1.0, c.H);  // M = dg/d(dqtmp)
            store_rowmajor<NW>(c, c.H, tM);
            bsync<NW>();
            E::columns(c, 0.0, 1.0, 0.0, -1.0 / c.c, c.H);  // D = df/dqdot = -(1/cK) dg/dqdot
            store_rowmajor<NW>(c, c.H, tD);
            if (want_J) {
                // J(idxM(body), :) : column of joint k is Ad(E_body^-1) s_k for ancestors-or-self k of the body, else 0
                double* Jb = E::jrows(c);
                if (t < c.n && c.jc[t].idx >= 0) {
                    double col[6] = {0, 0, 0, 0, 0, 0};
                    if (t <= jb && jb < c.jc[t].end) {
                        double Rb[9], pb[3], st[6];
                        E::body_frame(c, jb, Rb, pb);
                        E::screw(c, t, st);
                        xm_w2b(Rb, pb, st, col);
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) Jb[6 * c.jc[t].idx + i] = col[i];
                }
            }
            bsync<NW>();
        }
        if (diverged) {
            status |= 1;
            break;
        }
        if (t < nr) c.q[t] = __dadd_rn(c.q[t], dxt);
        bsync<NW>();
        if (conv) break;
        if (iter >= op.iterMax) {
            status |= 2;
            break;
        }
        ++iter;
    }
    return status;
}

// The same Newton iteration for the tensor-core adjoint kernel (Eval::TCA).  The assembly overwrites the composite blocks, and
// the tape needs three matrices of the last evaluation point (H, M, D): whether an evaluation is the last one is known from its
// residual before anything is assembled (converged, or iterMax reached), so M and D are formed first, their tiles going
// straight to the tape, and the Newton matrix last.  Only a diverging step (||dx|| > dxMax, known after the solve) has to
// evaluate the point once more for M and D.
template <class E>
__device__ __forceinline__ void adjoint_tape_md(typename E::C& c, double* __restrict__ tM, double* __restrict__ tD, bool want_J, int jb) {
    const int t = threadIdx.x;
    E::columns_global(c, 0.0, 0.0, 1.0, 1.0, tM);           // M = dg/d(dqtmp)
    E::columns_global(c, 0.0, 1.0, 0.0, -1.0 / c.c, tD);    // D = df/dqdot = -(1/cK) dg/dqdot
    if (want_J) {
        // J(idxM(body), :) : column of joint k is Ad(E_body^-1) s_k for ancestors-or-self k of the body, else 0
        double* Jb = E::jrows(c);
        const bool mine = t < c.n && c.jc[t].idx >= 0;
        double col[6] = {0, 0, 0, 0, 0, 0};
        if (mine && t <= jb && jb < c.jc[t].end) {
            double Rb[9], pb[3], st[6];
            E::body_frame(c, jb, Rb, pb);
            E::screw(c, t, st);
            xm_w2b(Rb, pb, st, col);
        }
        __syncwarp();  // the rows overlay the body frames just read (TcLayoutA::JROWS)
        if (mine) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Jb[6 * c.jc[t].idx + i] = col[i];
        }
        __syncwarp();
    }
}

template <class E>
__device__ __forceinline__ int newton_adjoint_tc(typename E::C& c, const StepOpts& op, int* perm, int& n_iter, bool save,
                                                 double* __restrict__ tA, double* __restrict__ tM, double* __restrict__ tD,
                                                 bool want_J, int jb) {
    const int t = threadIdx.x;
    const int nr = c.nr, ldt = h_ld(nr);
    constexpr int LD = TcLayoutA::LD;
    int status = 0;
    int iter = 1;
    bool redo = false;  // second evaluation of a diverged point, for M and D only (single call site for every phase)
    while (true) {
        E::base(c, true);
        const double gt = (t < nr) ? c.g[t] : 0.0;
        const double gsum = block_sum<1>(gt * gt, c.red);
        const bool conv = sqrt(gsum) < op.tol;
        const bool lastk = conv || iter >= op.iterMax;
        if ((lastk || redo) && save) adjoint_tape_md<E>(c, tM, tD, want_J, jb);
        if (redo) {
            status |= 1;
            break;
        }
        E::columns(c, 1.0, c.beta, 1.0, 1.0, c.H);
        E::factor_solve(c, perm, -1.0, false);
        const double dxt = (t < nr) ? c.dx[t] : 0.0;
        const double dxn = sqrt(block_sum<1>(dxt * dxt, c.red));
        ++n_iter;
        const bool diverged = dxn > op.dxMax;
        if ((lastk || diverged) && save) {
            // lu(H,'vector'): the blocked LU leaves every row where it was (row perm[k] is the k-th pivot row); the tape wants
            // the factors in pivot order, column-major with leading dimension nr|1
            for (int e = t; e < nr * ldt; e += 32) {
                const int cI = e / ldt, k = e - cI * ldt;
                tA[e] = (k < nr) ? c.H[cI * LD + perm[k]] : 0.0;
            }
            if (t < nr) reinterpret_cast<int*>(tA + (size_t)nr * ldt + nr)[t] = perm[t];
            __syncwarp();
        }
        if (diverged) {
            if (save && !lastk) {  // x stays at this evaluation point: evaluate it once more for M, D (and J)
                redo = true;
                continue;
            }
            status |= 1;
            break;
        }
        if (t < nr) c.q[t] = __dadd_rn(c.q[t], dxt);
        __syncwarp();
        if (conv) break;
        if (iter >= op.iterMax) {
            status |= 2;
            break;
        }
        ++iter;
    }
    return status;
}

// ---------------------------------------------------------------------------------------------
// Forward rollout kernel: simLoop of driverRedMaxBDF1.m:57-91 / driverRedMaxBDF2.m:57-125 (ADJ = false) and of
// driverRedMaxAdjointBDF1.m:65-102 / driverRedMaxAdjointBDF2.m:65-136 (ADJ = true), one block per rollout.
// ---------------------------------------------------------------------------------------------
// Next segment of a scheduled launch for this block (NW > 1) or warp (NW == 1; lockstep groups: every warp claims on its own):
// from list `victim` while it has unclaimed segments, else from the next list that has (32 candidates per probe round).
// Returns the segment index or -1 when every list is empty.  Uniform over the block / warp.
template <int NW>
__device__ __forceinline__ int claim_segment(const RolloutArgs& a, int& victim, int nslots, int* bc) {
    const int lane = threadIdx.x & 31;
    const bool w0 = threadIdx.x < 32;
    int seg = -1;
    if (w0) {
        while (true) {
            int idx = 0;
            if (lane == 0) idx = atomicAdd(a.cursor + victim, 1);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            const int lo = __ldg(a.seg_off + victim), hi = __ldg(a.seg_off + victim + 1);
            if (lo + idx < hi) {
                seg = lo + idx;
                break;
            }
            int found = -1;
            for (int base = 1; base < nslots && found < 0; base += 32) {
                int v = victim + base + lane;
                v -= (v >= nslots) ? nslots : 0;
                bool has = false;
                if (base + lane < nslots) {
                    const int left = __ldg(a.seg_off + v + 1) - __ldg(a.seg_off + v);
                    has = *reinterpret_cast<volatile int*>(a.cursor + v) < left;
                }
                const unsigned m = __ballot_sync(0xffffffffu, has);
                if (m) {
                    found = victim + base + __ffs(m) - 1;
                    found -= (found >= nslots) ? nslots : 0;
                }
            }
            if (found < 0) break;
            victim = found;
        }
    }
    if (NW > 1) {  // one block per rollout: warp 0 claimed for all
        __syncthreads();
        if (threadIdx.x == 0) {
            bc[0] = seg;
            bc[1] = victim;
        }
        __syncthreads();
        seg = bc[0];
        victim = bc[1];
    }
    return seg;
}

// Register budget of the one-warp tensor-core forward kernel.  Measured on B200 (profiles/r01_residency_probe.log,
// r01_regcap_ab.log): the kernel is latency-bound per warp -- time per block is almost flat in the number of resident blocks
// (59.7 us per rollout-step alone on an SM, 71.4 us with 8 resident) -- but capping registers to fit 10 blocks per SM
// (200 registers, ~400 B of spills) slows every block by 27 % and loses more than the third wave gains (34.1 vs 27.2 ms), and
// 184 registers is worse still.  So: no cap (255 registers, 8 blocks per SM).
#ifndef RMX_MAXNREG_FWD
#define RMX_MAXNREG_FWD 255
#endif
#define RMX_FWD_BOUNDS __maxnreg__((NW == 1 && !ADJ && LIN == 0 && IMPL == 2) ? RMX_MAXNREG_FWD : 255)
template <int NW, int GROUND, bool ADJ, int IMPL, int LIN>
__global__ void RMX_FWD_BOUNDS rollout_fwd_kernel(RolloutArgs a) {
    typedef Eval<IMPL, NW, GROUND, ADJ || LIN == 1, LIN, ADJ> E;
    extern __shared__ double2 smem_raw[];
    // lockstep group (group_barrier, rmx_device.cuh): warp threadIdx.y of the block is an independent rollout slot with its own
    // shared-memory region; slots are numbered so that neighbouring slots (the two parts of a cut rollout) sit in different blocks
    constexpr bool CAN_GROUP = NW == 1 && !ADJ && LIN == 0;
    constexpr int MAXG = CAN_GROUP ? RMX_MAX_GROUP : 1;
    const int G = CAN_GROUP ? (int)blockDim.y : 1;
    const int gy = CAN_GROUP ? (int)threadIdx.y : 0;
    double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)gy * a.group_stride;
    __shared__ int perm_all[32 * NW * MAXG];
    int* perm_s = perm_all + 32 * NW * gy;
    const long long slot = (long long)blockIdx.x + (long long)gy * gridDim.x, nslots = (long long)gridDim.x * G;
    const int t = threadIdx.x;
    const int nr = a.sc.nr;
    typename E::C c;
    E::setup(c, sm, a.sc, a.op);
    c.group = G;
    const StepOpts op = a.op;
    const double h = op.h;
    const double ah = __dmul_rn(SDIRK_A_CONST, h);
    const double bh = __dmul_rn(__dsub_rn(1.0, SDIRK_A_CONST), h);

    constexpr bool CAN_SCHED = !ADJ && LIN == 0;
    const bool sched = CAN_SCHED && a.seg != nullptr;
    long long it = slot;   // plain launches: rollouts slot, slot + nslots, ...
    // scheduled launches: the list this block is claiming segments from (claim_segment); slots beyond the lists only steal
    int victim = sched ? (int)(slot % a.nlists) : 0;
    while (true) {
        long long b = it;
        int k_begin = 0, k_end = op.nsteps, seg_flags = 0;
        if (sched) {
            const int si = claim_segment<NW>(a, victim, a.nlists, perm_s);  // (perm_s is free between rollouts)
            if (si < 0) break;
            const int4 sg = __ldg(a.seg + si);
            b = sg.x;
            k_begin = sg.y;
            k_end = sg.z;
            seg_flags = sg.w;
        } else {
            if (it >= a.B) break;
            it += nslots;
        }
        double qc = 0.0, qdc = 0.0;  // current state of dof t (joint.q / joint.qdot)
        double q0t = 0.0, qd0t = 0.0, q1t = 0.0, qdat = 0.0;
        int status = 0, n_iter = 0, n_ls = 0;
        if (CAN_SCHED && (seg_flags & 1)) {
            // second part of a cut rollout: wait until the block that runs the first part has published it (bounded spin: a
            // lost signal becomes a status bit, never a hang), then resume from the trajectory already in global memory
            if (G > 1) {  // a grouped warp keeps meeting its partners while it waits (they check in once per evaluation pass);
                          // a round lasts as long as the partners' passes, so the wait is bounded by time (~13 s), not by rounds
                unsigned long long t0 = 0;
                if (t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                int ready = 0;
                while (true) {
                    if (t == 0) {
                        ready = atomicAdd(a.flags + b, 0) != 0;
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (!ready && now - t0 > 13000000000ull) ready = 2;  // gave up
                    }
                    ready = __shfl_sync(0xffffffffu, ready, 0);
                    if (ready) break;
                    group_barrier(G, false);
                }
                if (ready == 2) status |= 16;
            } else if (t == 0) {
                unsigned spins = 0;
                while (atomicAdd(a.flags + b, 0) == 0 && ++spins < (1u << 26)) __nanosleep(200);
                if (spins >= (1u << 26)) status |= 16;
            }
            bsync<NW>();
            __threadfence();
        }
        if (t < nr) {
            if (CAN_SCHED && k_begin > 0) {
                const size_t o1 = ((size_t)b * op.nsteps + (k_begin - 1)) * nr + t;
                qc = __ldcg(a.q_out + o1);
                qdc = __ldcg(a.qd_out + o1);
                if (k_begin > 1) {  // BDF2 history: the state two steps back
                    c.hq1[t] = __ldcg(a.q_out + o1 - nr);
                    c.hqd1[t] = __ldcg(a.qd_out + o1 - nr);
                } else {            // after the SDIRK2 first step the history is the initial state (driverRedMaxBDF2.m:86-91)
                    c.hq1[t] = a.q0[b * nr + t];
                    c.hqd1[t] = a.qd0[b * nr + t];
                }
            } else {
                qc = a.q0[b * nr + t];
                qdc = a.qd0[b * nr + t];
                c.hq1[t] = qc;
                c.hqd1[t] = qdc;
            }
            if (ADJ)  // task.applyStep: joints{i}.tau = pscale*p(idxR)   (TaskBDF1PointPos.m:58-64)
                c.tau[t] = __dmul_rn(a.task.pscale, a.task.p[b * nr + t]);
            else
                c.tau[t] = (op.tau_mode == 1) ? a.tau[b * nr + t] : 0.0;
        }
        double tcur = 0.0, Pacc = 0.0;  // scene.t (Scene.m:122), task.P
        for (int k = k_begin; k < k_end; ++k) {
            if (!ADJ && op.tau_mode == 2 && t < nr) c.tau[t] = a.tau[((size_t)b * op.nsteps + k) * nr + t];
            // scene.t after this step; TaskBDF1PointPos.calcStep samples the objective when |t_target - t| < 1e-6
            const double tnext = __dadd_rn(tcur, h);
            const bool is_obj = ADJ && (fabs(__dsub_rn(a.task.t_target, tnext)) < 1e-6);
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
                if (ADJ) {
                    // the BDF2 adjoint driver keeps only the second SDIRK sub-solve's tape (driverRedMaxAdjointBDF2.m:88,96)
                    const bool save = stage != ST_SDIRK_A;
                    const size_t rec = (size_t)b * op.nsteps + k;
                    if constexpr (E::TCA)
                        status |= newton_adjoint_tc<E>(c, op, perm_s, n_iter, save, a.tape.A + rec * a.tape.sza,
                                                       a.tape.M + rec * nr * nr, a.tape.D + rec * nr * nr, save && is_obj,
                                                       a.task.body);
                    else
                        status |= newton_adjoint<E, NW>(c, op, perm_s, n_iter, save, a.tape.A + rec * a.tape.sza,
                                                        a.tape.M + rec * nr * nr, a.tape.D + rec * nr * nr, save && is_obj,
                                                        a.task.body);
                } else {
                    status |= newton_forward<E, NW>(c, op, perm_s, n_iter, n_ls);
                }
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
            tcur = tnext;
            if (a.sc.has_chart && t < a.sc.n) {
                // jroot.reparam() (driverRedMaxBDF1.m:78, driverRedMaxBDF2.m:112): the chart is fixed here, the place where
                // JointSpherical.reparam_ would leave it (|det T| <= 0.5, JointSpherical.m:63-67) is reported
                const JointConst& Jc = a.sc.jc[t];
                if (Jc.chart_mid) {
                    const double qm = c.q[Jc.idx];
                    const double detT = Jc.chart_mid == 2 ? sin(qm) : cos(qm);  // proper Euler / Tait-Bryan chart
                    if (!(fabs(detT) > 0.5)) status |= 32;
                }
            }
            if (t < nr) {
                const size_t o = ((size_t)b * op.nsteps + k) * nr + t;
                if (a.q_out) a.q_out[o] = qc;
                if (a.qd_out) a.qd_out[o] = qdc;
                if (a.q_host) __stcs(a.q_host + o, qc);
                if (a.qd_host) __stcs(a.qd_host + o, qdc);
            }
            if (ADJ) {
                // task.calcStep (TaskBDF1PointPos.m:67-107): objective and dP/dq_k at the stored (final) q, with the
                // tape's J, i.e. the Jacobian of the last Newton evaluation point (note N4 of SURVEY.md)
                double dPdq = 0.0;
                if (is_obj) {
                    bsync<NW>();
                    double jr[6] = {0, 0, 0, 0, 0, 0};  // this thread's row of the tape's J: read before the evaluation below
                    if (t < nr) {                        // rebuilds the body frames it may share storage with (TcLayoutA)
#pragma unroll
                        for (int i = 0; i < 6; ++i) jr[i] = E::jrows(c)[6 * t + i];
                    }
                    bsync<NW>();
                    E::base(c, false);  // c.q holds the final iterate: FK at history(k).q
                    double rb[9], pbt[3];
                    E::body_frame(c, a.task.body, rb, pbt);
                    double xw[3], dx3[3], y[3], v6[6];
                    mat3_vec(rb, a.task.xlocal, xw);
#pragma unroll
                    for (int i = 0; i < 3; ++i) dx3[i] = (xw[i] + pbt[i]) - a.task.xtarget[3 * b + i];
                    Pacc += a.task.wpos * 0.5 * (dx3[0] * dx3[0] + dx3[1] * dx3[1] + dx3[2] * dx3[2]);
                    mat3T_vec(rb, dx3, y);               // R' dx
                    cross3(a.task.xlocal, y, v6);        // Gamma' y = [xlocal x y ; y]
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        v6[i] *= a.task.wpos;
                        v6[3 + i] = y[i] * a.task.wpos;
                    }
                    if (t < nr) dPdq = dot6(jr, v6);
                }
                if (t < nr) {
                    double* recA = a.tape.A + ((size_t)b * op.nsteps + k) * a.tape.sza;
                    recA[(size_t)nr * h_ld(nr) + t] = dPdq;
                }
            }
            bsync<NW>();
        }
        // block-wide OR of the two per-thread conditions in one reduction: non-finite state (units) and a chart flag (4096s)
        double bad = (t < nr && !(isfinite(qc) && isfinite(qdc))) ? 1.0 : 0.0;
        if (status & 32) bad += 4096.0;
        bad = block_sum<NW>(bad, c.red);
        if (bad >= 4096.0) {
            status |= 32;
            bad -= 4096.0 * floor(bad / 4096.0);
        }
        if (bad > 0.0) status |= 8;
        if (CAN_SCHED && (seg_flags & 2)) {  // first part of a cut rollout: publish the trajectory written so far
            __threadfence();
            bsync<NW>();
            if (t == 0) atomicExch(a.flags + b, 1);
        }
        if (t == 0) {
            if (sched) {  // a rollout may arrive in two parts: accumulate (status / iters are zeroed before the launch)
                if (status) atomicOr(a.status + b, status);
                if (a.iters) {
                    atomicAdd(a.iters + 2 * b, n_iter);
                    atomicAdd(a.iters + 2 * b + 1, n_ls);
                }
            } else {
                a.status[b] = status;
                if (a.iters) {
                    a.iters[2 * b] = n_iter;
                    a.iters[2 * b + 1] = n_ls;
                }
            }
            if (ADJ) a.task.P[b] = Pacc;  // objective part; the regulariser is added by the backward kernel
            if (LIN == 1 && IMPL == 2 && a.kry_total) {
                atomicAdd(a.kry_total, (unsigned long long)krylov_count(c));
                krylov_reset(c);
            }
        }
        bsync<NW>();
    }
    if (CAN_GROUP && G > 1) {  // out of work: keep the block's barrier going until every warp of the group is
        while (!group_barrier(G, true)) {
        }
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

template <int NW, int GROUND, int IMPL>
__global__ void __launch_bounds__(32 * NW) eval_kernel(EvalArgs a) {
    typedef Eval<IMPL, NW, GROUND, true, 0> E;
    extern __shared__ double2 smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const int t = threadIdx.x;
    const int nr = a.sc.nr;
    typename E::C c;
    StepOpts op0;
    op0.shortcuts = 0;
    op0.lin_tol = 0.0;
    op0.lin_maxit = 0;
    E::setup(c, sm, a.sc, op0);
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
    E::base(c, true);
    if (t < nr && a.g) a.g[t] = c.g[t];
    const int ld = c.ld;
    for (int pass = 0; pass < 3; ++pass) {
        double* dst = pass == 0 ? a.H : (pass == 1 ? a.M : a.D);
        if (!dst) continue;
        if (pass == 0) E::columns(c, 1.0, c.beta, 1.0, 1.0, c.H);
        if (pass == 1) E::columns(c, 0.0, 0.0, 1.0, 1.0, c.H);
        if (pass == 2) E::columns(c, 0.0, 1.0, 0.0, -1.0 / c.c, c.H);
        for (int e = t; e < nr * nr; e += blockDim.x) {
            const int col = e / nr, row = e % nr;
            dst[e] = c.H[(size_t)col * ld + row];
        }
        bsync<NW>();
    }
}

// ---------------------------------------------------------------------------------------------
// Test hook: one Newton linear system exactly as the forward rollout kernel forms and solves it (same Eval instance:
// tensor-core assembly + blocked LU for one warp) -> H (before factorisation, nr x nr column-major) and dx = -H \ g.
// ---------------------------------------------------------------------------------------------
template <int NW, int GROUND>
__global__ void __launch_bounds__(32 * NW) eval_newton_kernel(EvalArgs a, double* dx_out) {
    typedef Eval<2, NW, GROUND, false, 0> E;
    extern __shared__ double2 smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    __shared__ int perm_s[32 * NW];
    const int t = threadIdx.x;
    const int nr = a.sc.nr;
    typename E::C c;
    StepOpts op0;
    op0.shortcuts = 0;
    op0.lin_tol = 0.0;
    op0.lin_maxit = 0;
    E::setup(c, sm, a.sc, op0);
    c.stage = ST_DIRECT;
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
    E::base(c, true);
    E::columns(c, 1.0, c.beta, 1.0, 1.0, c.H);
    const int ld = c.ld;
    if (a.H) {
        for (int e = t; e < nr * nr; e += blockDim.x) {
            const int col = e / nr, row = e % nr;
            a.H[e] = c.H[(size_t)col * ld + row];
        }
    }
    bsync<NW>();
    E::factor_solve(c, perm_s, -1.0, false);
    bsync<NW>();
    if (t < nr && dx_out) dx_out[t] = c.dx[t];
}

// ---------------------------------------------------------------------------------------------
// Test hook: the two operators of the Krylov solve at one evaluation point -- y = H x through the matrix-free tree sweeps
// and z = (J' blkdiag(M_j) J + Pr)^-1 x through the projected block-Jacobi preconditioner (notes.pdf Alg. 10).
// ---------------------------------------------------------------------------------------------
template <int NW, int GROUND>
__global__ void __launch_bounds__(32 * NW) eval_krylov_kernel(EvalArgs a, const double* x, double* hx, double* pinvx) {
    typedef Eval<2, NW, GROUND, true, 1> E;
    extern __shared__ double2 smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const int t = threadIdx.x;
    const int nr = a.sc.nr;
    typename E::C c;
    StepOpts op0;
    op0.shortcuts = 0;
    op0.lin_tol = 0.0;
    op0.lin_maxit = 0;
    E::setup(c, sm, a.sc, op0);
    c.stage = ST_DIRECT;
    c.h = 1.0;
    c.c = a.cK;
    c.beta = a.beta;
    double* xs = c.pm.vec;           // r slot
    double* ys = c.pm.vec + nr;      // r0 slot
    if (t < nr) {
        c.q[t] = a.q[t];
        c.hqd0[t] = a.qd[t];
        c.hq1[t] = a.dq[t];
        c.hq0[t] = 0;
        c.hqd1[t] = 0;
        c.tau[t] = a.tau ? a.tau[t] : 0.0;
        xs[t] = x[t];
    }
    bsync<NW>();
    E::base(c, true);
    E::columns(c, 1.0, c.beta, 1.0, 1.0, c.H);
    hx_apply<NW, GROUND>(c, c.pm, xs, ys);
    if (t < nr && hx) hx[t] = ys[t];
    bsync<NW>();
    precond_setup<NW, GROUND>(c, c.pm);
    precond_apply<NW, GROUND>(c, c.pm, xs, ys);
    if (t < nr && pinvx) pinvx[t] = ys[t];
}

}  // namespace rmx
