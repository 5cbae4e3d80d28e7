// rmx_adjoint.cuh -- adjoint backward recursion (Task*.calcFinal) and the energy kernel.
//
// Replaces, for B rollouts at once:
//   TaskBDF1.calcFinal   (matlab-diff/+redmax/TaskBDF1.m:45-81)
//   TaskBDF2.calcFinal   (matlab-diff/+redmax/TaskBDF2.m:45-108)
//   TaskBDF1PointPos / TaskBDF2PointPos dgdp (TaskBDF1PointPos.m:106, TaskBDF2PointPos.m:106)
//   Scene.saveHistory energies (Scene.m:155-160; Joint.m:616-637, Body.m:167-173, ForceGroundCuboid.m:156-183)
//
// The reference solves the block upper-triangular adjoint system backwards in time,
//     H_k' z_k = dPdq_k - sum_{d=1..4} B_{k,k+d}' z_{k+d},       z_k(Hp) = Hl' \ (Hu' \ y_k),
// re-reading M, D of steps k+1..k+4 for every k.  Here every step's tape is read exactly once: when z_j is known,
// u = M_j' z_j and v = D_j' z_j are formed and pushed into the (register-resident) right-hand sides of the up to
// four earlier steps that reference block j.  The LU record of step k-1 is fetched by a TMA bulk copy
// (cp.async.bulk + mbarrier) into the other half of a shared-memory double buffer while step k is being solved;
// M and D are streamed straight from HBM with coalesced loads (row-major tape, thread i owns column i of M').
// dPdp = wreg*p - z'*dgdp with dgdp_k = coef*h^2*pscale*I collapses to wreg*p - coef*h^2*pscale * sum_k z_k.
#pragma once
#include "rmx_rollout.cuh"

namespace rmx {

struct BwdArgs {
    int nr, nsteps, scheme;
    double h;
    long long B;
    TapeArgs tape;
    TaskArgs task;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16B-aligned addresses)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// NBUF = 2: double-buffered tape records; NBUF = 1 when two records do not fit in shared memory (nr > 80 or so).
template <int NW, int NBUF>
__global__ void __launch_bounds__(32 * NW) adjoint_bwd_kernel(BwdArgs a) {
    extern __shared__ double2 smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    __shared__ unsigned long long bars[2];
    const int t = threadIdx.x;
    const int nr = a.nr, ld = h_ld(nr), sza = a.tape.sza, ns = a.nsteps;
    double* buf[2] = {sm, sm + (NBUF == 2 ? sza : 0)};
    double* w = sm + (size_t)NBUF * sza;  // [nr] substitution scratch
    double* y = w + nr;                   // [nr] right-hand side
    double* z = y + nr;                   // [nr] solution
    Ctx c;
    c.nr = nr;
    c.ld = ld;
    c.red = z + nr;
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned rec_bytes = (unsigned)sza * 8u;
    const double acoef = SDIRK_A_CONST;
    unsigned phase[2] = {0u, 0u};
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double* A = a.tape.A + (size_t)b * ns * sza;
        const double* Mg = a.tape.M + (size_t)b * ns * nr * nr;
        const double* Dg = a.tape.D + (size_t)b * ns * nr * nr;
        double zsum = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
        int slot = 0;
        if (t == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bars[0], rec_bytes);
            bulk_g2s(buf[0], A + (size_t)(ns - 1) * sza, rec_bytes, &bars[0]);
        }
        for (int k = ns - 1; k >= 0; --k) {
            if (NBUF == 2 && k > 0 && t == 0) {
                // the other buffer was last read two steps ago; every thread has passed a block barrier since
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[slot ^ 1], rec_bytes);
                bulk_g2s(buf[slot ^ 1], A + (size_t)(k - 1) * sza, rec_bytes, &bars[slot ^ 1]);
            }
            mbar_wait(&bars[slot], phase[slot]);
            phase[slot] ^= 1u;
            const double* LU = buf[slot];
            const double* dPdq = LU + (size_t)nr * ld;
            const int* perm = reinterpret_cast<const int*>(dPdq + nr);
            if (t < nr) y[t] = dPdq[t] + p1;
            bsync<NW>();
            lu_solve_T<NW>(c, LU, perm, y, w, z);  // z(Hp) = Hl' \ (Hu' \ y)      (TaskBDF1.m:74-77)
            double u = 0.0, v = 0.0;
            if (t < nr) {
                zsum += z[t];
                const double* Mk = Mg + (size_t)k * nr * nr + t;
                const double* Dk = Dg + (size_t)k * nr * nr + t;
#pragma unroll 8
                for (int r = 0; r < nr; ++r) {
                    const double zr = z[r];
                    u += __ldg(Mk + (size_t)r * nr) * zr;
                    v += __ldg(Dk + (size_t)r * nr) * zr;
                }
            }
            // push block column k into the right-hand sides of steps k-1 .. k-4 (1-based target index = k, k-1, ...)
            const double hv = a.h * v;
            if (a.scheme == 1) {
                // TaskBDF1.m:58-72:  B_{k-1,k} = -2M + hD ;  B_{k-2,k} = M
                p1 = p2 + (2.0 * u - hv);
                p2 = -u;
            } else {
                // TaskBDF2.m:66-99; the SDIRK first step (1-based target 1 == 0-based 0) has its own coefficients
                const double c1 = (k - 1 == 0) ? -((8.0 / (9.0 * acoef)) + (4.0 / 3.0)) : -(8.0 / 3.0);
                const double c2 = (k - 2 == 0) ? ((2.0 / (9.0 * acoef)) + (19.0 / 9.0)) : (22.0 / 9.0);
                const double n1 = p2 - (c1 * u + (8.0 / 9.0) * hv);
                const double n2 = p3 - (c2 * u - (2.0 / 9.0) * hv);
                const double n3 = p4 + (8.0 / 9.0) * u;
                const double n4 = -(1.0 / 9.0) * u;
                p1 = n1;
                p2 = n2;
                p3 = n3;
                p4 = n4;
            }
            bsync<NW>();
            if (NBUF == 1 && k > 0 && t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[0], rec_bytes);
                bulk_g2s(buf[0], A + (size_t)(k - 1) * sza, rec_bytes, &bars[0]);
            }
            if (NBUF == 2) slot ^= 1;
        }
        if (t < nr) {
            const double pt = a.task.p[b * nr + t];
            // dgdp_k = -h^2*pscale*I (BDF1) / -(4/9)h^2*pscale*I for every step incl. the SDIRK one (BDF2, note N7)
            const double coef = (a.scheme == 1 ? -1.0 : -(4.0 / 9.0)) * (a.h * a.h) * a.task.pscale;
            a.task.dPdp[b * nr + t] = a.task.wreg * pt - coef * zsum;
        }
        // P = task.P + wreg/2 p'p   (TaskBDF1.m:49)
        const double pp = (t < nr) ? a.task.p[b * nr + t] : 0.0;
        const double pn = block_sum<NW>(pp * pp, c.red);
        if (t == 0) a.task.P[b] = a.task.P[b] + a.task.wreg * 0.5 * pn;
        bsync<NW>();
    }
}

// ---------------------------------------------------------------------------------------------
// Energies of Scene.saveHistory: T = sum 1/2 phi' I phi ; V = sum -m grav'p_wi + joint springs/limits + ground
// penalty.  One block per state; reuses the FK of eval_base.
// ---------------------------------------------------------------------------------------------
struct EnergyArgs {
    DevScene sc;
    long long B;
    const double* q;
    const double* qd;
    double* T;
    double* V;
    // optional (rmx_body_frames): world frames E_wi of the bodies, 4x4 column-major each, [B][nbody][16]
    double* E_out;
    const int* body_int;  // [nbody] internal joint index of each body
    int nbody;
};

template <int NW, int GROUND, int IMPL>
__global__ void __launch_bounds__(32 * NW) energies_kernel(EnergyArgs a) {
    typedef Eval<IMPL, NW, GROUND, true, 0> E;
    extern __shared__ double2 smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const int t = threadIdx.x;
    const int n = a.sc.n, nr = a.sc.nr;
    typename E::C c;
    StepOpts op0;
    op0.shortcuts = 0;
    op0.lin_tol = 0.0;
    op0.lin_maxit = 0;
    E::setup(c, sm, a.sc, op0);
    c.stage = ST_DIRECT;
    c.h = 1.0;
    c.c = 1.0;
    c.beta = 1.0;
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        if (t < nr) {
            c.q[t] = a.q[b * nr + t];
            c.hqd0[t] = a.qd[b * nr + t];
            c.hq1[t] = 0;
            c.hq0[t] = 0;
            c.hqd1[t] = 0;
            c.tau[t] = 0;
        }
        bsync<NW>();
        E::base(c, false);
        double T = 0.0, V = 0.0;
        if (t < n) {
            const JointConst& J = c.jc[t];
            double r1[12], phi[6];
            E::body_frame(c, t, r1, r1 + 9);
            E::body_phi(c, t, phi);
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) acc += phi[i] * (J.I[i] * phi[i]);
            T = 0.5 * acc;
            V = -J.I[5] * (c.gx * r1[9] + c.gy * r1[10] + c.gz * r1[11]);
            if (J.idx >= 0) {
                const double qk = c.q[J.idx];
                const double dq = qk - J.qRest;
                V += 0.5 * J.stiff * (dq * dq);
                const double dL = (qk < J.qLimL) ? (J.qLimL - qk) : 0.0;
                const double dU = (qk > J.qLimU) ? (J.qLimU - qk) : 0.0;
                V += 0.5 * J.qLimK * (dL * dL + dU * dU);
            }
            if (GROUND && J.has_ground) {
                double nb[3];
                mat3T_vec(r1, J.gng, nb);
                const double dp = J.gng[0] * (r1[9] - J.gxg[0]) + J.gng[1] * (r1[10] - J.gxg[1]) + J.gng[2] * (r1[11] - J.gxg[2]);
                for (int ci = 0; ci < 8; ++ci) {
                    const double xl[3] = {(ci & 4) ? J.hs[0] : -J.hs[0], (ci & 2) ? J.hs[1] : -J.hs[1],
                                          (ci & 1) ? J.hs[2] : -J.hs[2]};
                    const double d = nb[0] * xl[0] + nb[1] * xl[1] + nb[2] * xl[2] + dp;
                    if (d > 0) continue;
                    V += 0.5 * J.gkn * (d * d);
                }
            }
        }
        if (GROUND == 2 && t < c.npf) {  // Force*.computeEnergy_ (ForcePointPoint.m:116, ForceSpringGeneric.m:146, ForceSpringMultiPointGeneric.m:193)
            const PointForce& P = c.pf[t];
            double xprev[3] = {0, 0, 0}, len = 0.0, l2 = 0.0;
            for (int k = 0; k < P.npts; ++k) {
                const double xl[3] = {P.x[k][0], P.x[k][1], P.x[k][2]};
                double xw[3];
                if (P.body[k] >= 0) {
                    double rb[12];
                    E::body_frame(c, P.body[k], rb, rb + 9);
                    mat3_vec(rb, xl, xw);
                    xw[0] += rb[9]; xw[1] += rb[10]; xw[2] += rb[11];
                } else {
                    xw[0] = xl[0]; xw[1] = xl[1]; xw[2] = xl[2];
                }
                if (k > 0) {
                    const double d0 = xw[0] - xprev[0], d1 = xw[1] - xprev[1], d2 = xw[2] - xprev[2];
                    l2 = d0 * d0 + d1 * d1 + d2 * d2;
                    len += sqrt(l2);
                }
                xprev[0] = xw[0]; xprev[1] = xw[1]; xprev[2] = xw[2];
            }
            if (P.kind == 0) {
                V += 0.5 * P.ks * l2;  // zero rest length, linear
            } else {
                const double strain = (len - P.L) / P.L;  // spring-damper; cable: only when stretched (ForceCable.m:70)
                if (P.kind == 1 || strain > 0) V += 0.5 * P.ks * strain * strain * P.L;
            }
        }
        if (a.E_out) {  // Body.update: E_wi = E_wj * E0_ji  (Body.m:70-80)
            for (int j = t; j < a.nbody; j += 32 * NW) {
                double rb[12];
                E::body_frame(c, a.body_int[j], rb, rb + 9);
                double* Eo = a.E_out + ((size_t)b * a.nbody + j) * 16;
#pragma unroll
                for (int col = 0; col < 3; ++col) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) Eo[4 * col + r] = rb[3 * r + col];
                    Eo[4 * col + 3] = 0.0;
                }
                Eo[12] = rb[9]; Eo[13] = rb[10]; Eo[14] = rb[11]; Eo[15] = 1.0;
            }
        }
        T = block_sum<NW>(T, c.red);
        V = block_sum<NW>(V, c.red);
        if (t == 0 && a.T) {
            a.T[b] = T;
            a.V[b] = V;
        }
        bsync<NW>();
    }
}

}  // namespace rmx
