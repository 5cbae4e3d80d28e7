/*
 * redmax_b200.h -- C ABI of the B200-native batched RedMax stepper.
 *
 * This is the drop-in boundary for the hot path of sueda/redmax `matlab-diff`: the time loop of
 * driverRedMaxBDF1.m:57-91 / driverRedMaxBDF2.m:57-125 (simLoop + newton + evalBDF* + computeValues and
 * every +redmax Joint/Body/Force method they call) and of driverRedMaxAdjointBDF1/2.m (simLoop + newton +
 * Task*.calcStep/calcFinal), run for B independent rollouts on the GPU.  The reference has no FFI layer;
 * the binding a maintainer adds is a MEX gateway (matlab/redmax_mex.cpp) that passes
 * mxGetDoubles() pointers straight into these functions.  The same symbols are driven from Python (ctypes)
 * by redmax_b200/_ffi.py for every test and benchmark.
 *
 * Conventions
 *   - All matrices are column-major float64 exactly as MATLAB stores them; the batch is the trailing
 *     dimension (q0 is nr x B, q_out is nr x nsteps x B).
 *   - Joints are listed parents-before-children, as the reference requires (Joint.m:134-146).  Reduced
 *     indices follow the reference's leaf-to-root numbering (Scene.m:69-71): the LAST joint in the list owns
 *     q(1), a joint's own DOFs are consecutive (Joint.m:152); the library computes that numbering itself from
 *     the `ndof` implied by `jtype`.
 *   - Host-pointer entry points copy H2D/D2H internally and block until done.  `_dev` entry points take
 *     device pointers on the current CUDA device and enqueue on the given stream without synchronising.
 *   - Errors: 0 on success, negative RMX_E* otherwise; rmx_last_error() returns a message.  No exceptions
 *     cross the ABI.  Not re-entrant per scene handle (MATLAB calls from its single main thread).
 *   - There is no CPU fallback: every compute entry point fails with RMX_ENOGPU when no CUDA device exists.
 */
#ifndef REDMAX_B200_H
#define REDMAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMX_VERSION 109

/* error codes */
#define RMX_OK 0
#define RMX_EINVAL (-1)  /* bad argument / unsupported scene */
#define RMX_ENOGPU (-2)  /* no CUDA device / driver */
#define RMX_ECUDA (-3)   /* CUDA runtime error (message in rmx_last_error) */
#define RMX_ENOMEM (-4)
#define RMX_ELIMIT (-5)  /* scene exceeds kernel limits (n <= 128 joints) */

/* joint types (matlab-diff/+redmax/Joint*.m); ndof in brackets */
#define RMX_JOINT_FIXED 0         /* JointFixed.m          [0] */
#define RMX_JOINT_REVOLUTE 1      /* JointRevolute.m       [1] rotation about `axis` */
#define RMX_JOINT_PRISMATIC 2     /* JointPrismatic.m      [1] translation along `axis` */
#define RMX_JOINT_PLANAR 3        /* JointPlanar.m         [2] translation in the plane spanned by `axis`, `axis2` */
#define RMX_JOINT_TRANSLATIONAL 4 /* JointTranslational.m  [3] translation x, y, z */
#define RMX_JOINT_FREE2D 5        /* JointFree2D.m         [3] translation x, y then rotation about z */
#define RMX_JOINT_UNIVERSAL 6     /* JointUniversal.m      [2] rotation about x then y */
#define RMX_JOINT_SPHERICAL 7     /* JointSpherical.m      [3] Euler angles R = R_a(q1) R_b(q2) R_c(q3) in the chart given by
                                     rmx_scene_desc.chart (default XYZ, the constructor's: JointSpherical.m:33).  A rollout never
                                     changes chart by itself: where the reference re-parameterises (|det T| <= 0.5 after a step,
                                     JointSpherical.m:63-67) it sets RMX_ST_CHART and carries on; the caller re-expresses that step
                                     and resumes under the new chart (rmx_rollout_resume) */
#define RMX_JOINT_FREE3D 8        /* JointFree3D.m         [6] translation x, y, z (q1..q3) then a spherical joint (q4..q6) */
#define RMX_MAX_JOINT_DOF 6
#define RMX_MAX_POINTFORCE 8
#define RMX_MAX_CABLE_POINTS 4

/* two-point forces (pf_* arrays) */
#define RMX_FORCE_POINTPOINT 0   /* ForcePointPoint.m: linear, zero rest length: f = ks dx + kd dv */
#define RMX_FORCE_SPRINGDAMPER 1 /* ForceSpringDamper.m (ForceSpringGeneric.m): along the line, f = ks (l-L)/L + kd ldot/L */
#define RMX_FORCE_CABLE 2        /* ForceCable.m (ForceSpringMultiPointGeneric.m): routed through cable_* points, pulls only when
                                    stretched; not a pf_kind value -- cables are described by the cable_* arrays */

/* integrators (driverRedMaxBDF1.m, driverRedMaxBDF2.m) */
#define RMX_SCHEME_BDF1 1
#define RMX_SCHEME_BDF2 2 /* one SDIRK2 step (two sub-solves) then BDF2, driverRedMaxBDF2.m:62-107 */

/* Newton linear solve */
#define RMX_LINSOLVE_LU 0  /* in-block partial-pivot LU == MATLAB `H\g` / lu(H,'vector') (parity path) */
#define RMX_LINSOLVE_PCG 1 /* in-block preconditioned Krylov solve (BiCGStab: the BDF Newton matrix is unsymmetric) with the
                              projected block-Jacobi preconditioner of c++/PCG Solver.cpp:81 + ConstraintJoint.cpp:1236,1455;
                              forward rollouts only (the adjoint tape stores LU factors) */

/* per-rollout status bits (reference prints and continues, driverRedMaxBDF1.m:118-121,135-138,150-153) */
#define RMX_ST_DIVERGED 1  /* 'Newton diverged'            (||dx|| > dxMax) in some step */
#define RMX_ST_MAXITER 2   /* 'Newton did not converge'    (iter >= iterMax) in some step */
#define RMX_ST_LSFAIL 4    /* line search exhausted iterLsMax halvings in some step (silent in reference) */
#define RMX_ST_NAN 8       /* non-finite state produced */
#define RMX_ST_SCHED 16    /* a load-balanced launch gave up waiting for the earlier steps of this rollout (blocks not co-resident,
                              e.g. a shared GPU): the trajectory is NOT valid.  rmx_rollout re-runs such rollouts itself and never
                              returns this bit; callers of rmx_rollout_dev check for it after synchronising their stream */
#define RMX_ST_CHART 32    /* a spherical / Free3D joint left the well-conditioned part of its Euler chart after some step
                              (|det T| <= 0.5: |cos q2| in the charts 7..12, |sin q2| in 1..6); steps after the first such step
                              are the same motion in coordinates the reference would have left.  driverRedMaxBDF2 switches
                              charts there (JointSpherical.m:63-103 -> rmx_rollout_resume); driverRedMaxBDF1 and JointFree3D
                              stop with an error at that point (chart1 is never set) */

/* tau layout for rmx_rollout */
#define RMX_TAU_NONE 0     /* tau == NULL: joint.tau = 0 */
#define RMX_TAU_CONST 1    /* tau is nr x B: constant over the rollout (TaskBDF1PointPos.applyStep) */
#define RMX_TAU_PER_STEP 2 /* tau is nr x nsteps x B */

/*
 * Flattened +redmax scene (what scenesRedMax.m builds with redmax.Scene / BodyCuboid / JointRevolute / JointFixed /
 * JointPrismatic / JointPlanar / JointTranslational / JointFree2D / JointUniversal / JointSpherical / JointFree3D /
 * ForceGroundCuboid /
 * ForcePointPoint / ForceSpringDamper / ForceCable after
 * scene.init(), Scene.m:59-119).  All pointers are host pointers and are
 * copied by rmx_scene_create.
 */
typedef struct rmx_scene_desc {
    int32_t n;               /* number of joints == number of bodies (Scene.m:64) */
    const int32_t* parent;   /* [n] index of parent joint, -1 for a root; parent[j] < j */
    const int32_t* jtype;    /* [n] RMX_JOINT_* */
    const double* E0_pj;     /* [16*n] joint wrt parent joint at q=0   (Joint.setJointTransform, Joint.m:95) */
    const double* E0_ji;     /* [16*n] body wrt joint                  (Body.setBodyTransform, Body.m:46) */
    const double* axis;      /* [3*n]  revolute / prismatic axis, unit length (JointRevolute.m:14, JointPrismatic.m:14);
                                       planar: first in-plane direction (JointPlanar.m:16); ignored otherwise */
    const double* axis2;     /* [3*n]  planar: second in-plane direction (JointPlanar.m:17); may be NULL when the scene
                                       has no planar joint (default [0 1 0]) */
    const double* I_i;       /* [6*n]  diagonal body inertia [Ixx Iyy Izz m m m] (se3.inertiaCuboid, se3.m:366) */
    const double* sides;     /* [3*n]  cuboid side lengths             (BodyCuboid.m:13) */
    const double* stiffness; /* [n] Joint.m:102 */
    const double* damping;   /* [n] Joint.m:108 */
    const double* qRest;     /* [RMX_MAX_JOINT_DOF*n] rest configuration = q at scene.init() (Joint.m:157): entry
                                       [RMX_MAX_JOINT_DOF*j + d] belongs to DOF d of joint j; unused entries ignored */
    const double* qLimL;     /* [n] Joint.m:114  (default -1e8) */
    const double* qLimU;     /* [n] Joint.m:119  (default  1e8) */
    const double* qLimK;     /* [n] Joint.m:124  (default  1e8) */
    const double* qLimD;     /* [n] Joint.m:129  (default  0)   */
    double grav[3];          /* Scene.m:48 */
    int32_t nground;         /* number of ForceGroundCuboid forces (at most one per body) */
    const int32_t* ground_body; /* [nground] body (== joint) index      (ForceGroundCuboid.m:18) */
    const double* ground_E;  /* [16*nground] ground frame, Z-up        (ForceGroundCuboid.m:29) */
    const double* ground_kn; /* [nground] normal stiffness             (ForceGroundCuboid.m:34) */
    const double* ground_kt; /* [nground] tangential stiffness */
    const double* ground_kd; /* [nground] damping                      (ForceGroundCuboid.m:40) */
    const double* ground_mu; /* [nground] friction coefficient         (ForceGroundCuboid.m:45) */
    int32_t npointforce;     /* number of ForcePointPoint / ForceSpringDamper forces (at most RMX_MAX_POINTFORCE) */
    const int32_t* pf_body1; /* [npointforce] body (== joint) index or -1 for the world (ForcePointPoint.m:15) */
    const int32_t* pf_body2; /* [npointforce] */
    const double* pf_x1;     /* [3*npointforce] application point in body-1 (or world) coordinates */
    const double* pf_x2;     /* [3*npointforce] */
    const double* pf_ks;     /* [npointforce] stiffness (ForcePointPoint.m:37) */
    const double* pf_kd;     /* [npointforce] damping   (ForcePointPoint.m:42) */
    const int32_t* pf_kind;  /* [npointforce] RMX_FORCE_*; NULL = all RMX_FORCE_POINTPOINT */
    const double* pf_L;      /* [npointforce] rest length of a spring-damper (ForceSpringDamper.m:31); <= 0 or NULL = the
                                distance of the two points in the initial configuration (ForceSpringDamper.m:38-62) */
    int32_t ncable;          /* number of ForceCable forces; npointforce + ncable <= RMX_MAX_POINTFORCE */
    const int32_t* cable_npts; /* [ncable] points per cable, 2 .. RMX_MAX_CABLE_POINTS (ForceSpringMultiPointGeneric.addBodyPoint) */
    const int32_t* cable_body; /* [RMX_MAX_CABLE_POINTS*ncable] body index of each point or -1 for the world */
    const double* cable_x;   /* [3*RMX_MAX_CABLE_POINTS*ncable] the points in body (or world) coordinates */
    const double* cable_ks;  /* [ncable] ForceCable.m:20 */
    const double* cable_kd;  /* [ncable] ForceCable.m:25 */
    const double* cable_L;   /* [ncable] rest length (ForceCable.m:30); <= 0 or NULL = routed length in the initial configuration */
    const int32_t* chart;    /* [n] Euler chart of a spherical / Free3D joint in the reference's numbering 1..12 (XYX XZX YZY YXY
                                ZXZ ZYZ XYZ XZY YZX YXZ ZXY ZYX, JointSpherical.m:5-16); 0 or NULL = XYZ, the constructor's default */
} rmx_scene_desc;

/* Solver constants hard-coded in the reference's newton() (driverRedMaxBDF1.m:95-98;
 * driverRedMaxAdjointBDF1.m:106-108).  rmx_opts_default() fills the reference's values. */
typedef struct rmx_opts {
    int32_t scheme;          /* RMX_SCHEME_* */
    int32_t nsteps;          /* Scene.m:117  nsteps = ceil(tEnd/h) */
    double h;                /* Scene.m:38 */
    double tol;              /* 1e-9 */
    double dxMax;            /* 1e3 */
    int32_t iterMaxFactor;   /* forward: iterMax = 10*nr; adjoint driver: 5*nr */
    int32_t iterLsMax;       /* 20 */
    int32_t linsolve;        /* RMX_LINSOLVE_* */
    int32_t ngpus;           /* host-pointer entry points only: shard the batch over this many devices (>=1) */
    int32_t tau_mode;        /* RMX_TAU_* */
    int32_t pcg_maxit;       /* RMX_LINSOLVE_PCG: max Krylov iterations per solve (c++/PCG Solver.h:43: 1000; default here 4*nr) */
    double pcg_tol;          /* RMX_LINSOLVE_PCG: relative residual tolerance ||r|| < tol ||r0|| (Solver.h:43, Solver.cpp:137: 1e-6) */
} rmx_opts;

/* TaskBDF1PointPos / TaskBDF2PointPos (matlab-diff/+redmax/TaskBDF1PointPos.m:27-55) */
typedef struct rmx_task_pointpos {
    int32_t body;            /* task body (== joint) index, setBody */
    int32_t reserved;
    double xlocal[3];        /* setPoint */
    double t_target;         /* setTime: objective sampled when |t_target - t| < 1e-6 (TaskBDF1PointPos.m:75) */
    double pscale;           /* setScale */
    double wreg;             /* setWeights(wreg, wpos) */
    double wpos;
} rmx_task_pointpos;

typedef struct rmx_scene rmx_scene;

int rmx_version(void);
const char* rmx_last_error(void);
int rmx_device_count(void);

void rmx_opts_default(rmx_opts* o, int32_t scheme, int32_t adjoint);

/* Scene.init() + flattening.  Replaces the object graph walked by Joint.update/computeJacobian etc. */
int rmx_scene_create(const rmx_scene_desc* d, rmx_scene** out);
void rmx_scene_destroy(rmx_scene* s);
int rmx_scene_nr(const rmx_scene* s); /* redmax.Scene.countR(), Scene.m:411 */
int rmx_scene_nm(const rmx_scene* s); /* redmax.Scene.countM(), Scene.m:398 */

/* Forward rollouts: replaces simLoop(scene) of driverRedMaxBDF1.m:57 / driverRedMaxBDF2.m:57 for B scenes
 * that differ in (q0, qdot0, tau).  q_out/qdot_out[:, k, b] = history(k+1).q/.qdot (Scene.m:136-137).
 * status: B x int32 (RMX_ST_* bits).  iters: 2 x B int32 = total Newton iterations, total residual-only
 * (line-search) evaluations; may be NULL.  qdot_out may be NULL. */
int rmx_rollout(rmx_scene* s, const rmx_opts* o, int64_t B, const double* q0, const double* qdot0,
                const double* tau, double* q_out, double* qdot_out, int32_t* status, int32_t* iters);
/* Device-resident rollouts on G GPUs driven by ONE process: the batch is sharded contiguously, rollouts
 * [B g/G, B (g+1)/G) on devices[g], with no communication while the rollouts run.  q0[g], qdot0[g], tau[g], status[g], iters[g]
 * are device pointers on devices[g] holding that device's SHARD; q_out[g] (and qdot_out[g], optional) point to FULL-size arrays
 * (nr x nsteps x B) on devices[g]: every device integrates its shard into its own slice, then -- gather != 0 -- ONE in-place
 * ncclAllGather per array over NVLink gives every device all trajectories (needs B % G == 0; NCCL is loaded at run time from
 * libnccl.so.2, the call fails with RMX_ENOGPU if it cannot be).  gather == 0: no collective, the other slices stay untouched.
 * Blocking; status bits as rmx_rollout_dev. */
int rmx_rollout_multi_dev(rmx_scene* s, const rmx_opts* o, int32_t G, const int32_t* devices, int64_t B, const double* const* q0,
                          const double* const* qdot0, const double* const* tau, double* const* q_out, double* const* qdot_out,
                          int32_t* const* status, int32_t* const* iters, int32_t gather);

/* Continue rollouts mid-way (host pointers, current device): rollout b runs the steps k_begin[b] .. k_end[b]-1
 * (0 <= k_begin <= k_end <= nsteps; k_end == NULL: to nsteps) from the states the caller provides in q_out / qdot_out --
 * step k_begin-1 is the current state, step k_begin-2 (q0 / qdot0 when k_begin == 1) the BDF2 history joint.q1 / qdot1;
 * k_begin == 0 starts from q0 / qdot0 as rmx_rollout does.  Steps outside [k_begin, k_end) are left as given, status / iters
 * count the steps run by this call only.  This is the hook for what the reference does
 * between steps on the host side of simLoop -- jroot.reparam() (driverRedMaxBDF2.m:112, JointSpherical.m:63-103): the caller
 * re-parameterises the flagged step (RMX_ST_CHART), rebuilds the scene with the new rmx_scene_desc.chart and resumes. */
int rmx_rollout_resume(rmx_scene* s, const rmx_opts* o, int64_t B, const int32_t* k_begin, const int32_t* k_end,
                       const double* q0, const double* qdot0, const double* tau, double* q_out, double* qdot_out,
                       int32_t* status, int32_t* iters);
int rmx_rollout_dev(rmx_scene* s, const rmx_opts* o, int64_t B, const double* q0, const double* qdot0,
                    const double* tau, double* q_out, double* qdot_out, int32_t* status, int32_t* iters,
                    void* cuda_stream);

/* Total Krylov iterations of the last rmx_rollout* call with linsolve = RMX_LINSOLVE_PCG on this scene (summed over
 * rollouts, steps and Newton iterations; the counterpart of SolverDataTracker::num_iterations, c++/PCG Solver.h:19-23). */
int rmx_linsolve_stats(rmx_scene* s, int64_t* krylov_iterations);

/* Objective + gradient: replaces taskObjective(p,scene) of driverRedMaxAdjointBDF1.m:39 / ...BDF2.m:39
 * (scene.reset, task.init, adjoint simLoop with newton:105, saveHistory tape, task.calcStep, task.calcFinal).
 * p: np x B with np == nr (TaskBDF1PointPos.m:18-24); xtarget: 3 x B; P: B; dPdp: np x B;
 * q_out (nr x nsteps x B) optional. */
int rmx_rollout_adjoint(rmx_scene* s, const rmx_opts* o, const rmx_task_pointpos* t, int64_t B,
                        const double* q0, const double* qdot0, const double* p, const double* xtarget,
                        double* P, double* dPdp, double* q_out, int32_t* status);
int rmx_rollout_adjoint_dev(rmx_scene* s, const rmx_opts* o, const rmx_task_pointpos* t, int64_t B,
                            const double* q0, const double* qdot0, const double* p, const double* xtarget,
                            double* P, double* dPdp, double* q_out, int32_t* status, void* cuda_stream);
/* bytes of device scratch (the adjoint tape) rmx_rollout_adjoint* needs for a batch of B */
int64_t rmx_adjoint_tape_bytes(const rmx_scene* s, const rmx_opts* o, int64_t B);

/* Test hook (B = 1, host pointers): one evaluation of the implicit-step residual and its Jacobian,
 *   g = M*dqtmp - cK*f,  H = M - cD*D - cK*K + sum_i dMdq(:,:,i)*dqtmp      (driverRedMaxBDF1.m:173-185)
 * at state (q, qdot) with qdot = beta*(q - const), i.e. cD = cK*beta.  Any of g,H,M,D,f may be NULL.
 * H, M, D are nr x nr column-major; g, f are nr. */
int rmx_eval(rmx_scene* s, const double* q, const double* qdot, const double* dqtmp, const double* tau,
             double cK, double beta, double* g, double* H, double* M, double* D, double* f);

/* Test hook (B = 1, host pointers): the Newton linear system exactly as the forward rollout kernel forms and solves it
 * (the same assembly and in-block LU code path, which rmx_eval's M / D passes do not take): H = dg/dq (nr x nr
 * column-major, before factorisation) and dx = -H \ g  (driverRedMaxBDF1.m:115).  H or dx may be NULL. */
int rmx_eval_newton(rmx_scene* s, const double* q, const double* qdot, const double* dqtmp, const double* tau,
                    double cK, double beta, double* H, double* dx);

/* Test hook: the two operators of the Krylov linear solve (opts.linsolve = RMX_LINSOLVE_PCG, after c++/PCG) at one evaluation
 * point, arguments as rmx_eval.  Hx = H x applied matrix-free -- a root-to-leaves and a leaves-to-root sweep over the joint
 * tree, as the reference applies J x, LHS and J' y (c++/PCG/src/ConstraintJoint.cpp:1090, 1137, 1188) -- and
 * Pinv_x = (J' blkdiag(M_j) J + Pr)^-1 x, the projected block-Jacobi preconditioner (ConstraintJoint.cpp:1236, 1455;
 * notes.pdf Alg. 10).  Either output may be NULL. */
int rmx_eval_krylov(rmx_scene* s, const double* q, const double* qdot, const double* dqtmp, const double* tau, double cK, double beta,
                    const double* x, double* Hx, double* Pinv_x);

/* Test hook (host only): the load-balancing plan of a forward launch -- B rollouts x nsteps steps over `slots` co-resident
 * blocks (McNaughton wrap-around: a rollout is cut at most once; its first part opens one block's list and signals, its second
 * part closes the previous block's list and waits).  seg: 4 ints per segment {rollout, first step, end step, flags: 1 wait,
 * 2 signal}; off: slots + 1 offsets.  Returns the number of segments or a negative RMX_E* code. */
int rmx_debug_schedule(int64_t B, int32_t nsteps, int64_t slots, int32_t seg_capacity, int32_t* seg, int32_t* off);

/* Measurement aid: sustained FP64 rate of the current device in TFLOP/s, from two micro-kernels run on every SM at full
 * occupancy -- independent DFMA chains (2 flop per lane and instruction) and independent DMMA.8x8x4 chains (mma.sync m8n8k4 f64,
 * 512 flop per warp instruction), the two instruction kinds the rollout kernels spend their FP64 time in.  bench.py reports
 * them as the measured FP64 roofline (MEASURED_PEAKS.json holds no FP64 figure).  Either pointer may be NULL. */
int rmx_fp64_probe(double* dfma_tflops, double* dmma_tflops);

/* Page-lock / release a caller-owned host buffer (cudaHostRegister, portable + mapped).  rmx_rollout stores q(t), qdot(t)
 * straight into page-locked output buffers while the kernel runs (no device-to-host copy afterwards); pageable buffers are
 * served through page-locked staging owned by the library (written the same way, copied on by host threads as sub-batches
 * finish).  Registration costs about as much as one copy of the buffer, so it pays for buffers that are reused (MPC loops),
 * not for arrays allocated per call. */
int rmx_host_register(void* p, size_t bytes);
int rmx_host_unregister(void* p);

/* Scene.saveHistory energies (Scene.m:155-160; Joint.m:616, Body.m:167, ForceGroundCuboid.m:156) for B states:
 * T, V: B each. */
int rmx_energies(rmx_scene* s, int64_t B, const double* q, const double* qdot, double* T, double* V);

/* World frames of the bodies, body.E_wi after Joint.update / Body.update (Joint.m:382-434, Body.m:70-80), for B configurations:
 * q is nr x B, E is 4 x 4 x nbodies x B (column-major 4x4 blocks, bodies in the scene's order).  What Scene.draw and the
 * reference's trajectory export read. */
int rmx_body_frames(rmx_scene* s, int64_t B, const double* q, double* E);

#ifdef __cplusplus
}
#endif
#endif /* REDMAX_B200_H */
