/*
 * redmax_oracle_c.c -- compiled twin of oracle/redmax_oracle.py (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C restatement of the forward hot path of sueda/redmax `matlab-diff` in the reference's own DENSE formulation
 * (dense J, Jdot, dJdq, dJdotdq, Mm, Km, Dm; O(n^3) Jacobian derivative recursion; O(n^4) computeValues), one rollout
 * per OpenMP thread.  It exists (i) to be the CPU baseline `bench.py` times on the host cores (`cpu_baseline`,
 * `--impl reference`) and (ii) to check the CUDA path at sizes the NumPy oracle cannot finish in seconds.  It is
 * itself pinned against the NumPy oracle (tests/test_oracle_c.py), which is pinned against the reference's golden
 * energies.  Nothing under redmax_b200/ links or loads this file.
 *
 * Reference lines followed (paths relative to /root/reference/matlab-diff/):
 *   se3.m:11 inv, :44 Ad, :55 ad, :89 brac, :38 Gamma, :111 aaToMat
 *   +redmax/Joint.m:382 update, :437 computeForce, :490-613 computeJacobian (2- and 4-output branches)
 *   +redmax/JointRevolute.m:29 update_ ; JointFixed.m
 *   +redmax/Body.m:70 update, :83 computeMassGrav
 *   +redmax/ForceGroundCuboid.m:54-153 computeValues_
 *   driverRedMaxBDF1.m:57 simLoop, :94 newton, :160 evalBDF1, :190 computeValues
 *   driverRedMaxBDF2.m:57 simLoop, :194 evalSDIRK2a, :228 evalSDIRK2b, :263 evalBDF2
 * MATLAB's `H\g` (LAPACK dgesv) is restated as partial-pivot Gaussian elimination (first maximum on ties, as idamax).
 * Loop-invariant products the reference recomputes inside its loops (J'*Mm, driverRedMaxBDF1.m:222,232,233,240) are
 * formed once, as in the NumPy oracle; everything else keeps the reference's dense shapes.
 *
 * Matrices are row-major here; 4x4 transforms E[16], 6x6 adjoints A[36].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define THRESH 1e-9
#define SDIRK_A 0.29289321881345254 /* (2-sqrt(2))/2, driverRedMaxBDF2.m:75 */

typedef struct {
    int32_t n;
    const int32_t* parent;
    const int32_t* jtype; /* 0 fixed, 1 revolute */
    const double* E0_pj;  /* [16n] row-major */
    const double* E0_ji;  /* [16n] row-major */
    const double* axis;   /* [3n] */
    const double* I_i;    /* [6n] */
    const double* sides;  /* [3n] */
    const double* stiffness;
    const double* damping;
    const double* qRest;
    const double* qLimL;
    const double* qLimU;
    const double* qLimK;
    const double* qLimD;
    double grav[3];
    const int32_t* has_ground; /* [n] */
    const double* ground_E;    /* [16n] row-major (only where has_ground) */
    const double* ground_kn;
    const double* ground_kt;
    const double* ground_kd;
    const double* ground_mu;
} oc_desc;

/* ------------------------------------------------------------------ small dense helpers */
static void mm(int m, int k, int n, const double* A, const double* B, double* C) { /* C = A(mxk) B(kxn) */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) C[i * n + j] = 0.0;
    for (int i = 0; i < m; ++i)
        for (int l = 0; l < k; ++l) {
            const double a = A[i * k + l];
            if (a == 0.0) continue;
            for (int j = 0; j < n; ++j) C[i * n + j] += a * B[l * n + j];
        }
}
static void mm_dense(int m, int k, int n, const double* A, const double* B, double* C) { /* no zero skipping */
    for (int i = 0; i < m; ++i) {
        double* c = C + (size_t)i * n;
        for (int j = 0; j < n; ++j) c[j] = 0.0;
        for (int l = 0; l < k; ++l) {
            const double a = A[(size_t)i * k + l];
            const double* b = B + (size_t)l * n;
            for (int j = 0; j < n; ++j) c[j] += a * b[j];
        }
    }
}
static void mtm_dense(int k, int m, int n, const double* A, const double* B, double* C) { /* C = A'(m x k) B(k x n), A is k x m */
    for (int i = 0; i < m * n; ++i) C[i] = 0.0;
    for (int l = 0; l < k; ++l) {
        const double* a = A + (size_t)l * m;
        const double* b = B + (size_t)l * n;
        for (int i = 0; i < m; ++i) {
            const double ai = a[i];
            double* c = C + (size_t)i * n;
            for (int j = 0; j < n; ++j) c[j] += ai * b[j];
        }
    }
}
static void mv(int m, int n, const double* A, const double* x, double* y) {
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += A[(size_t)i * n + j] * x[j];
        y[i] = s;
    }
}
static void mtv(int m, int n, const double* A, const double* x, double* y) { /* y = A' x, A m x n */
    for (int j = 0; j < n; ++j) y[j] = 0.0;
    for (int i = 0; i < m; ++i) {
        const double xi = x[i];
        for (int j = 0; j < n; ++j) y[j] += A[(size_t)i * n + j] * xi;
    }
}
static void eye4(double* E) {
    memset(E, 0, 16 * sizeof(double));
    E[0] = E[5] = E[10] = E[15] = 1.0;
}
static void se3_inv(const double* E, double* Ei) { /* se3.m:11 */
    eye4(Ei);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Ei[4 * i + j] = E[4 * j + i];
    for (int i = 0; i < 3; ++i) {
        double s = 0.0;
        for (int j = 0; j < 3; ++j) s += E[4 * j + i] * E[4 * j + 3];
        Ei[4 * i + 3] = -s;
    }
}
static void brac(const double* x, double* S) { /* se3.m:89 */
    S[0] = 0; S[1] = -x[2]; S[2] = x[1];
    S[3] = x[2]; S[4] = 0; S[5] = -x[0];
    S[6] = -x[1]; S[7] = x[0]; S[8] = 0;
}
static void se3_Ad(const double* E, double* A) { /* se3.m:44 : [R 0; [p]R R] */
    memset(A, 0, 36 * sizeof(double));
    double R[9], P[9], PR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = E[4 * i + j];
    double p[3] = {E[3], E[7], E[11]};
    brac(p, P);
    mm(3, 3, 3, P, R, PR);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[6 * i + j] = R[3 * i + j];
            A[6 * (i + 3) + j + 3] = R[3 * i + j];
            A[6 * (i + 3) + j] = PR[3 * i + j];
        }
}
static void se3_ad(const double* phi, double* a) { /* se3.m:55 : [W 0; [v] W] */
    memset(a, 0, 36 * sizeof(double));
    double W[9], Vv[9];
    brac(phi, W);
    brac(phi + 3, Vv);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            a[6 * i + j] = W[3 * i + j];
            a[6 * (i + 3) + j + 3] = W[3 * i + j];
            a[6 * (i + 3) + j] = Vv[3 * i + j];
        }
}
static void aaToMat(const double* axis, double angle, double* R) { /* se3.m:111-176 */
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    double ax = axis[0], ay = axis[1], az = axis[2];
    double mag = sqrt(ax * ax + ay * ay + az * az);
    if (!(mag > THRESH)) return;
    mag = 1.0 / mag;
    ax *= mag; ay *= mag; az *= mag;
    double s, c;
    if (fabs(ax) < THRESH && fabs(ay) < THRESH) {
        if (az < 0) angle = -angle;
        s = sin(angle); c = cos(angle);
        R[0] = c; R[1] = -s; R[3] = s; R[4] = c;
    } else if (fabs(ay) < THRESH && fabs(az) < THRESH) {
        if (ax < 0) angle = -angle;
        s = sin(angle); c = cos(angle);
        R[4] = c; R[5] = -s; R[7] = s; R[8] = c;
    } else if (fabs(az) < THRESH && fabs(ax) < THRESH) {
        if (ay < 0) angle = -angle;
        s = sin(angle); c = cos(angle);
        R[0] = c; R[2] = s; R[6] = -s; R[8] = c;
    } else {
        s = sin(angle); c = cos(angle);
        const double t = 1.0 - c, xz = ax * az, xy = ax * ay, yz = ay * az;
        R[0] = t * ax * ax + c; R[1] = t * xy - s * az; R[2] = t * xz + s * ay;
        R[3] = t * xy + s * az; R[4] = t * ay * ay + c; R[5] = t * yz - s * ax;
        R[6] = t * xz - s * ay; R[7] = t * yz + s * ax; R[8] = t * az * az + c;
    }
}

/* ------------------------------------------------------------------ per-rollout workspace */
typedef struct {
    /* constants derived once */
    int n, nr, nm;
    int* idxR; /* [n] reduced index or -1 */
    int* idxM; /* [n] first maximal index */
    double *E0_jp, *E0_ij, *A0_ij;
    /* joint / body state */
    double *q, *qdot, *q0, *qdot0, *q1, *qdot1, *tau;                 /* per reduced dof [nr] (reference numbering) */
    double *Q, *invQ, *E_pj, *E_jp, *E_wj, *E_wi;                     /* [16n] */
    double *A, *invA, *Adot, *dAdq, *dAdotdq, *A_jp;                  /* [36n] */
    double *V, *phi;                                                  /* [6n] */
    /* dense work */
    double *J, *Jdot;           /* nm x nr */
    double *dJdq, *dJdotdq;     /* [nr][nm x nr] : slice i is contiguous */
    double *Mm, *Km, *Dm;       /* nm x nm */
    double *fm, *fr, *Kr, *Dr;  /* nm, nr, nr x nr, nr x nr */
    double *JtMm, *JtX, *M, *K, *D, *H, *f, *g, *tmpa, *tmpb, *dMdq; /* dMdq [nr][nr x nr] */
    double *v1, *v2, *v3, *x, *dx, *x0, *g0, *LU;
} oc_work;

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

static oc_work* work_create(const oc_desc* d) {
    oc_work* w = (oc_work*)calloc(1, sizeof(oc_work));
    const int n = d->n;
    w->n = n;
    w->idxR = (int*)calloc(n, sizeof(int));
    w->idxM = (int*)calloc(n, sizeof(int));
    int nr = 0, nm = 0;
    for (int j = n - 1; j >= 0; --j) { /* Scene.m:69-71: countDofs for i = n:-1:1 */
        w->idxR[j] = d->jtype[j] == 1 ? nr++ : -1;
        w->idxM[j] = nm;
        nm += 6;
    }
    w->nr = nr;
    w->nm = nm;
    w->E0_jp = dalloc(16 * n); w->E0_ij = dalloc(16 * n); w->A0_ij = dalloc(36 * n);
    for (int j = 0; j < n; ++j) {
        se3_inv(d->E0_pj + 16 * j, w->E0_jp + 16 * j);
        se3_inv(d->E0_ji + 16 * j, w->E0_ij + 16 * j);
        se3_Ad(w->E0_ij + 16 * j, w->A0_ij + 36 * j);
    }
    w->q = dalloc(nr); w->qdot = dalloc(nr); w->q0 = dalloc(nr); w->qdot0 = dalloc(nr);
    w->q1 = dalloc(nr); w->qdot1 = dalloc(nr); w->tau = dalloc(nr);
    w->Q = dalloc(16 * n); w->invQ = dalloc(16 * n); w->E_pj = dalloc(16 * n); w->E_jp = dalloc(16 * n);
    w->E_wj = dalloc(16 * n); w->E_wi = dalloc(16 * n);
    w->A = dalloc(36 * n); w->invA = dalloc(36 * n); w->Adot = dalloc(36 * n); w->dAdq = dalloc(36 * n);
    w->dAdotdq = dalloc(36 * n); w->A_jp = dalloc(36 * n);
    w->V = dalloc(6 * n); w->phi = dalloc(6 * n);
    w->J = dalloc((size_t)nm * nr); w->Jdot = dalloc((size_t)nm * nr);
    w->dJdq = dalloc((size_t)nm * nr * nr); w->dJdotdq = dalloc((size_t)nm * nr * nr);
    w->Mm = dalloc((size_t)nm * nm); w->Km = dalloc((size_t)nm * nm); w->Dm = dalloc((size_t)nm * nm);
    w->fm = dalloc(nm); w->fr = dalloc(nr); w->Kr = dalloc((size_t)nr * nr); w->Dr = dalloc((size_t)nr * nr);
    w->JtMm = dalloc((size_t)nr * nm); w->JtX = dalloc((size_t)nr * nm);
    w->M = dalloc((size_t)nr * nr); w->K = dalloc((size_t)nr * nr); w->D = dalloc((size_t)nr * nr);
    w->H = dalloc((size_t)nr * nr); w->LU = dalloc((size_t)nr * nr);
    w->f = dalloc(nr); w->g = dalloc(nr); w->tmpa = dalloc((size_t)nr * nr > (size_t)nm ? (size_t)nr * nr : nm);
    w->tmpb = dalloc((size_t)nr * nr > (size_t)nm ? (size_t)nr * nr : nm);
    w->dMdq = dalloc((size_t)nr * nr * nr);
    w->v1 = dalloc(nm); w->v2 = dalloc(nm); w->v3 = dalloc(nm);
    w->x = dalloc(nr); w->dx = dalloc(nr); w->x0 = dalloc(nr); w->g0 = dalloc(nr);
    return w;
}
static void work_destroy(oc_work* w) {
    double** p[] = {&w->E0_jp, &w->E0_ij, &w->A0_ij, &w->q, &w->qdot, &w->q0, &w->qdot0, &w->q1, &w->qdot1, &w->tau, &w->Q,
                    &w->invQ, &w->E_pj, &w->E_jp, &w->E_wj, &w->E_wi, &w->A, &w->invA, &w->Adot, &w->dAdq, &w->dAdotdq,
                    &w->A_jp, &w->V, &w->phi, &w->J, &w->Jdot, &w->dJdq, &w->dJdotdq, &w->Mm, &w->Km, &w->Dm, &w->fm, &w->fr,
                    &w->Kr, &w->Dr, &w->JtMm, &w->JtX, &w->M, &w->K, &w->D, &w->H, &w->LU, &w->f, &w->g, &w->tmpa, &w->tmpb,
                    &w->dMdq, &w->v1, &w->v2, &w->v3, &w->x, &w->dx, &w->x0, &w->g0};
    for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); ++i) free(*p[i]);
    free(w->idxR);
    free(w->idxM);
    free(w);
}

/* ------------------------------------------------------------------ Joint.update (Joint.m:382) + JointRevolute.update_ + Body.update */
static void scene_update(const oc_desc* d, oc_work* w) {
    const int n = w->n;
    for (int j = 0; j < n; ++j) {
        double* Q = w->Q + 16 * j;
        double* A = w->A + 36 * j;
        double* Adot = w->Adot + 36 * j;
        double* dAdq = w->dAdq + 36 * j;
        double* dAdotdq = w->dAdotdq + 36 * j;
        eye4(Q);
        memset(Adot, 0, 36 * sizeof(double));
        memset(dAdq, 0, 36 * sizeof(double));
        memset(dAdotdq, 0, 36 * sizeof(double));
        if (d->jtype[j] == 1) { /* JointRevolute.m:29-53 */
            const double q = w->q[w->idxR[j]], qdot = w->qdot[w->idxR[j]];
            const double* a = d->axis + 3 * j;
            double R[9], ab[9], dRdq[9], d2[9];
            aaToMat(a, q, R);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) Q[4 * r + c] = R[3 * r + c];
            brac(a, ab);
            mm(3, 3, 3, R, ab, dRdq);
            mm(3, 3, 3, dRdq, ab, d2);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    const double rd = dRdq[3 * r + c] * qdot;
                    Adot[6 * r + c] = rd;
                    Adot[6 * (r + 3) + c + 3] = rd;
                    dAdq[6 * r + c] = dRdq[3 * r + c];
                    dAdq[6 * (r + 3) + c + 3] = dRdq[3 * r + c];
                    const double t2 = d2[3 * r + c] * qdot;
                    dAdotdq[6 * r + c] = t2;
                    dAdotdq[6 * (r + 3) + c + 3] = t2;
                }
        }
        se3_Ad(Q, A);
        se3_inv(Q, w->invQ + 16 * j);
        se3_Ad(w->invQ + 16 * j, w->invA + 36 * j);
        mm(4, 4, 4, d->E0_pj + 16 * j, Q, w->E_pj + 16 * j);
        se3_inv(w->E_pj + 16 * j, w->E_jp + 16 * j);
        se3_Ad(w->E_jp + 16 * j, w->A_jp + 36 * j);
        const int p = d->parent[j];
        if (p < 0)
            memcpy(w->E_wj + 16 * j, w->E_pj + 16 * j, 16 * sizeof(double));
        else
            mm(4, 4, 4, w->E_wj + 16 * p, w->E_pj + 16 * j, w->E_wj + 16 * j);
        double* V = w->V + 6 * j;
        for (int i = 0; i < 6; ++i) V[i] = 0.0;
        if (d->jtype[j] == 1) {
            const double qdot = w->qdot[w->idxR[j]];
            for (int i = 0; i < 3; ++i) V[i] = d->axis[3 * j + i] * qdot; /* S = [a;0] */
        }
        if (p >= 0) {
            double t[6];
            mv(6, 6, w->A_jp + 36 * j, w->V + 6 * p, t);
            for (int i = 0; i < 6; ++i) V[i] += t[i];
        }
        /* Body.update (Body.m:70) */
        mm(4, 4, 4, w->E_wj + 16 * j, d->E0_ji + 16 * j, w->E_wi + 16 * j);
        mv(6, 6, w->A0_ij + 36 * j, V, w->phi + 6 * j);
    }
}

/* ------------------------------------------------------------------ Joint.computeJacobian (Joint.m:490-613) */
static void ad_from(const double* E0_BiJi, const double* invQ, const double* E0_JiBp, double* A_BiBp, double* Aleft,
                    double* Aright) {
    double T1[16], T2[16], T3[16];
    mm(4, 4, 4, E0_BiJi, invQ, T1);
    mm(4, 4, 4, T1, E0_JiBp, T2);
    se3_Ad(T2, A_BiBp);
    se3_Ad(T1, Aleft);
    for (int i = 0; i < 36; ++i) Aleft[i] = -Aleft[i];
    mm(4, 4, 4, invQ, E0_JiBp, T3);
    se3_Ad(T3, Aright);
}

static void compute_jacobian(const oc_desc* d, oc_work* w, int deriv) {
    const int n = w->n, nr = w->nr, nm = w->nm;
    double* J = w->J;
    double* Jd = w->Jdot;
    memset(J, 0, sizeof(double) * nm * nr);
    memset(Jd, 0, sizeof(double) * nm * nr);
    if (deriv) {
        memset(w->dJdq, 0, sizeof(double) * nm * nr * nr);
        memset(w->dJdotdq, 0, sizeof(double) * nm * nr * nr);
    }
    const size_t sl = (size_t)nm * nr; /* slice stride */
    for (int j = 0; j < n; ++j) {
        const int i0 = w->idxM[j];
        const int ri = w->idxR[j];
        const double* A0 = w->A0_ij + 36 * j;
        if (ri >= 0) { /* J(idxmI,idxrI) = A0_BiJi*S ; Sdot = 0, dSdq = 0 for a revolute joint */
            for (int r = 0; r < 6; ++r) {
                double s = 0.0;
                for (int c = 0; c < 3; ++c) s += A0[6 * r + c] * d->axis[3 * j + c];
                J[(size_t)(i0 + r) * nr + ri] = s;
            }
        }
        const int p = d->parent[j];
        if (p < 0) continue;
        const int p0 = w->idxM[p];
        double E0_JiBp[16], A_BiBp[36], Aleft[36], Aright[36], Adot_BiBp[36], T[36], T2[36];
        mm(4, 4, 4, w->E0_jp + 16 * j, d->E0_ji + 16 * p, E0_JiBp);
        ad_from(w->E0_ij + 16 * j, w->invQ + 16 * j, E0_JiBp, A_BiBp, Aleft, Aright);
        mm(6, 6, 6, Aleft, w->Adot + 36 * j, T);
        mm(6, 6, 6, T, Aright, Adot_BiBp);
        double dAdq_BiBp[36], dAdotdq_BiBp[36];
        if (deriv && ri >= 0) { /* Joint.m:573-580 */
            const double* dAdq = w->dAdq + 36 * j;
            const double* invA = w->invA + 36 * j;
            const double* Adot = w->Adot + 36 * j;
            double t1[36], t2[36], u[36];
            mm(6, 6, 6, dAdq, invA, T);
            mm(6, 6, 6, T, Adot, t1);
            mm(6, 6, 6, Adot, invA, T);
            mm(6, 6, 6, T, dAdq, t2);
            mm(6, 6, 6, Aleft, dAdq, T);
            mm(6, 6, 6, T, Aright, dAdq_BiBp);
            for (int i = 0; i < 36; ++i) u[i] = w->dAdotdq[36 * j + i] - t1[i] - t2[i];
            mm(6, 6, 6, Aleft, u, T2);
            mm(6, 6, 6, T2, Aright, dAdotdq_BiBp);
        }
        for (int a = p; a >= 0; a = d->parent[a]) {
            const int ra = w->idxR[a];
            if (ra < 0) continue;
            double JPA[6], JdPA[6], o1[6], o2[6], o3[6];
            for (int r = 0; r < 6; ++r) {
                JPA[r] = J[(size_t)(p0 + r) * nr + ra];
                JdPA[r] = Jd[(size_t)(p0 + r) * nr + ra];
            }
            mv(6, 6, A_BiBp, JPA, o1);
            mv(6, 6, A_BiBp, JdPA, o2);
            mv(6, 6, Adot_BiBp, JPA, o3);
            for (int r = 0; r < 6; ++r) {
                J[(size_t)(i0 + r) * nr + ra] = o1[r];
                Jd[(size_t)(i0 + r) * nr + ra] = o2[r] + o3[r];
            }
            if (!deriv) continue;
            if (ri >= 0) { /* Joint.m:585-590 */
                double* dj = w->dJdq + sl * ri;
                double* djd = w->dJdotdq + sl * ri;
                mv(6, 6, dAdq_BiBp, JPA, o1);
                mv(6, 6, dAdq_BiBp, JdPA, o2);
                mv(6, 6, dAdotdq_BiBp, JPA, o3);
                for (int r = 0; r < 6; ++r) {
                    dj[(size_t)(i0 + r) * nr + ra] = o1[r];
                    djd[(size_t)(i0 + r) * nr + ra] = o2[r] + o3[r];
                }
            }
            for (int k = p; k >= 0; k = d->parent[k]) { /* Joint.m:591-604 */
                const int rk = w->idxR[k];
                if (rk < 0) continue;
                double* dj = w->dJdq + sl * rk;
                double* djd = w->dJdotdq + sl * rk;
                double a1[6], a2[6];
                for (int r = 0; r < 6; ++r) {
                    a1[r] = dj[(size_t)(p0 + r) * nr + ra];
                    a2[r] = djd[(size_t)(p0 + r) * nr + ra];
                }
                mv(6, 6, A_BiBp, a1, o1);
                mv(6, 6, A_BiBp, a2, o2);
                mv(6, 6, Adot_BiBp, a1, o3);
                for (int r = 0; r < 6; ++r) {
                    dj[(size_t)(i0 + r) * nr + ra] = o1[r];
                    djd[(size_t)(i0 + r) * nr + ra] = o2[r] + o3[r];
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ Body.computeMassGrav (Body.m:83-135) */
static void compute_mass_grav(const oc_desc* d, oc_work* w, int deriv) {
    const int n = w->n, nm = w->nm;
    for (int j = 0; j < n; ++j) {
        const int i0 = w->idxM[j];
        const double* I = d->I_i + 6 * j;
        const double* phi = w->phi + 6 * j;
        for (int r = 0; r < 6; ++r) w->Mm[(size_t)(i0 + r) * nm + i0 + r] = I[r];
        double ad[36], adtM[36], Mphi[6], fcor[6];
        se3_ad(phi, ad);
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) adtM[6 * r + c] = ad[6 * c + r] * I[c]; /* ad' * M_i */
        (void)Mphi;
        mv(6, 6, adtM, phi, fcor);
        const double* E = w->E_wi + 16 * j;
        double gi[3];
        for (int r = 0; r < 3; ++r) gi[r] = E[0 + r] * d->grav[0] + E[4 + r] * d->grav[1] + E[8 + r] * d->grav[2]; /* R' g */
        const double mass = I[3];
        double fgrav[6] = {0, 0, 0, mass * gi[0], mass * gi[1], mass * gi[2]};
        for (int r = 0; r < 6; ++r) w->fm[i0 + r] += fcor[r] + fgrav[r];
        if (deriv) {
            double S[9];
            brac(fgrav + 3, S);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) w->Km[(size_t)(i0 + 3 + r) * nm + i0 + c] += S[3 * r + c];
            double Iw[3] = {I[0] * phi[0], I[1] * phi[1], I[2] * phi[2]};
            double mvv[3] = {mass * phi[3], mass * phi[4], mass * phi[5]};
            double blk[36];
            memset(blk, 0, sizeof(blk));
            for (int k = 0; k < 3; ++k) { /* columns [e_k] Iw, [e_k] mv */
                double e[3] = {0, 0, 0}, Sk[9], c1[3], c2[3];
                e[k] = 1.0;
                brac(e, Sk);
                mv(3, 3, Sk, Iw, c1);
                mv(3, 3, Sk, mvv, c2);
                for (int r = 0; r < 3; ++r) {
                    blk[6 * r + k] = c1[r];
                    blk[6 * r + 3 + k] = c2[r];
                    blk[6 * (r + 3) + k] = c2[r];
                }
            }
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) w->Dm[(size_t)(i0 + r) * nm + i0 + c] += adtM[6 * r + c] - blk[6 * r + c];
        }
    }
}

/* ------------------------------------------------------------------ Joint.computeForce (Joint.m:437-487) */
static void compute_joint_force(const oc_desc* d, oc_work* w, int deriv) {
    const int nr = w->nr;
    for (int j = 0; j < w->n; ++j) {
        const int r = w->idxR[j];
        if (r < 0) continue;
        const double q = w->q[r], qdot = w->qdot[r];
        w->fr[r] += w->tau[r] + d->stiffness[j] * (d->qRest[j] - q) - d->damping[j] * qdot;
        const double hitL = q < d->qLimL[j] ? 1.0 : 0.0, hitU = q > d->qLimU[j] ? 1.0 : 0.0;
        w->fr[r] += hitL * (d->qLimK[j] * (d->qLimL[j] - q) - d->qLimD[j] * qdot);
        w->fr[r] += hitU * (d->qLimK[j] * (d->qLimU[j] - q) - d->qLimD[j] * qdot);
        if (deriv) {
            w->Kr[(size_t)r * nr + r] -= d->stiffness[j];
            w->Dr[(size_t)r * nr + r] -= d->damping[j];
            w->Kr[(size_t)r * nr + r] -= hitL * d->qLimK[j];
            w->Kr[(size_t)r * nr + r] -= hitU * d->qLimK[j];
            w->Dr[(size_t)r * nr + r] -= hitL * d->qLimD[j];
            w->Dr[(size_t)r * nr + r] -= hitU * d->qLimD[j];
        }
    }
}

/* ------------------------------------------------------------------ ForceGroundCuboid.computeValues_ (ForceGroundCuboid.m:54-153) */
static void gt_mul(const double* xl, const double* X, int cols, double* out) { /* out(6 x cols) = G' X(3 x cols), G = [brac(xl)' I] */
    double S[9];
    brac(xl, S); /* G' = [brac(xl); I] */
    for (int c = 0; c < cols; ++c) {
        for (int r = 0; r < 3; ++r) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += S[3 * r + k] * X[k * cols + c];
            out[r * cols + c] = s;
            out[(r + 3) * cols + c] = X[r * cols + c];
        }
    }
}
static void ground_force(const oc_desc* d, oc_work* w, int j, int deriv) {
    const int nm = w->nm, i0 = w->idxM[j];
    const double* Eg = d->ground_E + 16 * j;
    const double kn = d->ground_kn[j], kt = d->ground_kt[j], kd = d->ground_kd[j], mu = d->ground_mu[j];
    const double xg[3] = {Eg[3], Eg[7], Eg[11]}, ng[3] = {Eg[2], Eg[6], Eg[10]};
    double N[9], T[9], R[9], Rt[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            N[3 * r + c] = ng[r] * ng[c];
            T[3 * r + c] = (r == c ? 1.0 : 0.0) - N[3 * r + c];
            R[3 * r + c] = w->E_wi[16 * j + 4 * r + c];
            Rt[3 * c + r] = R[3 * r + c];
        }
    const double p[3] = {w->E_wi[16 * j + 3], w->E_wi[16 * j + 7], w->E_wi[16 * j + 11]};
    const double* phi = w->phi + 6 * j;
    double RtN[9], RNR[9], RtT[9], B[9], pd[3], v3[3], pxg[9];
    mm(3, 3, 3, Rt, N, RtN);
    mm(3, 3, 3, RtN, R, RNR);
    mm(3, 3, 3, Rt, T, RtT);
    mm(3, 3, 3, RtT, R, B);
    for (int r = 0; r < 3; ++r) pd[r] = p[r] - xg[r];
    mv(3, 3, RtN, pd, v3);
    brac(v3, pxg);
    double Kacc[36], Dacc[36], facc[6];
    memset(Kacc, 0, sizeof(Kacc));
    memset(Dacc, 0, sizeof(Dacc));
    memset(facc, 0, sizeof(facc));
    for (int ci = 0; ci < 8; ++ci) {
        const double xl[3] = {((ci & 4) ? 0.5 : -0.5) * d->sides[3 * j], ((ci & 2) ? 0.5 : -0.5) * d->sides[3 * j + 1],
                              ((ci & 1) ? 0.5 : -0.5) * d->sides[3 * j + 2]};
        double xw[3];
        mv(3, 3, R, xl, xw);
        for (int r = 0; r < 3; ++r) xw[r] += p[r];
        const double dd = ng[0] * (xw[0] - xg[0]) + ng[1] * (xw[1] - xg[1]) + ng[2] * (xw[2] - xg[2]);
        if (dd > 0) continue;
        double xlb[9], Gphi[3], vw[3], Nv[3], fc[3], Rtf[3], w6[6];
        brac(xl, xlb);
        /* G*phi = brac(xl)'*w + v */
        for (int r = 0; r < 3; ++r) Gphi[r] = xlb[0 * 3 + r] * phi[0] + xlb[1 * 3 + r] * phi[1] + xlb[2 * 3 + r] * phi[2] + phi[3 + r];
        mv(3, 3, R, Gphi, vw);
        mv(3, 3, N, vw, Nv);
        for (int r = 0; r < 3; ++r) fc[r] = -kn * ng[r] * dd - kd * Nv[r];
        mv(3, 3, Rt, fc, Rtf);
        gt_mul(xl, Rtf, 1, w6);
        for (int r = 0; r < 6; ++r) facc[r] += w6[r];
        if (deriv) {
            double RNRxl[3], RNRG[3], Gb[9], t1[9], t2[9], RX[9], RGb[9], X[18], o[36];
            mv(3, 3, RNR, xl, RNRxl);
            mv(3, 3, RNR, Gphi, RNRG);
            brac(Gphi, Gb);
            mm(3, 3, 3, RNR, xlb, RX);
            mm(3, 3, 3, RNR, Gb, RGb);
            for (int k = 0; k < 3; ++k) {
                double e[3] = {0, 0, 0}, Sk[9], c1[3], c2[3];
                e[k] = 1.0;
                brac(e, Sk);
                mv(3, 3, Sk, RNRxl, c1);
                mv(3, 3, Sk, RNRG, c2);
                for (int r = 0; r < 3; ++r) {
                    t1[3 * r + k] = -c1[r] - RX[3 * r + k] + pxg[3 * r + k];
                    t2[3 * r + k] = -c2[r] - RGb[3 * r + k];
                }
            }
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    X[6 * r + c] = kn * t1[3 * r + c] + kd * t2[3 * r + c];
                    X[6 * r + 3 + c] = kn * RNR[3 * r + c];
                }
            gt_mul(xl, X, 6, o);
            for (int i = 0; i < 36; ++i) Kacc[i] -= o[i];
            /* Dm -= kd G' RNR G ; G = [xlb' I] */
            double RG[18];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double s = 0.0;
                    for (int k = 0; k < 3; ++k) s += RNR[3 * r + k] * xlb[3 * c + k];
                    RG[6 * r + c] = s;
                    RG[6 * r + 3 + c] = RNR[3 * r + c];
                }
            gt_mul(xl, RG, 6, o);
            for (int i = 0; i < 36; ++i) Dacc[i] -= kd * o[i];
        }
        if (mu == 0) continue;
        double a[3];
        mv(3, 3, T, vw, a);
        const double anorm = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        if (mu * fabs(kn * dd) > kt * anorm) { /* static, :121-133 */
            double fs[3] = {-kt * a[0], -kt * a[1], -kt * a[2]};
            mv(3, 3, Rt, fs, Rtf);
            gt_mul(xl, Rtf, 1, w6);
            for (int r = 0; r < 6; ++r) facc[r] += w6[r];
            if (deriv) {
                double BG[18], o[36], X[18];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        double s = 0.0;
                        for (int k = 0; k < 3; ++k) s += B[3 * r + k] * xlb[3 * c + k];
                        BG[6 * r + c] = s;
                        BG[6 * r + 3 + c] = B[3 * r + c];
                    }
                gt_mul(xl, BG, 6, o);
                for (int i = 0; i < 36; ++i) Dacc[i] += -kt * o[i];
                memset(X, 0, sizeof(X));
                for (int k = 0; k < 3; ++k) {
                    double e[3] = {0, 0, 0}, Sk[9], BS[9], SB[9], col[3];
                    e[k] = 1.0;
                    brac(e, Sk);
                    mm(3, 3, 3, B, Sk, BS);
                    mm(3, 3, 3, Sk, B, SB);
                    for (int i = 0; i < 9; ++i) BS[i] -= SB[i];
                    mv(3, 3, BS, Gphi, col);
                    for (int r = 0; r < 3; ++r) X[6 * r + k] = col[r];
                }
                gt_mul(xl, X, 6, o);
                for (int i = 0; i < 36; ++i) Kacc[i] += -kt * o[i];
            }
        } else { /* dynamic, :134-150 */
            const double mukn = mu * kn;
            double t[3] = {a[0] / anorm, a[1] / anorm, a[2] / anorm};
            double fd[3] = {-mukn * dd * t[0], -mukn * dd * t[1], -mukn * dd * t[2]};
            mv(3, 3, Rt, fd, Rtf);
            gt_mul(xl, Rtf, 1, w6);
            for (int r = 0; r < 6; ++r) facc[r] += w6[r];
            if (deriv) {
                const double a2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], an3 = anorm * anorm * anorm;
                double Am[9], RtA[9], RtAT[9], C[9], CG[18], o[36], Rtt[3], X[18], Gb[9], CGb[9];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) Am[3 * r + c] = ((r == c ? a2 : 0.0) - a[r] * a[c]) / an3;
                mm(3, 3, 3, Rt, Am, RtA);
                mm(3, 3, 3, RtA, T, RtAT);
                mm(3, 3, 3, RtAT, R, C); /* R' A T R */
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        double s = 0.0;
                        for (int k = 0; k < 3; ++k) s += C[3 * r + k] * xlb[3 * c + k];
                        CG[6 * r + c] = s;
                        CG[6 * r + 3 + c] = C[3 * r + c];
                    }
                gt_mul(xl, CG, 6, o);
                for (int i = 0; i < 36; ++i) Dacc[i] += -mukn * dd * o[i];
                mv(3, 3, Rt, t, Rtt);
                brac(Gphi, Gb);
                mm(3, 3, 3, C, Gb, CGb);
                double ngR[3], ngRG[6];
                mtv(3, 3, R, ng, ngR); /* ng' R */
                for (int c = 0; c < 3; ++c) {
                    double s = 0.0;
                    for (int k = 0; k < 3; ++k) s += ngR[k] * xlb[3 * c + k];
                    ngRG[c] = s;
                    ngRG[3 + c] = ngR[c];
                }
                memset(X, 0, sizeof(X));
                for (int k = 0; k < 3; ++k) {
                    double e[3] = {0, 0, 0}, Sk[9], col[3];
                    e[k] = 1.0;
                    brac(e, Sk);
                    mv(3, 3, Sk, Rtt, col);
                    for (int r = 0; r < 3; ++r) X[6 * r + k] = -dd * col[r] - dd * CGb[3 * r + k];
                }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 6; ++c) X[6 * r + c] += Rtt[r] * ngRG[c];
                gt_mul(xl, X, 6, o);
                for (int i = 0; i < 36; ++i) Kacc[i] += -mukn * o[i];
            }
        }
    }
    for (int r = 0; r < 6; ++r) w->fm[i0 + r] += facc[r];
    if (deriv)
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                w->Km[(size_t)(i0 + r) * nm + i0 + c] += Kacc[6 * r + c];
                w->Dm[(size_t)(i0 + r) * nm + i0 + c] += Dacc[6 * r + c];
            }
}

/* ------------------------------------------------------------------ computeValues (driverRedMaxBDF1.m:190-243) */
static void compute_values(const oc_desc* d, oc_work* w, int deriv) {
    const int nr = w->nr, nm = w->nm;
    memset(w->Mm, 0, sizeof(double) * nm * nm);
    memset(w->fm, 0, sizeof(double) * nm);
    memset(w->fr, 0, sizeof(double) * nr);
    if (deriv) {
        memset(w->Km, 0, sizeof(double) * nm * nm);
        memset(w->Dm, 0, sizeof(double) * nm * nm);
        memset(w->Kr, 0, sizeof(double) * nr * nr);
        memset(w->Dr, 0, sizeof(double) * nr * nr);
    }
    compute_jacobian(d, w, deriv);
    compute_mass_grav(d, w, deriv);
    compute_joint_force(d, w, deriv);
    if (d->has_ground)
        for (int j = 0; j < w->n; ++j)
            if (d->has_ground[j]) ground_force(d, w, j, deriv);
    const double* J = w->J;
    const double* Jd = w->Jdot;
    mtm_dense(nm, nr, nm, J, w->Mm, w->JtMm);    /* JtMm = J' Mm */
    mm_dense(nr, nm, nr, w->JtMm, J, w->M);      /* M = J' Mm J */
    mv(nm, nr, Jd, w->qdot, w->v1);              /* Jdot qdot */
    mv(nr, nm, w->JtMm, w->v1, w->f);            /* -fqvv */
    mtv(nm, nr, J, w->fm, w->tmpa);              /* J' fm */
    for (int r = 0; r < nr; ++r) w->f[r] = w->fr[r] + w->tmpa[r] - w->f[r];
    if (!deriv) return;
    const size_t sl = (size_t)nm * nr;
    for (int i = 0; i < nr; ++i) { /* dMdq(:,:,i) = tmp' + tmp, tmp = J' Mm dJdq(:,:,i) */
        double* dM = w->dMdq + (size_t)i * nr * nr;
        mm_dense(nr, nm, nr, w->JtMm, w->dJdq + sl * i, w->tmpa);
        for (int r = 0; r < nr; ++r)
            for (int c = 0; c < nr; ++c) dM[(size_t)r * nr + c] = w->tmpa[(size_t)c * nr + r] + w->tmpa[(size_t)r * nr + c];
    }
    /* Dqvv = -J' Mm Jdot ; MmJdotqdot = Mm Jdot qdot */
    mm_dense(nr, nm, nr, w->JtMm, Jd, w->D);
    for (int i = 0; i < nr * nr; ++i) w->D[i] = -w->D[i];
    mv(nm, nm, w->Mm, w->v1, w->v2); /* MmJdotqdot */
    memset(w->K, 0, sizeof(double) * nr * nr);
    for (int i = 0; i < nr; ++i) {
        const double* dJ = w->dJdq + sl * i;
        const double* dJd = w->dJdotdq + sl * i;
        mtv(nm, nr, dJ, w->v2, w->tmpa);           /* dJdqi' MmJdotqdot */
        mv(nm, nr, dJd, w->qdot, w->v3);           /* dJdotdqi qdot */
        mv(nr, nm, w->JtMm, w->v3, w->tmpb);       /* JtMm * that */
        for (int r = 0; r < nr; ++r) w->K[(size_t)r * nr + i] = -w->tmpa[r] - w->tmpb[r]; /* Kqvv(:,i) */
        mv(nm, nr, dJ, w->qdot, w->v3);            /* dJdqi qdot */
        mv(nr, nm, w->JtMm, w->v3, w->tmpb);
        for (int r = 0; r < nr; ++r) w->D[(size_t)r * nr + i] -= w->tmpb[r]; /* Dqvv(:,i) */
    }
    /* K = Kr + J' Km J + Kqvv ; D = Dr + J' Dm J + Dqvv */
    mtm_dense(nm, nr, nm, J, w->Km, w->JtX);
    mm_dense(nr, nm, nr, w->JtX, J, w->tmpa);
    for (int i = 0; i < nr * nr; ++i) w->K[i] += w->Kr[i] + w->tmpa[i];
    mtm_dense(nm, nr, nm, J, w->Dm, w->JtX); /* JtDm */
    mm_dense(nr, nm, nr, w->JtX, J, w->tmpa);
    for (int i = 0; i < nr * nr; ++i) w->D[i] += w->Dr[i] + w->tmpa[i];
    for (int i = 0; i < nr; ++i) { /* K(:,i) += dJdqi' fm + JtDm dJdqi qdot */
        const double* dJ = w->dJdq + sl * i;
        mtv(nm, nr, dJ, w->fm, w->tmpa);
        mv(nm, nr, dJ, w->qdot, w->v3);
        mv(nr, nm, w->JtX, w->v3, w->tmpb);
        for (int r = 0; r < nr; ++r) w->K[(size_t)r * nr + i] += w->tmpa[r] + w->tmpb[r];
    }
}

/* ------------------------------------------------------------------ evalBDF1 / evalSDIRK2a / evalSDIRK2b / evalBDF2 */
enum { ST_BDF1 = 0, ST_SDIRK_A = 1, ST_SDIRK_B = 2, ST_BDF2 = 3 };

static void eval_stage(const oc_desc* d, oc_work* w, int stage, double h, const double* x, int deriv) {
    const int nr = w->nr;
    const double a = SDIRK_A;
    double cD, cK;
    double* dq = w->dx; /* reuse as dqtmp holder? no: keep separate */
    (void)dq;
    double* dqtmp = w->g0 + 0; /* placeholder, overwritten below */
    (void)dqtmp;
    double* dqt = (double*)alloca(sizeof(double) * nr);
    if (stage == ST_BDF1) {
        for (int i = 0; i < nr; ++i) {
            dqt[i] = x[i] - w->q0[i] - h * w->qdot0[i];
            w->qdot[i] = (x[i] - w->q0[i]) / h;
        }
        cD = h;
        cK = h * h;
    } else if (stage == ST_SDIRK_A) {
        const double ah = a * h;
        for (int i = 0; i < nr; ++i) {
            dqt[i] = x[i] - w->q0[i] - ah * w->qdot0[i];
            w->qdot[i] = (x[i] - w->q0[i]) / ah;
        }
        cD = ah;
        cK = ah * ah;
    } else if (stage == ST_SDIRK_B) {
        const double ah = a * h;
        for (int i = 0; i < nr; ++i) { /* q1/qdot1 slots hold (qa, qdota) */
            dqt[i] = x[i] - w->q0[i] - (2 * a - 1) * h * w->qdot0[i] - 2 * (1 - a) * h * w->qdot1[i];
            w->qdot[i] = (x[i] - w->q0[i] - (1 - a) * h * w->qdot1[i]) / ah;
        }
        cD = ah;
        cK = ah * ah;
    } else {
        for (int i = 0; i < nr; ++i) {
            const double e = x[i] - (4.0 / 3.0) * w->q1[i] + (1.0 / 3.0) * w->q0[i];
            dqt[i] = e - (8.0 / 9.0) * h * w->qdot1[i] + (2.0 / 9.0) * h * w->qdot0[i];
            w->qdot[i] = (3.0 / (2.0 * h)) * e;
        }
        cD = (2.0 / 3.0) * h;
        cK = (4.0 / 9.0) * (h * h);
    }
    for (int i = 0; i < nr; ++i) w->q[i] = x[i];
    scene_update(d, w);
    compute_values(d, w, deriv);
    mv(nr, nr, w->M, dqt, w->g);
    for (int i = 0; i < nr; ++i) w->g[i] -= cK * w->f[i];
    if (!deriv) return;
    for (int i = 0; i < nr * nr; ++i) w->H[i] = w->M[i] - cD * w->D[i] - cK * w->K[i];
    for (int i = 0; i < nr; ++i) { /* H(:,i) += dMdq(:,:,i) dqtmp */
        mv(nr, nr, w->dMdq + (size_t)i * nr * nr, dqt, w->tmpa);
        for (int r = 0; r < nr; ++r) w->H[(size_t)r * nr + i] += w->tmpa[r];
    }
}

/* dx = -H\g by partial-pivot elimination (dgesv) */
static void solve_neg(int n, const double* H, const double* g, double* LU, double* dx) {
    memcpy(LU, H, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) dx[i] = -g[i];
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(LU[(size_t)k * n + k]);
        for (int r = k + 1; r < n; ++r)
            if (fabs(LU[(size_t)r * n + k]) > best) {
                best = fabs(LU[(size_t)r * n + k]);
                p = r;
            }
        if (p != k) {
            for (int c = 0; c < n; ++c) {
                const double t = LU[(size_t)k * n + c];
                LU[(size_t)k * n + c] = LU[(size_t)p * n + c];
                LU[(size_t)p * n + c] = t;
            }
            const double t = dx[k];
            dx[k] = dx[p];
            dx[p] = t;
        }
        const double piv = LU[(size_t)k * n + k];
        for (int r = k + 1; r < n; ++r) {
            const double l = LU[(size_t)r * n + k] / piv;
            LU[(size_t)r * n + k] = l;
            for (int c = k + 1; c < n; ++c) LU[(size_t)r * n + c] -= l * LU[(size_t)k * n + c];
            dx[r] -= l * dx[k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = dx[k];
        for (int c = k + 1; c < n; ++c) s -= LU[(size_t)k * n + c] * dx[c];
        dx[k] = s / LU[(size_t)k * n + k];
    }
}

/* newton (driverRedMaxBDF1.m:94-157); x in/out; returns status bits; counts in it[0] (iterations), it[1] (ls evals) */
static int newton(const oc_desc* d, oc_work* w, int stage, double h, double* x, int* it) {
    const int nr = w->nr;
    const double tol = 1e-9, dxMax = 1e3;
    const int iterMax = 10 * nr, iterLsMax = 20;
    int iter = 1, status = 0;
    while (1) {
        eval_stage(d, w, stage, h, x, 1);
        solve_neg(nr, w->H, w->g, w->LU, w->dx);
        it[0]++;
        double dn = 0.0, f0 = 0.0;
        for (int i = 0; i < nr; ++i) dn += w->dx[i] * w->dx[i];
        if (sqrt(dn) > dxMax) {
            status |= 1;
            break;
        }
        for (int i = 0; i < nr; ++i) {
            f0 += w->g[i] * w->g[i];
            w->x0[i] = x[i];
        }
        f0 *= 0.5;
        double alpha = 1.0, gn = 0.0;
        int iterLs = 1;
        while (1) {
            for (int i = 0; i < nr; ++i) x[i] = w->x0[i] + alpha * w->dx[i];
            eval_stage(d, w, stage, h, x, 0);
            it[1]++;
            gn = 0.0;
            for (int i = 0; i < nr; ++i) gn += w->g[i] * w->g[i];
            if (0.5 * gn < f0) break;
            if (iterLs >= iterLsMax) {
                status |= 4;
                break;
            }
            alpha = 0.5 * alpha;
            iterLs++;
        }
        if (sqrt(gn) < tol) break;
        if (iter >= iterMax) {
            status |= 2;
            break;
        }
        iter++;
    }
    return status;
}

/* simLoop of driverRedMaxBDF1.m:57-91 / driverRedMaxBDF2.m:57-125 for one rollout */
static int rollout_one(const oc_desc* d, oc_work* w, int scheme, double h, int nsteps, const double* q0, const double* qd0,
                       const double* tau, double* q_out, double* qd_out, int* it) {
    const int nr = w->nr;
    const double a = SDIRK_A;
    int status = 0;
    double* qc = (double*)alloca(sizeof(double) * nr);
    double* qdc = (double*)alloca(sizeof(double) * nr);
    double* x = w->x;
    for (int i = 0; i < nr; ++i) {
        qc[i] = q0[i];
        qdc[i] = qd0[i];
        w->tau[i] = tau ? tau[i] : 0.0;
        w->q1[i] = qc[i];
        w->qdot1[i] = qdc[i];
    }
    for (int k = 0; k < nsteps; ++k) {
        if (scheme == 1) {
            for (int i = 0; i < nr; ++i) {
                w->q0[i] = qc[i];
                w->qdot0[i] = qdc[i];
                x[i] = qc[i] + h * qdc[i];
            }
            status |= newton(d, w, ST_BDF1, h, x, it);
            for (int i = 0; i < nr; ++i) {
                qdc[i] = (x[i] - w->q0[i]) / h;
                qc[i] = x[i];
            }
        } else if (k == 0) {
            double* qa = (double*)alloca(sizeof(double) * nr);
            double* qda = (double*)alloca(sizeof(double) * nr);
            for (int i = 0; i < nr; ++i) {
                w->q0[i] = qc[i];
                w->qdot0[i] = qdc[i];
                x[i] = qc[i] + a * h * qdc[i];
            }
            status |= newton(d, w, ST_SDIRK_A, h, x, it);
            for (int i = 0; i < nr; ++i) {
                qa[i] = x[i];
                qda[i] = (x[i] - w->q0[i]) / (a * h);
                w->q1[i] = qa[i];
                w->qdot1[i] = qda[i];
                x[i] = qa[i] + (1 - a) * h * qda[i];
            }
            status |= newton(d, w, ST_SDIRK_B, h, x, it);
            for (int i = 0; i < nr; ++i) {
                qdc[i] = (x[i] - w->q0[i] - (1 - a) * h * qda[i]) / (a * h);
                qc[i] = x[i];
                w->q1[i] = w->q0[i];
                w->qdot1[i] = w->qdot0[i];
            }
        } else {
            for (int i = 0; i < nr; ++i) {
                w->q0[i] = w->q1[i];
                w->qdot0[i] = w->qdot1[i];
                w->q1[i] = qc[i];
                w->qdot1[i] = qdc[i];
                x[i] = qc[i] + h * qdc[i];
            }
            status |= newton(d, w, ST_BDF2, h, x, it);
            for (int i = 0; i < nr; ++i) {
                qdc[i] = (3.0 / (2.0 * h)) * (x[i] - (4.0 / 3.0) * w->q1[i] + (1.0 / 3.0) * w->q0[i]);
                qc[i] = x[i];
            }
        }
        for (int i = 0; i < nr; ++i) {
            q_out[(size_t)k * nr + i] = qc[i];
            if (qd_out) qd_out[(size_t)k * nr + i] = qdc[i];
        }
    }
    return status;
}

/* ------------------------------------------------------------------ exported */
int oc_nr(const oc_desc* d) {
    int nr = 0;
    for (int j = 0; j < d->n; ++j) nr += d->jtype[j] == 1;
    return nr;
}

int oc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* B rollouts; q0, qd0, tau: [B][nr]; q_out, qd_out: [B][nsteps][nr]; stats: [B][3] = iterations, ls evals, status */
int oc_rollout(const oc_desc* d, int scheme, double h, int nsteps, int B, const double* q0, const double* qd0,
               const double* tau, double* q_out, double* qd_out, int32_t* stats, int threads) {
    const int nr = oc_nr(d);
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        oc_work* w = work_create(d);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            int it[2] = {0, 0};
            const int st = rollout_one(d, w, scheme, h, nsteps, q0 + (size_t)b * nr, qd0 + (size_t)b * nr,
                                       tau ? tau + (size_t)b * nr : NULL, q_out + (size_t)b * nsteps * nr,
                                       qd_out ? qd_out + (size_t)b * nsteps * nr : NULL, it);
            if (stats) {
                stats[3 * b] = it[0];
                stats[3 * b + 1] = it[1];
                stats[3 * b + 2] = st;
            }
        }
        work_destroy(w);
    }
    return 0;
}

/* one BDF1-style evaluation for cross-checks: state (q, qdot) given directly, dqtmp given; outputs g, H, M, D, K, f */
int oc_eval(const oc_desc* d, const double* q, const double* qdot, const double* dqtmp, const double* tau, double cD,
            double cK, double* g, double* H, double* M, double* D, double* K, double* f) {
    oc_work* w = work_create(d);
    const int nr = w->nr;
    for (int i = 0; i < nr; ++i) {
        w->q[i] = q[i];
        w->qdot[i] = qdot[i];
        w->tau[i] = tau ? tau[i] : 0.0;
    }
    scene_update(d, w);
    compute_values(d, w, 1);
    mv(nr, nr, w->M, dqtmp, w->g);
    for (int i = 0; i < nr; ++i) w->g[i] -= cK * w->f[i];
    for (int i = 0; i < nr * nr; ++i) w->H[i] = w->M[i] - cD * w->D[i] - cK * w->K[i];
    for (int i = 0; i < nr; ++i) {
        mv(nr, nr, w->dMdq + (size_t)i * nr * nr, dqtmp, w->tmpa);
        for (int r = 0; r < nr; ++r) w->H[(size_t)r * nr + i] += w->tmpa[r];
    }
    if (g) memcpy(g, w->g, sizeof(double) * nr);
    if (f) memcpy(f, w->f, sizeof(double) * nr);
    if (H) memcpy(H, w->H, sizeof(double) * nr * nr);
    if (M) memcpy(M, w->M, sizeof(double) * nr * nr);
    if (D) memcpy(D, w->D, sizeof(double) * nr * nr);
    if (K) memcpy(K, w->K, sizeof(double) * nr * nr);
    work_destroy(w);
    return 0;
}
