// rmx_api.cu -- host side of the C ABI declared in include/redmax_b200.h.
//
// Scene flattening follows Scene.init (matlab-diff/+redmax/Scene.m:59-119): joints listed parents-first,
// reduced indices assigned leaf-to-root (Scene.m:69-71 + Joint.countDofs, Joint.m:149).  Internally joints are
// renumbered in DFS preorder so that every subtree is a contiguous index range.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "rmx_host.h"

using namespace rmx;

static thread_local std::string g_err;
int rmx_fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
static int fail(int code, const std::string& msg) { return rmx_fail(code, msg); }

struct rmx_scene {
    int n = 0, nr = 0, nm = 0;
    int is_chain = 0, has_ground = 0, has_chart = 0;
    double grav[3] = {0, 0, 0};
    std::vector<JointConst> jc;   // internal (preorder) order
    std::vector<int> ends_list;
    std::vector<int> user2int;    // expanded (virtual) joint index -> internal (preorder) index
    std::vector<int> body2int;    // user joint/body index -> internal index of the virtual joint that carries the body
    std::vector<PointForce> pf;   // ForcePointPoint forces (internal body indices)
    std::vector<int> pf_ep;       // per-body attachment lists (JointConst::pf_ptr / pf_cnt), entry = PF_MAXPTS*force + attachment
    int pf_doubles = 0;           // shared-memory scratch of the forces (attachment records + cross blocks)
    std::vector<int> anc;         // [nrounds][n] 2^r-th ancestors (internal indices) for the pointer-jumping scans
    int nrounds = 0;
    int impl = 2;                 // 1 = sweep kernels (rmx_device.cuh), 2 = composite kernels (rmx_fast.cuh)
    std::map<int, DevCopy> dev;   // per CUDA device
    std::vector<int> kry_devs;    // devices that ran the last Krylov-solve rollout
};

extern "C" int rmx_version(void) { return RMX_VERSION; }
extern "C" const char* rmx_last_error(void) { return g_err.c_str(); }
extern "C" int rmx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" void rmx_opts_default(rmx_opts* o, int32_t scheme, int32_t adjoint) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->scheme = scheme;
    o->nsteps = 100;          // ceil(tEnd/h) with Scene.m:38,41 defaults
    o->h = 1e-2;              // Scene.m:41
    o->tol = 1e-9;            // driverRedMaxBDF1.m:95
    o->dxMax = 1e3;           // :96
    o->iterMaxFactor = adjoint ? 5 : 10;  // :97 ; driverRedMaxAdjointBDF1.m:108
    o->iterLsMax = 20;        // :98
    o->linsolve = RMX_LINSOLVE_LU;
    o->ngpus = 1;
    o->tau_mode = RMX_TAU_NONE;
    o->pcg_maxit = 0;         // 0 -> 4*nr
    o->pcg_tol = 1e-6;        // c++/PCG Solver.h:43
}

static void colmajor4_to_Rp(const double* E, double* R, double* p) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = E[4 * c + r];
        p[r] = E[12 + r];
    }
}

// A multi-DOF joint whose Q(q) is a product of one-parameter motions is integrated as a chain of one-DOF "virtual"
// joints; only the last one carries the body (the others have zero inertia):
//   JointPlanar        Q = trans(B q)                 -> prismatic b1, prismatic b2          (JointPlanar.m:24-28)
//   JointTranslational Q = trans(q)                   -> prismatic x, y, z                   (JointTranslational.m:20-23)
//   JointFree2D        Q = [Rz(q3) [q1 q2 0]'; 0 1]   -> prismatic x, prismatic y, revolute z (JointFree2D.m:20-31)
//   JointUniversal     R = X(q1) Y(q2)                -> revolute x, revolute y              (JointUniversal.m:73-75)
// g(q) and H = dg/dq are functions of the motion only, so the reference's S(q), Sdot, dSdq ... of these joints are
// reproduced by the world-frame screws of the virtual joints (parity: tests/test_gpu_joints.py vs the oracle's restatement
// of the reference classes).
struct ExpandedScene {
    int n = 0;
    std::vector<int> parent, type, idx, user;  // type: RMX_JOINT_FIXED / REVOLUTE / PRISMATIC; idx: reduced index or -1
    std::vector<double> E0_pj, E0_ji, axis, I_i, sides, stiffness, damping, qRest, qLimL, qLimU, qLimK, qLimD;
    std::vector<int> body_of_user;             // user joint/body -> index (in this list) of the virtual joint carrying the body
    std::vector<int> chart_mid;                // middle Euler angle of a spherical / Free3D joint
};

static int joint_ndof(int jtype) {
    switch (jtype) {
        case RMX_JOINT_FIXED: return 0;
        case RMX_JOINT_REVOLUTE: case RMX_JOINT_PRISMATIC: return 1;
        case RMX_JOINT_PLANAR: case RMX_JOINT_UNIVERSAL: return 2;
        case RMX_JOINT_TRANSLATIONAL: case RMX_JOINT_FREE2D: case RMX_JOINT_SPHERICAL: return 3;
        case RMX_JOINT_FREE3D: return 6;
        default: return -1;
    }
}

static void expand_scene(const rmx_scene_desc* d, const std::vector<int>& base, ExpandedScene& x) {
    static const double ID4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    static const double EX[3] = {1, 0, 0}, EY[3] = {0, 1, 0}, EZ[3] = {0, 0, 1}, Z3[3] = {0, 0, 0};
    const int n = d->n;
    x.body_of_user.assign(n, -1);
    for (int j = 0; j < n; ++j) {
        const int jt = d->jtype[j];
        const double* ax1 = d->axis + 3 * j;
        const double* ax2 = d->axis2 ? d->axis2 + 3 * j : EY;
        int vt[6] = {RMX_JOINT_FIXED, 0, 0, 0, 0, 0};
        const double* va[6] = {Z3, Z3, Z3, Z3, Z3, Z3};
        int nv = 1, mid = -1;
        // Euler charts in the reference's numbering (JointSpherical.m:5-16): axes of R = R_a(q1) R_b(q2) R_c(q3); det T is
        // sin q2 for the proper Euler charts 1..6 and cos q2 for the Tait-Bryan charts 7..12
        static const int CH[13][3] = {{0, 1, 2}, {0, 1, 0}, {0, 2, 0}, {1, 2, 1}, {1, 0, 1}, {2, 0, 2}, {2, 1, 2},
                                      {0, 1, 2}, {0, 2, 1}, {1, 2, 0}, {1, 0, 2}, {2, 0, 1}, {2, 1, 0}};
        const double* EA[3] = {EX, EY, EZ};
        int ch = d->chart ? d->chart[j] : 0;
        if (ch < 1 || ch > 12) ch = 7;
        const int midkind = ch <= 6 ? 2 : 1;
        switch (jt) {
            case RMX_JOINT_REVOLUTE: vt[0] = RMX_JOINT_REVOLUTE; va[0] = ax1; break;
            case RMX_JOINT_PRISMATIC: vt[0] = RMX_JOINT_PRISMATIC; va[0] = ax1; break;
            case RMX_JOINT_PLANAR: nv = 2; vt[0] = vt[1] = RMX_JOINT_PRISMATIC; va[0] = ax1; va[1] = ax2; break;
            case RMX_JOINT_TRANSLATIONAL: nv = 3; vt[0] = vt[1] = vt[2] = RMX_JOINT_PRISMATIC; va[0] = EX; va[1] = EY; va[2] = EZ; break;
            case RMX_JOINT_FREE2D: nv = 3; vt[0] = vt[1] = RMX_JOINT_PRISMATIC; vt[2] = RMX_JOINT_REVOLUTE; va[0] = EX; va[1] = EY; va[2] = EZ; break;
            case RMX_JOINT_UNIVERSAL: nv = 2; vt[0] = vt[1] = RMX_JOINT_REVOLUTE; va[0] = EX; va[1] = EY; break;
            // Euler chart XYZ, R = X(q1) Y(q2) Z(q3) (JointSpherical.m:33,1086): g and dg/dq depend on the motion only, so the
            // reference's T(q), Tdot and their derivatives are reproduced by three revolute virtual joints
            case RMX_JOINT_SPHERICAL:
                nv = 3;
                vt[0] = vt[1] = vt[2] = RMX_JOINT_REVOLUTE;
                for (int v = 0; v < 3; ++v) va[v] = EA[CH[ch][v]];
                mid = 1;
                break;
            // Q = [R p; 0 1] (JointFree3D.m:56-58): translate in the parent frame, then rotate
            case RMX_JOINT_FREE3D:
                nv = 6;
                vt[0] = vt[1] = vt[2] = RMX_JOINT_PRISMATIC;
                vt[3] = vt[4] = vt[5] = RMX_JOINT_REVOLUTE;
                va[0] = EX; va[1] = EY; va[2] = EZ;
                for (int v = 0; v < 3; ++v) va[3 + v] = EA[CH[ch][v]];
                mid = 4;
                break;
            default: break;
        }
        for (int v = 0; v < nv; ++v) {
            const bool first = v == 0, last = v == nv - 1;
            x.parent.push_back(first ? (d->parent[j] < 0 ? -1 : x.body_of_user[d->parent[j]]) : x.n - 1);
            x.type.push_back(vt[v]);
            x.user.push_back(j);
            x.chart_mid.push_back(v == mid ? midkind : 0);
            x.idx.push_back(vt[v] == RMX_JOINT_FIXED ? -1 : base[j] + v);
            const double* Epj = first ? d->E0_pj + 16 * j : ID4;
            const double* Eji = last ? d->E0_ji + 16 * j : ID4;
            x.E0_pj.insert(x.E0_pj.end(), Epj, Epj + 16);
            x.E0_ji.insert(x.E0_ji.end(), Eji, Eji + 16);
            x.axis.insert(x.axis.end(), va[v], va[v] + 3);
            for (int i = 0; i < 6; ++i) x.I_i.push_back(last ? d->I_i[6 * j + i] : 0.0);
            for (int i = 0; i < 3; ++i) x.sides.push_back(last ? d->sides[3 * j + i] : 0.0);
            x.stiffness.push_back(d->stiffness ? d->stiffness[j] : 0.0);
            x.damping.push_back(d->damping ? d->damping[j] : 0.0);
            x.qRest.push_back(d->qRest ? d->qRest[RMX_MAX_JOINT_DOF * j + v] : 0.0);
            x.qLimL.push_back(d->qLimL ? d->qLimL[j] : -1e8);  // Joint.m:77-80
            x.qLimU.push_back(d->qLimU ? d->qLimU[j] : 1e8);
            x.qLimK.push_back(d->qLimK ? d->qLimK[j] : 1e8);
            x.qLimD.push_back(d->qLimD ? d->qLimD[j] : 0.0);
            if (last) x.body_of_user[j] = x.n;
            x.n++;
        }
    }
}

// Host forward kinematics at the initial configuration (what Scene.init's update() leaves in body.E_wi, Scene.m:93): only
// needed for quantities the reference derives from it at init time -- the rest length of ForceSpringDamper
// (ForceSpringDamper.m:38-62).  jc: internal (preorder) joints; out: world frame of every body, R row-major and p.
static void host_aa_to_mat(const JointConst& J, double angle, double* R) {
    if (J.axtype != 0 && J.axsign < 0) angle = -angle;
    const double sn = std::sin(angle), cs = std::cos(angle);
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(R, I, sizeof(I));
    if (J.axtype == 3) {
        R[0] = cs; R[1] = -sn; R[3] = sn; R[4] = cs;
    } else if (J.axtype == 1) {
        R[4] = cs; R[5] = -sn; R[7] = sn; R[8] = cs;
    } else if (J.axtype == 2) {
        R[0] = cs; R[2] = sn; R[6] = -sn; R[8] = cs;
    } else {
        const double ax = J.axn[0], ay = J.axn[1], az = J.axn[2], t = 1.0 - cs;
        const double xz = ax * az, xy = ax * ay, yz = ay * az;
        R[0] = t * ax * ax + cs; R[1] = t * xy - sn * az; R[2] = t * xz + sn * ay;
        R[3] = t * xy + sn * az; R[4] = t * ay * ay + cs; R[5] = t * yz - sn * ax;
        R[6] = t * xz - sn * ay; R[7] = t * yz + sn * ax; R[8] = t * az * az + cs;
    }
}
static void h_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void h_mv(const double* A, const double* x, double* y) {
    for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
static void host_body_frames(const std::vector<JointConst>& jc, std::vector<double>& Rb, std::vector<double>& pb) {
    const int n = (int)jc.size();
    std::vector<double> Rw(9 * n), pw(3 * n);
    Rb.assign(9 * n, 0.0);
    pb.assign(3 * n, 0.0);
    for (int k = 0; k < n; ++k) {
        const JointConst& J = jc[k];
        double Rl[9], pl[3] = {J.p0[0], J.p0[1], J.p0[2]};
        if (J.idx >= 0 && !J.prismatic) {
            double Rq[9];
            host_aa_to_mat(J, J.qRest, Rq);
            h_mul(J.R0, Rq, Rl);
        } else {
            std::memcpy(Rl, J.R0, sizeof(Rl));
            if (J.idx >= 0) {
                const double aq[3] = {J.axis[0] * J.qRest, J.axis[1] * J.qRest, J.axis[2] * J.qRest};
                double t3[3];
                h_mv(J.R0, aq, t3);
                for (int i = 0; i < 3; ++i) pl[i] += t3[i];
            }
        }
        if (J.parent < 0) {
            std::memcpy(&Rw[9 * k], Rl, sizeof(Rl));
            std::memcpy(&pw[3 * k], pl, sizeof(pl));
        } else {
            double t3[3];
            h_mul(&Rw[9 * J.parent], Rl, &Rw[9 * k]);
            h_mv(&Rw[9 * J.parent], pl, t3);
            for (int i = 0; i < 3; ++i) pw[3 * k + i] = pw[3 * J.parent + i] + t3[i];
        }
        double t3[3];
        h_mul(&Rw[9 * k], J.Rji, &Rb[9 * k]);
        h_mv(&Rw[9 * k], J.pji, t3);
        for (int i = 0; i < 3; ++i) pb[3 * k + i] = pw[3 * k + i] + t3[i];
    }
}

extern "C" int rmx_scene_create(const rmx_scene_desc* d, rmx_scene** out) {
    if (!d || !out) return fail(RMX_EINVAL, "rmx_scene_create: null argument");
    *out = nullptr;
    const int n_user = d->n;
    if (n_user < 1) return fail(RMX_EINVAL, "rmx_scene_create: n < 1");
    if (!d->parent || !d->jtype || !d->E0_pj || !d->E0_ji || !d->axis || !d->I_i || !d->sides)
        return fail(RMX_EINVAL, "rmx_scene_create: missing required array");
    for (int j = 0; j < n_user; ++j) {
        if (d->parent[j] >= j || d->parent[j] < -1)
            return fail(RMX_EINVAL, "rmx_scene_create: joints must be listed parents-before-children (Joint.m:134)");
        if (d->chart && (d->chart[j] < 0 || d->chart[j] > 12)) return fail(RMX_EINVAL, "rmx_scene_create: chart must be 0 (default) or 1..12");
        if (joint_ndof(d->jtype[j]) < 0)
            return fail(RMX_EINVAL, "rmx_scene_create: unknown joint type (fixed, revolute, prismatic, planar, translational, "
                                    "free2d, universal, spherical and free3d are on the hot path)");
    }
    if (d->nground < 0) return fail(RMX_EINVAL, "rmx_scene_create: nground < 0");
    if (d->nground > 0 && (!d->ground_body || !d->ground_E || !d->ground_kn || !d->ground_kt || !d->ground_kd || !d->ground_mu))
        return fail(RMX_EINVAL, "rmx_scene_create: missing ground_* array");
    // reference numbering: countDofs is called for i = n..1 (Scene.m:69-71); a joint's DOFs are consecutive (Joint.m:152)
    std::vector<int> base(n_user, 0);
    int nr = 0;
    for (int j = n_user - 1; j >= 0; --j) {
        base[j] = nr;
        nr += joint_ndof(d->jtype[j]);
    }
    ExpandedScene x;
    expand_scene(d, base, x);
    const int n = x.n;
    if (n > 128) return fail(RMX_ELIMIT, "rmx_scene_create: more than 128 (virtual) joints not supported by the in-block solver");
    rmx_scene* s = new rmx_scene();
    s->n = n;
    const std::vector<int>& idxR = x.idx;
    s->nr = nr;
    s->nm = 6 * n_user;
    if (nr < 1) {
        delete s;
        return fail(RMX_EINVAL, "rmx_scene_create: scene has no degrees of freedom");
    }
    // DFS preorder
    std::vector<std::vector<int>> children(n);
    std::vector<int> roots;
    for (int j = 0; j < n; ++j) {
        if (x.parent[j] < 0)
            roots.push_back(j);
        else
            children[x.parent[j]].push_back(j);
    }
    std::vector<int> order;  // internal -> user
    order.reserve(n);
    {
        std::vector<int> stack;
        for (int ri = (int)roots.size() - 1; ri >= 0; --ri) stack.push_back(roots[ri]);
        while (!stack.empty()) {
            int j = stack.back();
            stack.pop_back();
            order.push_back(j);
            for (int ci = (int)children[j].size() - 1; ci >= 0; --ci) stack.push_back(children[j][ci]);
        }
    }
    s->user2int.assign(n, -1);
    for (int k = 0; k < n; ++k) s->user2int[order[k]] = k;
    s->jc.assign(n, JointConst());
    std::vector<int> size(n, 1);
    s->is_chain = 1;
    for (int k = 0; k < n; ++k) {
        const int j = order[k];
        JointConst& J = s->jc[k];
        std::memset(&J, 0, sizeof(J));
        colmajor4_to_Rp(x.E0_pj.data() + 16 * j, J.R0, J.p0);
        colmajor4_to_Rp(x.E0_ji.data() + 16 * j, J.Rji, J.pji);
        for (int i = 0; i < 3; ++i) {
            J.axis[i] = x.axis[3 * j + i];
            J.hs[i] = 0.5 * x.sides[3 * j + i];  // ForceGroundCuboid.m:71
        }
        for (int i = 0; i < 6; ++i) J.I[i] = x.I_i[6 * j + i];
        J.stiff = x.stiffness[j];
        J.damp = x.damping[j];
        J.qRest = x.qRest[j];
        J.qLimL = x.qLimL[j];
        J.qLimU = x.qLimU[j];
        J.qLimK = x.qLimK[j];
        J.qLimD = x.qLimD[j];
        J.parent = x.parent[j] < 0 ? -1 : s->user2int[x.parent[j]];
        if (J.parent != k - 1) s->is_chain = 0;
        J.prismatic = x.type[j] == RMX_JOINT_PRISMATIC;
        J.chart_mid = x.chart_mid[j];
        if (J.chart_mid) s->has_chart = 1;
        J.idx = idxR[j];
        // se3.aaToMat classification (se3.m:118-176)
        J.axtype = 0;
        J.axsign = 1;
        if (x.type[j] == RMX_JOINT_PRISMATIC) {
            const double mag = std::sqrt(J.axis[0] * J.axis[0] + J.axis[1] * J.axis[1] + J.axis[2] * J.axis[2]);
            if (!(mag > 1e-9)) {
                delete s;
                return fail(RMX_EINVAL, "rmx_scene_create: zero prismatic axis");
            }
        }
        if (x.type[j] == RMX_JOINT_REVOLUTE) {
            double ax = J.axis[0], ay = J.axis[1], az = J.axis[2];
            double mag = std::sqrt(ax * ax + ay * ay + az * az);
            if (!(mag > 1e-9)) {
                delete s;
                return fail(RMX_EINVAL, "rmx_scene_create: zero revolute axis");
            }
            mag = 1.0 / mag;
            ax = ax * mag;
            ay = ay * mag;
            az = az * mag;
            J.axn[0] = ax;
            J.axn[1] = ay;
            J.axn[2] = az;
            const double TH = 1e-9;
            if (std::fabs(ax) < TH && std::fabs(ay) < TH) {
                J.axtype = 3;
                J.axsign = az < 0 ? -1 : 1;
            } else if (std::fabs(ay) < TH && std::fabs(az) < TH) {
                J.axtype = 1;
                J.axsign = ax < 0 ? -1 : 1;
            } else if (std::fabs(az) < TH && std::fabs(ax) < TH) {
                J.axtype = 2;
                J.axsign = ay < 0 ? -1 : 1;
            }
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        s->jc[k].end = k + size[k];
        if (s->jc[k].parent >= 0) size[s->jc[k].parent] += size[k];
    }
    // ends lists: for each index m, joints k (k < m) with end == m
    {
        std::vector<std::vector<int>> ends(n + 1);
        for (int k = 0; k < n; ++k) ends[s->jc[k].end].push_back(k);
        for (int m = 0; m < n; ++m) {
            s->jc[m].ends_ptr = (int)s->ends_list.size();
            s->jc[m].ends_cnt = (int)ends[m].size();
            for (int k : ends[m]) s->ends_list.push_back(k);
        }
        if (s->ends_list.empty()) s->ends_list.push_back(0);
    }
    // ancestor tables: anc[0] = parent, anc[r+1][j] = anc[r][anc[r][j]]
    {
        std::vector<int> cur(n);
        for (int k = 0; k < n; ++k) cur[k] = s->jc[k].parent;
        while (true) {
            bool any = false;
            for (int k = 0; k < n; ++k) any = any || cur[k] >= 0;
            if (!any) break;
            s->anc.insert(s->anc.end(), cur.begin(), cur.end());
            s->nrounds++;
            std::vector<int> nxt(n);
            for (int k = 0; k < n; ++k) nxt[k] = cur[k] >= 0 ? cur[cur[k]] : -1;
            cur.swap(nxt);
        }
        if (s->anc.empty()) s->anc.push_back(-1);
    }
    {
        const char* e = std::getenv("RMX_IMPL");  // developer switch: 1 forces the sweep kernels
        s->impl = (e && e[0] == '1') ? 1 : 2;
        if (n > 64) s->impl = 1;  // the composite path keeps ~90 doubles per joint in shared memory
        bool any_prismatic = false;
        for (int k = 0; k < n; ++k) any_prismatic = any_prismatic || s->jc[k].prismatic;
        if (s->impl == 1 && any_prismatic) {
            delete s;
            return fail(RMX_ELIMIT, "rmx_scene_create: translational joint DOFs need the composite kernels (at most 64 virtual joints)");
        }
    }
    for (int i = 0; i < 3; ++i) s->grav[i] = d->grav[i];
    for (int f = 0; f < d->nground; ++f) {
        const int b = d->ground_body[f];
        if (b < 0 || b >= n_user) {
            delete s;
            return fail(RMX_EINVAL, "rmx_scene_create: ground_body out of range");
        }
        JointConst& J = s->jc[s->user2int[x.body_of_user[b]]];
        if (J.has_ground) {
            delete s;
            return fail(RMX_EINVAL, "rmx_scene_create: at most one ForceGroundCuboid per body");
        }
        J.has_ground = 1;
        const double* E = d->ground_E + 16 * f;
        for (int i = 0; i < 3; ++i) {
            J.gxg[i] = E[12 + i];  // E(1:3,4)  ForceGroundCuboid.m:56
            J.gng[i] = E[8 + i];   // E(1:3,3)  ForceGroundCuboid.m:57
        }
        J.gkn = d->ground_kn[f];
        J.gkt = d->ground_kt[f];
        J.gkd = d->ground_kd[f];
        J.gmu = d->ground_mu[f];
        s->has_ground = 1;
    }
    s->body2int.assign(n_user, -1);
    for (int j = 0; j < n_user; ++j) s->body2int[j] = s->user2int[x.body_of_user[j]];
    // forces between body points: ForcePointPoint / ForceSpringDamper (pf_*) and ForceCable (cable_*)
    const int nforce = (d->npointforce > 0 ? d->npointforce : 0) + (d->ncable > 0 ? d->ncable : 0);
    if (d->npointforce < 0 || d->ncable < 0 || nforce > RMX_MAX_POINTFORCE) {
        delete s;
        return fail(RMX_ELIMIT, "rmx_scene_create: at most RMX_MAX_POINTFORCE point forces and cables");
    }
    if (nforce > 0) {
        if ((d->npointforce > 0 && (!d->pf_body1 || !d->pf_body2 || !d->pf_x1 || !d->pf_x2 || !d->pf_ks || !d->pf_kd)) ||
            (d->ncable > 0 && (!d->cable_npts || !d->cable_body || !d->cable_x || !d->cable_ks || !d->cable_kd))) {
            delete s;
            return fail(RMX_EINVAL, "rmx_scene_create: missing point-force / cable array");
        }
        if (s->impl != 2) {
            delete s;
            return fail(RMX_ELIMIT, "rmx_scene_create: point forces need the composite kernels (at most 64 virtual joints)");
        }
        static_assert(RMX_MAX_CABLE_POINTS == PF_MAXPTS, "cable capacity");
        for (int f = 0; f < nforce; ++f) {
            PointForce P;
            std::memset(&P, 0, sizeof(P));
            int ub[PF_MAXPTS] = {-1, -1, -1, -1};
            if (f < d->npointforce) {
                P.kind = d->pf_kind ? d->pf_kind[f] : RMX_FORCE_POINTPOINT;
                if (P.kind != RMX_FORCE_POINTPOINT && P.kind != RMX_FORCE_SPRINGDAMPER) {
                    delete s;
                    return fail(RMX_EINVAL, "rmx_scene_create: unknown point-force kind");
                }
                P.npts = 2;
                ub[0] = d->pf_body1[f];
                ub[1] = d->pf_body2[f];
                for (int i = 0; i < 3; ++i) {
                    P.x[0][i] = d->pf_x1[3 * f + i];
                    P.x[1][i] = d->pf_x2[3 * f + i];
                }
                P.ks = d->pf_ks[f];
                P.kd = d->pf_kd[f];
                P.L = (d->pf_L && P.kind == RMX_FORCE_SPRINGDAMPER) ? d->pf_L[f] : 0.0;
            } else {
                const int cI = f - d->npointforce;
                P.kind = RMX_FORCE_CABLE;
                P.npts = d->cable_npts[cI];
                if (P.npts < 2 || P.npts > PF_MAXPTS) {
                    delete s;
                    return fail(RMX_ELIMIT, "rmx_scene_create: a cable has 2 .. RMX_MAX_CABLE_POINTS points");
                }
                for (int k = 0; k < P.npts; ++k) {
                    ub[k] = d->cable_body[PF_MAXPTS * cI + k];
                    for (int i = 0; i < 3; ++i) P.x[k][i] = d->cable_x[3 * (PF_MAXPTS * cI + k) + i];
                }
                P.ks = d->cable_ks[cI];
                P.kd = d->cable_kd[cI];
                P.L = d->cable_L ? d->cable_L[cI] : 0.0;
            }
            for (int k = 0; k < P.npts; ++k) {
                if (ub[k] < -1 || ub[k] >= n_user) {
                    delete s;
                    return fail(RMX_EINVAL, "rmx_scene_create: point-force body out of range");
                }
                P.body[k] = ub[k] < 0 ? -1 : s->body2int[ub[k]];
                for (int k2 = 0; k2 < k; ++k2)
                    if (P.body[k] >= 0 && P.body[k] == P.body[k2]) {
                        delete s;
                        return fail(RMX_EINVAL, "rmx_scene_create: the points of a force must lie on different bodies (or the world)");
                    }
            }
            s->pf.push_back(P);
        }
        // shared-memory scratch offsets, attachment lists, rest lengths
        std::vector<std::vector<int>> ep(n);
        std::vector<double> Rb, pb;
        bool have = false;
        int off = 0;
        for (int f = 0; f < nforce; ++f) {
            PointForce& P = s->pf[f];
            P.rec_off = off;
            off += P.npts * PF_REC;
            P.blk_off = off;
            off += P.npts * P.npts * PF_BLK;
            for (int k = 0; k < P.npts; ++k)
                if (P.body[k] >= 0) ep[P.body[k]].push_back(PF_MAXPTS * f + k);
            if (P.kind == RMX_FORCE_POINTPOINT || P.L > 0) continue;
            // rest length not given: (routed) distance of the points in the initial configuration
            // (ForceSpringDamper.m:38-62, ForceCable.m:36-63)
            if (!have) {
                host_body_frames(s->jc, Rb, pb);
                have = true;
            }
            double prev[3] = {0, 0, 0};
            P.L = 0.0;
            for (int k = 0; k < P.npts; ++k) {
                double xw[3];
                if (P.body[k] >= 0) {
                    h_mv(&Rb[9 * P.body[k]], P.x[k], xw);
                    for (int i = 0; i < 3; ++i) xw[i] += pb[3 * P.body[k] + i];
                } else {
                    for (int i = 0; i < 3; ++i) xw[i] = P.x[k][i];
                }
                if (k > 0) {
                    const double d0 = xw[0] - prev[0], d1 = xw[1] - prev[1], d2 = xw[2] - prev[2];
                    P.L += std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                }
                for (int i = 0; i < 3; ++i) prev[i] = xw[i];
            }
            if (!(P.L > 0)) {
                delete s;
                return fail(RMX_EINVAL, "rmx_scene_create: spring / cable with zero rest length");
            }
        }
        s->pf_doubles = off;
        for (int k = 0; k < n; ++k) {
            s->jc[k].pf_ptr = (int)s->pf_ep.size();
            s->jc[k].pf_cnt = (int)ep[k].size();
            for (int e : ep[k]) s->pf_ep.push_back(e);
        }
        s->has_ground = 1;  // external-force fields (AEXT / CEXT) of the composite kernels
    }
    *out = s;
    return RMX_OK;
}

extern "C" void rmx_scene_destroy(rmx_scene* s) {
    if (!s) return;
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : s->dev) {
        cudaSetDevice(kv.first);
        cudaFree(kv.second.jc);
        cudaFree(kv.second.ends);
        cudaFree(kv.second.anc);
        cudaFree(kv.second.pf);
        cudaFree(kv.second.pf_ep);
        cudaFree(kv.second.kry);
        for (auto& b : kv.second.buf) cudaFree(b.p);
        if (kv.second.stream) cudaStreamDestroy(kv.second.stream);
        if (kv.second.copy_stream) cudaStreamDestroy(kv.second.copy_stream);
        for (cudaEvent_t e : kv.second.chunk_done)
            if (e) cudaEventDestroy(e);
        for (void* h : kv.second.stage)
            if (h) cudaFreeHost(h);
    }
    cudaSetDevice(cur);
    cudaGetLastError();
    delete s;
}
extern "C" int rmx_scene_nr(const rmx_scene* s) { return s ? s->nr : 0; }
extern "C" int rmx_scene_nm(const rmx_scene* s) { return s ? s->nm : 0; }

static int scene_on_device(rmx_scene* s, int dev, DevCopy** out) {
    auto it = s->dev.find(dev);
    if (it == s->dev.end()) {
        DevCopy dc;
        CUDA_TRY(cudaMalloc(&dc.jc, sizeof(JointConst) * s->n));
        CUDA_TRY(cudaMemcpy(dc.jc, s->jc.data(), sizeof(JointConst) * s->n, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&dc.ends, sizeof(int) * s->ends_list.size()));
        CUDA_TRY(cudaMemcpy(dc.ends, s->ends_list.data(), sizeof(int) * s->ends_list.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc(&dc.anc, sizeof(int) * s->anc.size()));
        CUDA_TRY(cudaMemcpy(dc.anc, s->anc.data(), sizeof(int) * s->anc.size(), cudaMemcpyHostToDevice));
        if (!s->pf.empty()) {
            CUDA_TRY(cudaMalloc(&dc.pf, sizeof(PointForce) * s->pf.size()));
            CUDA_TRY(cudaMemcpy(dc.pf, s->pf.data(), sizeof(PointForce) * s->pf.size(), cudaMemcpyHostToDevice));
            const size_t ne = s->pf_ep.empty() ? 1 : s->pf_ep.size();
            CUDA_TRY(cudaMalloc(&dc.pf_ep, sizeof(int) * ne));
            if (!s->pf_ep.empty())
                CUDA_TRY(cudaMemcpy(dc.pf_ep, s->pf_ep.data(), sizeof(int) * s->pf_ep.size(), cudaMemcpyHostToDevice));
        }
        CUDA_TRY(cudaMalloc(&dc.kry, sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(dc.kry, 0, sizeof(unsigned long long)));
        CUDA_TRY(cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking));
        it = s->dev.emplace(dev, dc).first;
    }
    *out = &it->second;
    return RMX_OK;
}

int rmx_dev_reserve(DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return RMX_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    CUDA_TRY(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return RMX_OK;
}

static int dev_reserve(DevBuf& b, size_t bytes) { return rmx_dev_reserve(b, bytes); }

static DevScene make_devscene(const rmx_scene* s, const DevCopy* dc) {
    DevScene ds;
    ds.n = s->n;
    ds.nr = s->nr;
    ds.is_chain = s->is_chain;
    ds.has_ground = s->has_ground;
    for (int i = 0; i < 3; ++i) ds.grav[i] = s->grav[i];
    ds.jc = dc->jc;
    ds.ends_list = dc->ends;
    ds.anc = dc->anc;
    ds.nrounds = s->nrounds;
    ds.pf = dc->pf;
    ds.pf_ep = dc->pf_ep;
    ds.npf = (int)s->pf.size();
    ds.has_chart = s->has_chart;
    return ds;
}

// external-force level of the kernels a scene needs: 0 none, 1 ground contact, 2 ground contact + forces between body points
static int ext_level(const rmx_scene* s) { return !s->pf.empty() ? 2 : (s->has_ground ? 1 : 0); }

static int warps_for(const rmx_scene* s) {
    const int m = s->n > s->nr ? s->n : s->nr;
    if (m <= 32) return 1;
    if (m <= 64) return 2;
    return 4;
}

static size_t scene_smem_doubles(const rmx_scene* s, bool keep = true) {
    const bool g = s->has_ground != 0;
    if (s->impl != 2) return smem_doubles(s->n, s->nr, g);
    if (!s->pf.empty()) return pf_offset_doubles(s->n, s->nr, g, keep) + (size_t)s->pf_doubles;
    return smem_doubles2(s->n, s->nr, g, keep);
}

static int check_opts(const rmx_scene* s, const rmx_opts* o, StepOpts* so, int adjoint) {
    if (!s || !o) return fail(RMX_EINVAL, "null scene/opts");
    if (o->scheme != RMX_SCHEME_BDF1 && o->scheme != RMX_SCHEME_BDF2) return fail(RMX_EINVAL, "opts.scheme must be BDF1 or BDF2");
    if (o->nsteps < 1) return fail(RMX_EINVAL, "opts.nsteps < 1");
    if (!(o->h > 0)) return fail(RMX_EINVAL, "opts.h <= 0");
    if (o->linsolve != RMX_LINSOLVE_LU && o->linsolve != RMX_LINSOLVE_PCG) return fail(RMX_EINVAL, "opts.linsolve");
    so->scheme = o->scheme;
    so->nsteps = o->nsteps;
    so->iterMax = (o->iterMaxFactor > 0 ? o->iterMaxFactor : (adjoint ? 5 : 10)) * s->nr;
    so->iterLsMax = o->iterLsMax > 0 ? o->iterLsMax : 20;
    so->tau_mode = o->tau_mode;
    so->adjoint_newton = adjoint;
    {
        const char* e = std::getenv("RMX_NO_SHORTCUTS");  // developer switch: run stalled Newton solves the long way (A/B of bitwise equality)
        so->shortcuts = !(e && e[0] == '1');
    }
    so->h = o->h;
    so->tol = o->tol > 0 ? o->tol : 1e-9;
    so->dxMax = o->dxMax > 0 ? o->dxMax : 1e3;
    so->lin_tol = o->pcg_tol > 0 ? o->pcg_tol : 1e-6;
    so->lin_maxit = o->pcg_maxit > 0 ? o->pcg_maxit : 4 * s->nr;
    return RMX_OK;
}

static thread_local bool g_no_sched = false;  // set while rmx_rollout re-runs rollouts a load-balanced launch gave up on
bool rmx_sched_enabled() {
    const char* e = std::getenv("RMX_SCHED");  // developer switch: RMX_SCHED=0 launches one block per rollout
    return !g_no_sched && !(e && e[0] == '0');
}

// McNaughton wrap-around schedule of B rollouts x nsteps steps over `slots` co-resident blocks: every block gets
// T = ceil(B nsteps / slots) steps; a rollout that does not fit the rest of a block's quota is cut there -- its LAST steps fill
// the tail of that block (and wait), its FIRST steps open the next block's list (and signal).  nsteps <= T, so the two
// parts never overlap in time.
void rmx_build_plan(SchedPlan& p, long long B, int nsteps, long long slots) {
    p.B = B;
    p.nsteps = nsteps;
    p.slots = slots;
    p.seg.clear();
    p.off.assign((size_t)slots + 1, 0);
    const long long T = (B * nsteps + slots - 1) / slots;
    long long m = 0, t = 0;
    for (long long b = 0; b < B; ++b) {
        if (t + nsteps <= T) {
            p.seg.push_back(make_int4((int)b, 0, nsteps, 0));
            t += nsteps;
            if (t == T && b + 1 < B) {
                p.off[++m] = (int)p.seg.size();
                t = 0;
            }
        } else {
            const int d = (int)(T - t), first = nsteps - d;
            p.seg.push_back(make_int4((int)b, first, nsteps, 1));
            p.off[++m] = (int)p.seg.size();
            p.seg.push_back(make_int4((int)b, 0, first, 2));
            t = first;
        }
    }
    for (long long k = m + 1; k <= slots; ++k) p.off[k] = (int)p.seg.size();
}

// the launcher of one kernel instance (rmx_host.h), or null if that combination is not built
static rmx_fwd_launcher fwd_launcher(int impl, int nw, int ground, bool adj, int lin) {
#define X(IMPL, NW, G, A, L) \
    if (impl == IMPL && nw == NW && ground == G && adj == (A != 0) && lin == L) return RMX_FWD_NAME(IMPL, NW, G, A, L);
    RMX_FWD_ALL(X)
#undef X
    return nullptr;
}

// forward rollout with the Krylov linear solve (fast path only: it needs the world-frame fields of rmx_fast.cuh)
static int launch_fwd_pcg(const rmx_scene* s, const RolloutArgs& a, cudaStream_t st) {
    if (s->impl != 2) return fail(RMX_ELIMIT, "linsolve=PCG needs the composite kernels (n <= 64 joints)");
    const int nw = warps_for(s);
    const int g = ext_level(s);
    const size_t smem = (scene_smem_doubles(s, true) + pcg_doubles(s->n, s->nr)) * sizeof(double);
    if (smem > 227 * 1024) return fail(RMX_ELIMIT, "scene does not fit the 227 KB shared memory of one SM");
    rmx_fwd_launcher f = fwd_launcher(2, nw, g, false, 1);
    if (!f) return fail(RMX_ELIMIT, "linsolve=PCG: no kernel for this scene size");
    return f(a, smem, st, nullptr);
}

// Rollouts per block of a one-warp forward launch (lockstep groups, group_barrier in rmx_device.cuh).  The kernels with
// external forces are bound by instruction fetch, and warps that walk through the code together share the fetched lines:
// they run as many warps per block as one SM holds.  RMX_GROUP=<n> overrides (1 = one rollout per block everywhere).
static int fwd_group(const rmx_scene* s, int nw, int ground, size_t smem_per_rollout) {
    if (nw != 1 || s->impl != 2) return 1;
    const char* e = std::getenv("RMX_GROUP");  // read per launch: tests compare group sizes within one process
    const int forced = e ? std::atoi(e) : 0;
    int G = forced > 0 ? forced : (ground ? RMX_MAX_GROUP : 1);
    const size_t room = 227 * 1024;
    while (G > 1 && (size_t)G * smem_per_rollout > room) --G;
    return G < 1 ? 1 : G;
}

template <bool ADJ>
static int launch_fwd(const rmx_scene* s, const RolloutArgs& a, cudaStream_t st, DevCopy* dc = nullptr) {
    const int nw = warps_for(s);
    const int g = ext_level(s);
    size_t smem = (scene_smem_doubles(s, ADJ) + (ADJ ? 6 * (size_t)s->nr : 0)) * sizeof(double);
    // the one-warp adjoint forward kernel of a scene without external forces runs on the tensor-core path (TcLayoutA)
    if (ADJ && s->impl == 2 && nw == 1 && tc_adjoint(s->n, s->nr, g != 0)) smem = (size_t)TcLayoutA::TOTAL * sizeof(double);
    // developer knob for residency experiments (tools/residency_ab.sh): unused shared memory per block, fewer blocks per SM
    static const size_t pad = [] { const char* e = std::getenv("RMX_DEBUG_SMEM_PAD"); return e ? (size_t)std::atol(e) : (size_t)0; }();
    smem += pad & ~(size_t)15;
    if (smem > 227 * 1024) return fail(RMX_ELIMIT, "scene does not fit the 227 KB shared memory of one SM");
    rmx_fwd_launcher f = fwd_launcher(s->impl, nw, g, ADJ, 0);
    if (!f) return fail(RMX_ELIMIT, "no forward kernel for this scene size");
    RolloutArgs ag = a;
    ag.group = ADJ ? 1 : fwd_group(s, nw, g, smem);
    return f(ag, smem, st, dc);
}

// co-resident blocks of the forward kernel this scene runs on the current device (0 if unknown)
static long long fwd_slots(const rmx_scene* s, DevCopy* dc) {
    RolloutArgs a;
    std::memset(&a, 0, sizeof(a));
    dc->slots_query = 0;
    if (launch_fwd<false>(s, a, nullptr, dc) != RMX_OK) return 0;
    return dc->slots_query;
}

static int rollout_dev_impl(rmx_scene* s, const rmx_opts* o, int64_t B, const double* q0, const double* qdot0,
                            const double* tau, double* q_out, double* qdot_out, int32_t* status, int32_t* iters,
                            void* cuda_stream, double* q_host, double* qdot_host) {
    StepOpts so;
    int rc = check_opts(s, o, &so, 0);
    if (rc) return rc;
    if (B < 1 || !q0 || !qdot0 || !q_out || !status) return fail(RMX_EINVAL, "rmx_rollout_dev: bad arguments");
    if (so.tau_mode != RMX_TAU_NONE && !tau) return fail(RMX_EINVAL, "rmx_rollout_dev: tau_mode set but tau == NULL");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DevCopy* dc;
    rc = scene_on_device(s, dev, &dc);
    if (rc) return rc;
    RolloutArgs a;
    std::memset(&a, 0, sizeof(a));
    a.sc = make_devscene(s, dc);
    a.op = so;
    a.B = B;
    a.q0 = q0;
    a.qd0 = qdot0;
    a.tau = tau;
    a.q_out = q_out;
    a.qd_out = qdot_out;
    a.status = status;
    a.iters = iters;
    a.q_host = q_host;
    a.qd_host = qdot_host;
    if (o->linsolve == RMX_LINSOLVE_PCG) {
        CUDA_TRY(cudaMemsetAsync(dc->kry, 0, sizeof(unsigned long long), (cudaStream_t)cuda_stream));
        a.kry_total = dc->kry;
        bool seen = false;
        for (int d2 : s->kry_devs) seen = seen || d2 == dev;
        if (!seen) s->kry_devs.push_back(dev);
        return launch_fwd_pcg(s, a, (cudaStream_t)cuda_stream);
    }
    return launch_fwd<false>(s, a, (cudaStream_t)cuda_stream, dc);
}

extern "C" int rmx_rollout_dev(rmx_scene* s, const rmx_opts* o, int64_t B, const double* q0, const double* qdot0,
                               const double* tau, double* q_out, double* qdot_out, int32_t* status, int32_t* iters,
                               void* cuda_stream) {
    return rollout_dev_impl(s, o, B, q0, qdot0, tau, q_out, qdot_out, status, iters, cuda_stream, nullptr, nullptr);
}

// Staging for pageable outputs (rmx_rollout).  A device-to-host copy into pageable memory runs at half the rate of one into
// page-locked memory and blocks the calling thread (measured on the B200 boxes, profiles/r02_pageable_probe.log: 20.8 against
// 41.7 GB/s, no gain from more threads), while four host threads move page-locked memory to pageable memory at 50 GB/s.  So the
// kernel writes the trajectories into library-owned page-locked memory while it runs (the same mirrored stores a page-locked
// caller buffer gets) and the finished sub-batches are copied on by host threads.  RMX_STAGE=0 keeps the plain copies; so does
// an allocation that fails or a request above 2 GiB per array and device.
static bool stage_enabled() {
    const char* e = std::getenv("RMX_STAGE");
    return !(e && e[0] == '0');
}
static double* stage_reserve(DevCopy* dc, int which, size_t bytes) {
    if (bytes == 0 || bytes > ((size_t)2 << 30)) return nullptr;
    if (dc->stage_cap[which] < bytes) {
        if (dc->stage[which]) cudaFreeHost(dc->stage[which]);
        dc->stage[which] = nullptr;
        dc->stage_cap[which] = 0;
        void* h = nullptr;
        if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        dc->stage[which] = h;
        dc->stage_cap[which] = bytes;
    }
    return (double*)dc->stage[which];
}
static void host_copy_threads(void* dst, const void* src, size_t bytes) {
    const int nt = bytes >= ((size_t)8 << 20) ? 4 : 1;
    if (nt == 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    std::thread th[4];
    int started = 0;
    size_t done_to = 0;  // bytes [0, done_to) are covered by started threads
    for (int i = 0; i < nt; ++i) {
        const size_t lo = (bytes / 64 * i / nt) * 64, hi = i == nt - 1 ? bytes : (bytes / 64 * (i + 1) / nt) * 64;
        try {  // no exception may cross the C ABI: whatever cannot get a thread is copied here
            th[i] = std::thread([=] { std::memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
        } catch (...) {
            break;
        }
        ++started;
        done_to = hi;
    }
    if (done_to < bytes) std::memcpy((char*)dst + done_to, (const char*)src + done_to, bytes - done_to);
    for (int i = 0; i < started; ++i) th[i].join();
}

// Device pointer of a caller's host buffer if the running kernel can store to it directly: page-locked memory
// (cudaHostAlloc / cudaHostRegister; mapped on every device under unified addressing).  Null for pageable memory.
// RMX_ZEROCOPY=0 (developer switch) forces the staged device-to-host copy.
static double* mapped_host_ptr(double* p) {
    const char* e = std::getenv("RMX_ZEROCOPY");
    if (!p || (e && e[0] == '0')) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? (double*)at.devicePointer : nullptr;
}

// Inside the per-device loops of the host-pointer entries an error must not return: the streams of the devices already
// launched are synchronised and the caller's current device restored in the epilogue (their copies and mirrored stores are
// still writing into the caller's buffers).
#define CUDA_LOOP(x)                                                                   \
    {                                                                                  \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            cudaGetLastError();                                                        \
            ret = fail(RMX_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_));    \
            break;                                                                     \
        }                                                                              \
    }

// Host-pointer entry: shards the batch contiguously over o->ngpus devices (no communication during the rollout),
// each device copying its slice of the trajectories straight back into the caller's buffers.
extern "C" int rmx_rollout(rmx_scene* s, const rmx_opts* o, int64_t B, const double* q0, const double* qdot0,
                           const double* tau, double* q_out, double* qdot_out, int32_t* status, int32_t* iters) {
    StepOpts so;
    int rc = check_opts(s, o, &so, 0);
    if (rc) return rc;
    if (B < 1 || !q0 || !qdot0 || !q_out || !status) return fail(RMX_EINVAL, "rmx_rollout: bad arguments");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(RMX_ENOGPU, "rmx_rollout: no CUDA device (there is no CPU fallback)");
    int G = o->ngpus > 0 ? o->ngpus : 1;
    if (G > ndev) G = ndev;
    if (G > B) G = (int)B;
    int cur = 0;
    CUDA_TRY(cudaGetDevice(&cur));
    const int nr = s->nr;
    const size_t per = (size_t)nr * so.nsteps;
    const size_t tau_per = so.tau_mode == RMX_TAU_PER_STEP ? per : (size_t)nr;
    rmx_opts o1 = *o;
    s->kry_devs.clear();
    std::vector<int> devs;
    for (int gi = 0; gi < G; ++gi) devs.push_back(G == 1 ? cur : gi);
    int ret = RMX_OK;
    std::vector<int> nchunk(G, 1), staged(G, 0);
    for (int gi = 0; gi < G && ret == RMX_OK; ++gi) {
        const int64_t b0 = B * gi / G, b1 = B * (gi + 1) / G, nb = b1 - b0;
        CUDA_LOOP(cudaSetDevice(devs[gi]));
        DevCopy* dc;
        ret = scene_on_device(s, devs[gi], &dc);
        if (ret) break;
        size_t sz[7] = {nb * nr * sizeof(double), nb * nr * sizeof(double), tau ? nb * tau_per * sizeof(double) : 0,
                        nb * per * sizeof(double), qdot_out ? nb * per * sizeof(double) : 0, nb * sizeof(int),
                        iters ? 2 * nb * sizeof(int) : 0};
        for (int i = 0; i < 7 && ret == RMX_OK; ++i)
            if (sz[i]) ret = dev_reserve(dc->buf[i], sz[i]);
        if (ret) break;
        cudaStream_t st = dc->stream;
        CUDA_LOOP(cudaMemcpyAsync(dc->buf[0].p, q0 + b0 * nr, sz[0], cudaMemcpyHostToDevice, st));
        CUDA_LOOP(cudaMemcpyAsync(dc->buf[1].p, qdot0 + b0 * nr, sz[1], cudaMemcpyHostToDevice, st));
        if (tau) CUDA_LOOP(cudaMemcpyAsync(dc->buf[2].p, tau + b0 * tau_per, sz[2], cudaMemcpyHostToDevice, st));
        // Page-locked output buffers are written by the kernel itself while it runs (mirrored stores over PCIe, hidden
        // behind the rollout).  Pageable ones (what a MATLAB mxArray is) need a device-to-host copy, which blocks this thread:
        // the batch is launched as up to four sub-batches of at least one resident wave each, all enqueued first, and the
        // copy of each sub-batch (second loop below) runs while the later ones are still integrating.
        double* qh = mapped_host_ptr(q_out + b0 * per);
        double* qdh = qdot_out ? mapped_host_ptr(qdot_out + b0 * per) : nullptr;
        const bool paged = !qh || (qdot_out && !qdh);
        // pageable outputs: through the page-locked staging if it can be had (the kernel then writes it like a caller's
        // page-locked buffer, device pointer = host pointer under unified addressing)
        staged[gi] = 0;
        if (paged && stage_enabled()) {
            double* sq = !qh ? stage_reserve(dc, 0, sz[3]) : nullptr;
            double* sqd = (qdot_out && !qdh) ? stage_reserve(dc, 1, sz[4]) : nullptr;
            if ((qh || sq) && (!qdot_out || qdh || sqd)) {
                if (sq) {
                    qh = mapped_host_ptr(sq);
                    staged[gi] |= 1;
                }
                if (sqd) {
                    qdh = mapped_host_ptr(sqd);
                    staged[gi] |= 2;
                }
                if ((sq && !qh) || (sqd && !qdh)) {  // not mapped after all: plain copies
                    staged[gi] = 0;
                    qh = mapped_host_ptr(q_out + b0 * per);
                    qdh = qdot_out ? mapped_host_ptr(qdot_out + b0 * per) : nullptr;
                }
            }
        }
        int K = 1;
        if (paged && o1.linsolve == RMX_LINSOLVE_LU && !g_no_sched) {
            const long long slots = fwd_slots(s, dc);
            if (slots > 0 && nb >= 2 * slots) K = (int)std::min<long long>(4, nb / slots);
        }
        if (K > 1) {
            if (!dc->copy_stream) CUDA_LOOP(cudaStreamCreateWithFlags(&dc->copy_stream, cudaStreamNonBlocking));
            bool ok = true;
            for (int c = 0; c < K && ok; ++c)
                if (!dc->chunk_done[c]) ok = cudaEventCreateWithFlags(&dc->chunk_done[c], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) {
                cudaGetLastError();
                K = 1;
            }
        }
        nchunk[gi] = K;
        for (int c = 0; c < K && ret == RMX_OK; ++c) {
            const int64_t c0 = nb * c / K, cn = nb * (c + 1) / K - c0;
            ret = rollout_dev_impl(s, &o1, cn, (double*)dc->buf[0].p + c0 * nr, (double*)dc->buf[1].p + c0 * nr,
                                   tau ? (double*)dc->buf[2].p + c0 * tau_per : nullptr, (double*)dc->buf[3].p + c0 * per,
                                   qdot_out ? (double*)dc->buf[4].p + c0 * per : nullptr, (int*)dc->buf[5].p + c0,
                                   iters ? (int*)dc->buf[6].p + 2 * c0 : nullptr, st, qh ? qh + c0 * per : nullptr,
                                   qdh ? qdh + c0 * per : nullptr);
            if (ret == RMX_OK && K > 1) CUDA_LOOP(cudaEventRecord(dc->chunk_done[c], st));
        }
        if (ret) break;
    }
    // device-to-host copies: all devices are integrating by now
    for (int gi = 0; gi < G && ret == RMX_OK; ++gi) {
        const int64_t b0 = B * gi / G, b1 = B * (gi + 1) / G, nb = b1 - b0;
        CUDA_LOOP(cudaSetDevice(devs[gi]));
        DevCopy* dc = &s->dev.find(devs[gi])->second;
        cudaStream_t st = dc->stream;
        double* qh = mapped_host_ptr(q_out + b0 * per);
        double* qdh = qdot_out ? mapped_host_ptr(qdot_out + b0 * per) : nullptr;
        const int K = nchunk[gi];
        if (staged[gi]) {
            // the trajectories are in the staging already when a sub-batch's event has fired: host threads copy them on
            for (int c = 0; c < K && ret == RMX_OK; ++c) {
                const int64_t c0 = nb * c / K, cn = nb * (c + 1) / K - c0;
                CUDA_LOOP(K > 1 ? cudaEventSynchronize(dc->chunk_done[c]) : cudaStreamSynchronize(st));
                if (staged[gi] & 1)
                    host_copy_threads(q_out + (b0 + c0) * per, (const double*)dc->stage[0] + c0 * per, cn * per * sizeof(double));
                else if (!qh)
                    CUDA_LOOP(cudaMemcpyAsync(q_out + (b0 + c0) * per, (double*)dc->buf[3].p + c0 * per,
                                              cn * per * sizeof(double), cudaMemcpyDeviceToHost, st));
                if (staged[gi] & 2)
                    host_copy_threads(qdot_out + (b0 + c0) * per, (const double*)dc->stage[1] + c0 * per,
                                      cn * per * sizeof(double));
                else if (qdot_out && !qdh)
                    CUDA_LOOP(cudaMemcpyAsync(qdot_out + (b0 + c0) * per, (double*)dc->buf[4].p + c0 * per,
                                              cn * per * sizeof(double), cudaMemcpyDeviceToHost, st));
            }
            if (ret) break;
            CUDA_LOOP(cudaMemcpyAsync(status + b0, dc->buf[5].p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
            if (iters) CUDA_LOOP(cudaMemcpyAsync(iters + 2 * b0, dc->buf[6].p, 2 * nb * sizeof(int), cudaMemcpyDeviceToHost, st));
            continue;
        }
        for (int c = 0; c < K && ret == RMX_OK; ++c) {
            const int64_t c0 = nb * c / K, cn = nb * (c + 1) / K - c0;
            cudaStream_t cs = K > 1 ? dc->copy_stream : st;
            if (K > 1) CUDA_LOOP(cudaStreamWaitEvent(cs, dc->chunk_done[c], 0));
            if (!qh)
                CUDA_LOOP(cudaMemcpyAsync(q_out + (b0 + c0) * per, (double*)dc->buf[3].p + c0 * per, cn * per * sizeof(double),
                                          cudaMemcpyDeviceToHost, cs));
            if (qdot_out && !qdh)
                CUDA_LOOP(cudaMemcpyAsync(qdot_out + (b0 + c0) * per, (double*)dc->buf[4].p + c0 * per, cn * per * sizeof(double),
                                          cudaMemcpyDeviceToHost, cs));
        }
        if (ret) break;
        if (K > 1) CUDA_LOOP(cudaStreamSynchronize(dc->copy_stream));
        CUDA_LOOP(cudaMemcpyAsync(status + b0, dc->buf[5].p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (iters) CUDA_LOOP(cudaMemcpyAsync(iters + 2 * b0, dc->buf[6].p, 2 * nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    for (int gi = 0; gi < G; ++gi) {
        cudaSetDevice(devs[gi]);
        auto it = s->dev.find(devs[gi]);
        if (it != s->dev.end()) {
            cudaError_t e = cudaStreamSynchronize(it->second.stream);
            if (e != cudaSuccess && ret == RMX_OK) ret = fail(RMX_ECUDA, std::string("rollout: ") + cudaGetErrorString(e));
        }
    }
    cudaSetDevice(cur);
    if (ret != RMX_OK || g_no_sched) return ret;
    // A load-balanced launch hands the later steps of some rollouts to another block, which waits (bounded) for the earlier
    // steps to be published.  If that wait ever ran out (RMX_ST_SCHED: the blocks were not co-resident, e.g. a shared GPU), the
    // trajectory after the hand-over is not valid: run those rollouts again, one block per rollout.
    std::vector<int64_t> redo;
    for (int64_t b = 0; b < B; ++b)
        if (status[b] & RMX_ST_SCHED) redo.push_back(b);
    if (redo.empty()) return RMX_OK;
    const int64_t R = (int64_t)redo.size();
    std::vector<double> rq0((size_t)R * nr), rqd0((size_t)R * nr), rtau(tau ? (size_t)R * tau_per : 0), rq((size_t)R * per),
        rqd(qdot_out ? (size_t)R * per : 0);
    std::vector<int32_t> rst(R), rit(2 * R);
    for (int64_t i = 0; i < R; ++i) {
        std::memcpy(&rq0[i * nr], q0 + redo[i] * nr, nr * sizeof(double));
        std::memcpy(&rqd0[i * nr], qdot0 + redo[i] * nr, nr * sizeof(double));
        if (tau) std::memcpy(&rtau[i * tau_per], tau + redo[i] * tau_per, tau_per * sizeof(double));
    }
    g_no_sched = true;
    ret = rmx_rollout(s, o, R, rq0.data(), rqd0.data(), tau ? rtau.data() : nullptr, rq.data(), qdot_out ? rqd.data() : nullptr,
                      rst.data(), rit.data());
    g_no_sched = false;
    if (ret != RMX_OK) return ret;
    for (int64_t i = 0; i < R; ++i) {
        std::memcpy(q_out + redo[i] * per, &rq[i * per], per * sizeof(double));
        if (qdot_out) std::memcpy(qdot_out + redo[i] * per, &rqd[i * per], per * sizeof(double));
        status[redo[i]] = rst[i];
        if (iters) {
            iters[2 * redo[i]] = rit[2 * i];
            iters[2 * redo[i] + 1] = rit[2 * i + 1];
        }
    }
    return RMX_OK;
}

// Host-pointer resume (see the header): one launch on the current device, one block per rollout, each running the single
// segment {b, k_begin[b], nsteps} of the schedule mechanism the load-balanced launches use.
extern "C" int rmx_rollout_resume(rmx_scene* s, const rmx_opts* o, int64_t B, const int32_t* k_begin, const int32_t* k_end,
                                  const double* q0, const double* qdot0, const double* tau, double* q_out, double* qdot_out,
                                  int32_t* status, int32_t* iters) {
    StepOpts so;
    int rc = check_opts(s, o, &so, 0);
    if (rc) return rc;
    if (B < 1 || B >= (1ll << 30) || !k_begin || !q0 || !qdot0 || !q_out || !qdot_out || !status)
        return fail(RMX_EINVAL, "rmx_rollout_resume: bad arguments");
    if (o->linsolve != RMX_LINSOLVE_LU) return fail(RMX_EINVAL, "rmx_rollout_resume: LU linear solve only");
    if (so.tau_mode != RMX_TAU_NONE && !tau) return fail(RMX_EINVAL, "rmx_rollout_resume: tau_mode set but tau == NULL");
    for (int64_t b = 0; b < B; ++b) {
        const int ke = k_end ? k_end[b] : so.nsteps;
        if (k_begin[b] < 0 || k_begin[b] > ke || ke > so.nsteps)
            return fail(RMX_EINVAL, "rmx_rollout_resume: need 0 <= k_begin <= k_end <= nsteps");
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(RMX_ENOGPU, "rmx_rollout_resume: no CUDA device (there is no CPU fallback)");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DevCopy* dc;
    rc = scene_on_device(s, dev, &dc);
    if (rc) return rc;
    const int nr = s->nr;
    const size_t per = (size_t)nr * so.nsteps;
    const size_t tau_per = so.tau_mode == RMX_TAU_PER_STEP ? per : (size_t)nr;
    std::vector<int4> seg((size_t)B);
    std::vector<int> off((size_t)B + 1);
    for (int64_t b = 0; b < B; ++b) {
        seg[b] = make_int4((int)b, k_begin[b], k_end ? k_end[b] : so.nsteps, 0);
        off[b] = (int)b;
    }
    off[B] = (int)B;
    size_t sz[7] = {B * nr * sizeof(double), B * nr * sizeof(double), tau ? B * tau_per * sizeof(double) : 0,
                    B * per * sizeof(double), B * per * sizeof(double), B * sizeof(int), 2 * B * sizeof(int)};
    for (int i = 0; i < 7; ++i)
        if (sz[i] && (rc = dev_reserve(dc->buf[i], sz[i]))) return rc;
    const size_t sb = seg.size() * sizeof(int4), ob = off.size() * sizeof(int);
    if ((rc = dev_reserve(dc->buf[13], sb)) || (rc = dev_reserve(dc->buf[14], ob)) ||
        (rc = dev_reserve(dc->buf[15], (size_t)B * sizeof(int))))
        return rc;
    dc->plan = SchedPlan();  // the cached load-balancing plan no longer matches what is in buf[13], buf[14]
    dc->plan_dev = nullptr;
    cudaStream_t st = dc->stream;
    CUDA_TRY(cudaMemcpyAsync(dc->buf[0].p, q0, sz[0], cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dc->buf[1].p, qdot0, sz[1], cudaMemcpyHostToDevice, st));
    if (tau) CUDA_TRY(cudaMemcpyAsync(dc->buf[2].p, tau, sz[2], cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dc->buf[3].p, q_out, sz[3], cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dc->buf[4].p, qdot_out, sz[4], cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dc->buf[13].p, seg.data(), sb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dc->buf[14].p, off.data(), ob, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(dc->buf[5].p, 0, sz[5], st));
    CUDA_TRY(cudaMemsetAsync(dc->buf[6].p, 0, sz[6], st));
    RolloutArgs a;
    std::memset(&a, 0, sizeof(a));
    a.sc = make_devscene(s, dc);
    a.op = so;
    a.B = B;
    a.q0 = (const double*)dc->buf[0].p;
    a.qd0 = (const double*)dc->buf[1].p;
    a.tau = tau ? (const double*)dc->buf[2].p : nullptr;
    a.q_out = (double*)dc->buf[3].p;
    a.qd_out = (double*)dc->buf[4].p;
    a.status = (int*)dc->buf[5].p;
    a.iters = (int*)dc->buf[6].p;
    a.seg = (const int4*)dc->buf[13].p;
    a.seg_off = (const int*)dc->buf[14].p;
    a.cursor = (int*)dc->buf[15].p;  // one list per rollout: claimed once each (claim_segment)
    a.nlists = (int)B;
    CUDA_TRY(cudaMemsetAsync(dc->buf[15].p, 0, (size_t)B * sizeof(int), st));
    rc = launch_fwd<false>(s, a, st, nullptr);  // no DevCopy: the launch keeps our segments, grid = B
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(q_out, dc->buf[3].p, sz[3], cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(qdot_out, dc->buf[4].p, sz[4], cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(status, dc->buf[5].p, sz[5], cudaMemcpyDeviceToHost, st));
    if (iters) CUDA_TRY(cudaMemcpyAsync(iters, dc->buf[6].p, sz[6], cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(RMX_ECUDA, std::string("rollout_resume: ") + cudaGetErrorString(e));
    return RMX_OK;
}

extern "C" int rmx_linsolve_stats(rmx_scene* s, int64_t* krylov_iterations) {
    if (!s || !krylov_iterations) return fail(RMX_EINVAL, "rmx_linsolve_stats: null argument");
    int cur = 0;
    CUDA_TRY(cudaGetDevice(&cur));
    int64_t total = 0;
    for (int d : s->kry_devs) {
        auto it = s->dev.find(d);
        if (it == s->dev.end()) continue;
        unsigned long long v = 0;
        CUDA_TRY(cudaSetDevice(d));
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(&v, it->second.kry, sizeof(v), cudaMemcpyDeviceToHost));
        total += (int64_t)v;
    }
    cudaSetDevice(cur);
    *krylov_iterations = total;
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------------------
// rmx_eval test hook
// ---------------------------------------------------------------------------------------------------
extern "C" int rmx_eval(rmx_scene* s, const double* q, const double* qdot, const double* dqtmp, const double* tau,
                        double cK, double beta, double* g, double* H, double* M, double* D, double* f) {
    if (!s || !q || !qdot || !dqtmp) return fail(RMX_EINVAL, "rmx_eval: null argument");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(RMX_ENOGPU, "rmx_eval: no CUDA device");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DevCopy* dc;
    int rc = scene_on_device(s, dev, &dc);
    if (rc) return rc;
    const int nr = s->nr;
    const size_t v = nr * sizeof(double), m = (size_t)nr * nr * sizeof(double);
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 5 * v + 3 * m));
    double* dq_ = d;
    double* dqd = d + nr;
    double* ddq = d + 2 * nr;
    double* dtau = d + 3 * nr;
    double* dg = d + 4 * nr;
    double* dH = d + 5 * nr;
    double* dM = dH + (size_t)nr * nr;
    double* dD = dM + (size_t)nr * nr;
    cudaMemcpy(dq_, q, v, cudaMemcpyHostToDevice);
    cudaMemcpy(dqd, qdot, v, cudaMemcpyHostToDevice);
    cudaMemcpy(ddq, dqtmp, v, cudaMemcpyHostToDevice);
    if (tau)
        cudaMemcpy(dtau, tau, v, cudaMemcpyHostToDevice);
    else
        cudaMemset(dtau, 0, v);
    EvalArgs a;
    a.sc = make_devscene(s, dc);
    a.q = dq_;
    a.qd = dqd;
    a.dq = ddq;
    a.tau = dtau;
    a.cK = cK;
    a.beta = beta;
    a.g = dg;
    a.H = dH;
    a.M = dM;
    a.D = dD;
    const int nw = warps_for(s);
    const int gr = ext_level(s);
    const size_t smem = scene_smem_doubles(s) * sizeof(double);
    rc = rmx_launch_eval(s->impl, nw, gr, a, smem);
    if (rc == RMX_OK) {
        std::vector<double> hg(nr), hM((size_t)nr * nr);
        cudaMemcpy(hg.data(), dg, v, cudaMemcpyDeviceToHost);
        cudaMemcpy(hM.data(), dM, m, cudaMemcpyDeviceToHost);
        if (g) std::memcpy(g, hg.data(), v);
        if (H) cudaMemcpy(H, dH, m, cudaMemcpyDeviceToHost);
        if (M) std::memcpy(M, hM.data(), m);
        if (D) cudaMemcpy(D, dD, m, cudaMemcpyDeviceToHost);
        if (f) {
            // g = M*dqtmp - cK*f  =>  f = (M*dqtmp - g)/cK
            for (int r = 0; r < nr; ++r) {
                double acc = 0;
                for (int c2 = 0; c2 < nr; ++c2) acc += hM[(size_t)c2 * nr + r] * dqtmp[c2];
                f[r] = (acc - hg[r]) / cK;
            }
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(RMX_ECUDA, cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

extern "C" int rmx_eval_newton(rmx_scene* s, const double* q, const double* qdot, const double* dqtmp, const double* tau,
                               double cK, double beta, double* H, double* dx) {
    if (!s || !q || !qdot || !dqtmp) return fail(RMX_EINVAL, "rmx_eval_newton: null argument");
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(RMX_ENOGPU, "rmx_eval_newton: no CUDA device");
    }
    if (s->impl != 2) return fail(RMX_ELIMIT, "rmx_eval_newton: composite kernels only (n <= 64 joints)");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DevCopy* dc;
    int rc = scene_on_device(s, dev, &dc);
    if (rc) return rc;
    const int nr = s->nr;
    const size_t v = nr * sizeof(double), m = (size_t)nr * nr * sizeof(double);
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 5 * v + m));
    double *dq_ = d, *dqd = d + nr, *ddq = d + 2 * nr, *dtau = d + 3 * nr, *ddx = d + 4 * nr, *dH = d + 5 * nr;
    cudaMemcpy(dq_, q, v, cudaMemcpyHostToDevice);
    cudaMemcpy(dqd, qdot, v, cudaMemcpyHostToDevice);
    cudaMemcpy(ddq, dqtmp, v, cudaMemcpyHostToDevice);
    if (tau)
        cudaMemcpy(dtau, tau, v, cudaMemcpyHostToDevice);
    else
        cudaMemset(dtau, 0, v);
    EvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.sc = make_devscene(s, dc);
    a.q = dq_;
    a.qd = dqd;
    a.dq = ddq;
    a.tau = dtau;
    a.cK = cK;
    a.beta = beta;
    a.H = dH;
    const int nw = warps_for(s);
    const int gr = ext_level(s);
    const size_t smem = scene_smem_doubles(s, false) * sizeof(double);
    rc = rmx_launch_eval_newton(nw, gr, a, ddx, smem);
    if (rc == RMX_OK) {
        if (H) cudaMemcpy(H, dH, m, cudaMemcpyDeviceToHost);
        if (dx) cudaMemcpy(dx, ddx, v, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(RMX_ECUDA, cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

// rmx_eval_krylov test hook: the two operators of the Krylov linear solve at one evaluation point
extern "C" int rmx_eval_krylov(rmx_scene* s, const double* q, const double* qdot, const double* dqtmp, const double* tau,
                               double cK, double beta, const double* x, double* Hx, double* Pinv_x) {
    if (!s || !q || !qdot || !dqtmp || !x) return fail(RMX_EINVAL, "rmx_eval_krylov: null argument");
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(RMX_ENOGPU, "rmx_eval_krylov: no CUDA device");
    }
    if (s->impl != 2) return fail(RMX_ELIMIT, "rmx_eval_krylov: composite kernels only (n <= 64 joints)");
    if (!s->pf.empty()) return fail(RMX_ELIMIT, "rmx_eval_krylov: scenes with forces between body points use the dense operator");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DevCopy* dc;
    int rc = scene_on_device(s, dev, &dc);
    if (rc) return rc;
    const int nr = s->nr;
    const size_t v = nr * sizeof(double);
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 7 * v));
    double *dq_ = d, *dqd = d + nr, *ddq = d + 2 * nr, *dtau = d + 3 * nr, *dx = d + 4 * nr, *dhx = d + 5 * nr, *dpx = d + 6 * nr;
    cudaMemcpy(dq_, q, v, cudaMemcpyHostToDevice);
    cudaMemcpy(dqd, qdot, v, cudaMemcpyHostToDevice);
    cudaMemcpy(ddq, dqtmp, v, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, x, v, cudaMemcpyHostToDevice);
    if (tau)
        cudaMemcpy(dtau, tau, v, cudaMemcpyHostToDevice);
    else
        cudaMemset(dtau, 0, v);
    EvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.sc = make_devscene(s, dc);
    a.q = dq_;
    a.qd = dqd;
    a.dq = ddq;
    a.tau = dtau;
    a.cK = cK;
    a.beta = beta;
    const size_t smem = (scene_smem_doubles(s, true) + pcg_doubles(s->n, s->nr)) * sizeof(double);
    if (smem > 227 * 1024) {
        cudaFree(d);
        return fail(RMX_ELIMIT, "scene does not fit the 227 KB shared memory of one SM");
    }
    rc = rmx_launch_eval_krylov(warps_for(s), ext_level(s), a, dx, dhx, dpx, smem);
    if (rc == RMX_OK) {
        if (Hx) cudaMemcpy(Hx, dhx, v, cudaMemcpyDeviceToHost);
        if (Pinv_x) cudaMemcpy(Pinv_x, dpx, v, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(RMX_ECUDA, cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

// Test hook (host only, no GPU needed): the load-balancing plan a forward launch of B rollouts x nsteps steps would use on
// `slots` co-resident blocks.  seg: 4 ints per segment {rollout, first step, end step, flags (1 wait, 2 signal)}; off: slots+1.
// Returns the number of segments (at most B + slots), or a negative error code if seg_capacity is too small.
extern "C" int rmx_debug_schedule(int64_t B, int32_t nsteps, int64_t slots, int32_t seg_capacity, int32_t* seg, int32_t* off) {
    if (B < 1 || nsteps < 1 || slots < 1 || !seg || !off) return fail(RMX_EINVAL, "rmx_debug_schedule: bad arguments");
    if (B <= slots) return fail(RMX_EINVAL, "rmx_debug_schedule: B <= slots launches one block per rollout (no plan)");
    SchedPlan p;
    rmx_build_plan(p, B, nsteps, slots);
    if ((int64_t)p.seg.size() > seg_capacity) return fail(RMX_EINVAL, "rmx_debug_schedule: seg_capacity too small");
    for (size_t i = 0; i < p.seg.size(); ++i) {
        seg[4 * i] = p.seg[i].x;
        seg[4 * i + 1] = p.seg[i].y;
        seg[4 * i + 2] = p.seg[i].z;
        seg[4 * i + 3] = p.seg[i].w;
    }
    for (size_t i = 0; i < p.off.size(); ++i) off[i] = p.off[i];
    return (int)p.seg.size();
}

#include "rmx_api_adjoint.inc"
#include "rmx_api_multi.inc"
