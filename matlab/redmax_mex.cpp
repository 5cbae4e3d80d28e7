// redmax_mex.cpp -- MEX gateway from MATLAB to the C ABI of include/redmax_b200.h.
//
// Build on a machine with MATLAB (mex.h is not present in the build image, so this file is compiled there only):
//     mex -R2018a CXXFLAGS='$CXXFLAGS -std=c++14' -I<repo>/include redmax_mex.cpp -L<repo>/redmax_b200/lib -lredmax_b200
//
// Usage from MATLAB (all arrays double, column-major as MATLAB stores them; batch is the trailing dimension):
//     h        = redmax_mex('create', desc)                    desc: struct produced by +redmax/Scene.flatten
//     [q,qdot,status,iters] = redmax_mex('rollout', h, opts, q0, qdot0, tau)     q0,qdot0: nr x B ; q: nr x nsteps x B
//     [P,dPdp,status,q]     = redmax_mex('adjoint', h, opts, task, q0, qdot0, p, xtarget)
//     [T,V]    = redmax_mex('energies', h, q, qdot)
//     redmax_mex('destroy', h)
// The handle is a uint64 scalar.  Errors from the library are raised as MATLAB errors with rmx_last_error().
#ifdef MATLAB_MEX_FILE
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"
#include "redmax_b200.h"

static std::vector<rmx_scene*> g_scenes;

static void cleanup() {
    for (rmx_scene* s : g_scenes)
        if (s) rmx_scene_destroy(s);
    g_scenes.clear();
}

static const mxArray* field(const mxArray* s, const char* name, bool required = true) {
    const mxArray* f = mxGetField(s, 0, name);
    if (!f && required) mexErrMsgIdAndTxt("redmax:arg", "missing field '%s'", name);
    return f;
}
static double scalar(const mxArray* s, const char* name, double dflt) {
    const mxArray* f = mxGetField(s, 0, name);
    return f ? mxGetScalar(f) : dflt;
}
// Index-like arguments may arrive as double (MATLAB's default) or as int32 (e.g. zeros(1,n,'int32'), int32(k)); under
// -R2018a the typed accessors are only valid on arrays of their own class, so dispatch on the class.
static std::vector<int32_t> to_i32(const mxArray* a) {
    const size_t n = mxGetNumberOfElements(a);
    std::vector<int32_t> v(n);
    if (mxIsDouble(a) && !mxIsComplex(a)) {
        const double* p = mxGetDoubles(a);
        for (size_t i = 0; i < n; ++i) v[i] = (int32_t)p[i];
    } else if (mxIsInt32(a)) {
        const int32_t* p = (const int32_t*)mxGetInt32s(a);
        for (size_t i = 0; i < n; ++i) v[i] = p[i];
    } else if (n > 0) {
        mexErrMsgIdAndTxt("redmax:arg", "index arrays must be double or int32");
    }
    return v;
}
// real double array (or empty) -> its data; anything else is a caller error, not a crash
static const double* dbl(const mxArray* a, const char* what) {
    if (!a || mxIsEmpty(a)) return nullptr;
    if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("redmax:arg", "%s must be a real double array", what);
    return mxGetDoubles(a);
}
static const double* dfield(const mxArray* s, const char* name, bool required = true) {
    const mxArray* f = mxGetField(s, 0, name);
    if (!f && required) mexErrMsgIdAndTxt("redmax:arg", "missing field '%s'", name);
    return f ? dbl(f, name) : nullptr;
}
static void need(int nrhs, int n, const char* usage) {
    if (nrhs < n) mexErrMsgIdAndTxt("redmax:arg", "usage: %s", usage);
}
static void need_size(const mxArray* a, size_t n, const char* what) {
    if (mxGetNumberOfElements(a) != n) mexErrMsgIdAndTxt("redmax:arg", "%s has the wrong number of elements", what);
}
static void check(int rc, const char* what) {
    if (rc != RMX_OK) mexErrMsgIdAndTxt("redmax:lib", "%s failed (%d): %s", what, rc, rmx_last_error());
}
static rmx_scene* handle(const mxArray* a) {
    if (!mxIsUint64(a) || mxGetNumberOfElements(a) != 1) mexErrMsgIdAndTxt("redmax:arg", "scene handle must be a uint64 scalar");
    const uint64_t h = *(const uint64_t*)mxGetData(a);
    rmx_scene* s = (rmx_scene*)(uintptr_t)h;
    for (rmx_scene* p : g_scenes)
        if (p == s && s) return s;
    mexErrMsgIdAndTxt("redmax:arg", "stale or unknown scene handle");
    return nullptr;
}
static rmx_opts opts_from(const mxArray* o, int adjoint) {
    rmx_opts r;
    rmx_opts_default(&r, (int32_t)scalar(o, "scheme", 1), adjoint);
    r.nsteps = (int32_t)scalar(o, "nsteps", r.nsteps);
    r.h = scalar(o, "h", r.h);
    r.tol = scalar(o, "tol", r.tol);
    r.dxMax = scalar(o, "dxMax", r.dxMax);
    r.iterMaxFactor = (int32_t)scalar(o, "iterMaxFactor", r.iterMaxFactor);
    r.iterLsMax = (int32_t)scalar(o, "iterLsMax", r.iterLsMax);
    r.linsolve = (int32_t)scalar(o, "linsolve", r.linsolve);
    r.ngpus = (int32_t)scalar(o, "ngpus", 1);
    return r;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("redmax:arg", "first argument must be a command string");
    char cmd[32];
    mxGetString(prhs[0], cmd, sizeof(cmd));
    static bool locked = false;
    if (!locked) {
        mexLock();
        mexAtExit(cleanup);
        locked = true;
    }
    if (!std::strcmp(cmd, "create")) {
        need(nrhs, 2, "h = redmax_mex('create', desc)");
        const mxArray* d = prhs[1];
        if (!mxIsStruct(d)) mexErrMsgIdAndTxt("redmax:arg", "desc must be a struct");
        rmx_scene_desc sd;
        std::memset(&sd, 0, sizeof(sd));
        std::vector<int32_t> parent = to_i32(field(d, "parent")), jtype = to_i32(field(d, "jtype")), gbody;
        sd.n = (int32_t)parent.size();
        need_size(field(d, "jtype"), parent.size(), "jtype");
        need_size(field(d, "E0_pj"), 16 * parent.size(), "E0_pj");
        need_size(field(d, "E0_ji"), 16 * parent.size(), "E0_ji");
        need_size(field(d, "axis"), 3 * parent.size(), "axis");
        need_size(field(d, "axis2"), 3 * parent.size(), "axis2");
        need_size(field(d, "I_i"), 6 * parent.size(), "I_i");
        need_size(field(d, "sides"), 3 * parent.size(), "sides");
        need_size(field(d, "qRest"), RMX_MAX_JOINT_DOF * parent.size(), "qRest");
        sd.parent = parent.data();
        sd.jtype = jtype.data();
        sd.E0_pj = dfield(d, "E0_pj");  // 4 x 4 x n, column-major == what rmx_scene_desc wants
        sd.E0_ji = dfield(d, "E0_ji");
        sd.axis = dfield(d, "axis");    // 3 x n
        sd.axis2 = dfield(d, "axis2");  // 3 x n
        sd.I_i = dfield(d, "I_i");      // 6 x n
        sd.sides = dfield(d, "sides");  // 3 x n
        sd.stiffness = dfield(d, "stiffness");
        sd.damping = dfield(d, "damping");
        sd.qRest = dfield(d, "qRest");  // RMX_MAX_JOINT_DOF x n
        sd.qLimL = dfield(d, "qLimL");
        sd.qLimU = dfield(d, "qLimU");
        sd.qLimK = dfield(d, "qLimK");
        sd.qLimD = dfield(d, "qLimD");
        need_size(field(d, "grav"), 3, "grav");
        std::memcpy(sd.grav, dfield(d, "grav"), 3 * sizeof(double));
        std::vector<int32_t> chart;
        const mxArray* ch = field(d, "chart", false);  // Euler chart per joint (JointSpherical.chart), optional
        if (ch && mxGetNumberOfElements(ch) == (size_t)sd.n) {
            chart = to_i32(ch);
            sd.chart = chart.data();
        }
        std::vector<int32_t> pfb1, pfb2, pfkind;
        const mxArray* pf1 = field(d, "pf_body1", false);
        if (pf1 && mxGetNumberOfElements(pf1) > 0) {
            pfb1 = to_i32(pf1);
            pfb2 = to_i32(field(d, "pf_body2"));
            sd.npointforce = (int32_t)pfb1.size();
            sd.pf_body1 = pfb1.data();
            sd.pf_body2 = pfb2.data();
            sd.pf_x1 = dfield(d, "pf_x1");  // 3 x npointforce
            sd.pf_x2 = dfield(d, "pf_x2");
            sd.pf_ks = dfield(d, "pf_ks");
            sd.pf_kd = dfield(d, "pf_kd");
            pfkind = to_i32(field(d, "pf_kind"));
            sd.pf_kind = pfkind.data();
            sd.pf_L = dfield(d, "pf_L");
        }
        std::vector<int32_t> cnpts, cbody;
        const mxArray* cn = field(d, "cable_npts", false);
        if (cn && mxGetNumberOfElements(cn) > 0) {
            cnpts = to_i32(cn);
            cbody = to_i32(field(d, "cable_body"));  // RMX_MAX_CABLE_POINTS x ncable
            sd.ncable = (int32_t)cnpts.size();
            sd.cable_npts = cnpts.data();
            sd.cable_body = cbody.data();
            sd.cable_x = dfield(d, "cable_x");  // 3 x RMX_MAX_CABLE_POINTS x ncable
            sd.cable_ks = dfield(d, "cable_ks");
            sd.cable_kd = dfield(d, "cable_kd");
            sd.cable_L = dfield(d, "cable_L");
        }
        const mxArray* gb = field(d, "ground_body", false);
        if (gb && mxGetNumberOfElements(gb) > 0) {
            gbody = to_i32(gb);
            sd.nground = (int32_t)gbody.size();
            sd.ground_body = gbody.data();
            sd.ground_E = dfield(d, "ground_E");
            sd.ground_kn = dfield(d, "ground_kn");
            sd.ground_kt = dfield(d, "ground_kt");
            sd.ground_kd = dfield(d, "ground_kd");
            sd.ground_mu = dfield(d, "ground_mu");
        }
        rmx_scene* s = nullptr;
        check(rmx_scene_create(&sd, &s), "rmx_scene_create");
        g_scenes.push_back(s);
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)s;
    } else if (!std::strcmp(cmd, "destroy")) {
        need(nrhs, 2, "redmax_mex('destroy', h)");
        rmx_scene* s = handle(prhs[1]);
        for (auto& p : g_scenes)
            if (p == s) p = nullptr;
        rmx_scene_destroy(s);
    } else if (!std::strcmp(cmd, "rollout")) {
        need(nrhs, 5, "[q,qdot,status,iters] = redmax_mex('rollout', h, opts, q0, qdot0, tau)");
        rmx_scene* s = handle(prhs[1]);
        rmx_opts o = opts_from(prhs[2], 0);
        const int nr = rmx_scene_nr(s);
        const mwSize B = mxGetN(prhs[3]);
        if (B < 1 || mxGetM(prhs[3]) != (size_t)nr) mexErrMsgIdAndTxt("redmax:arg", "q0 must be nr x B");
        need_size(prhs[4], (size_t)nr * B, "qdot0");
        const mxArray* tau = nrhs > 5 && !mxIsEmpty(prhs[5]) ? prhs[5] : nullptr;
        o.tau_mode = !tau ? RMX_TAU_NONE : (mxGetNumberOfElements(tau) == (size_t)nr * B ? RMX_TAU_CONST : RMX_TAU_PER_STEP);
        if (o.tau_mode == RMX_TAU_PER_STEP) need_size(tau, (size_t)nr * o.nsteps * B, "tau (nr x nsteps x B)");
        const mwSize dims[3] = {(mwSize)nr, (mwSize)o.nsteps, B};
        plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        mxArray* qd = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        mxArray* st = mxCreateNumericMatrix(B, 1, mxINT32_CLASS, mxREAL);
        mxArray* it = mxCreateNumericMatrix(2, B, mxINT32_CLASS, mxREAL);
        check(rmx_rollout(s, &o, (int64_t)B, dbl(prhs[3], "q0"), dbl(prhs[4], "qdot0"), tau ? dbl(tau, "tau") : nullptr,
                          mxGetDoubles(plhs[0]), mxGetDoubles(qd), (int32_t*)mxGetData(st), (int32_t*)mxGetData(it)),
              "rmx_rollout");
        if (nlhs > 1) plhs[1] = qd; else mxDestroyArray(qd);
        if (nlhs > 2) plhs[2] = st; else mxDestroyArray(st);
        if (nlhs > 3) plhs[3] = it; else mxDestroyArray(it);
    } else if (!std::strcmp(cmd, "resume")) {
        // [q, qdot, status, iters] = redmax_mex('resume', h, opts, kbegin, q0, qdot0, tau, q, qdot): continue the rollouts from
        // step kbegin(b) (0-based) with the states in q / qdot (nr x nsteps x B) -- after jroot.reparam() re-expressed a step
        need(nrhs, 9, "[q,qdot,status,iters] = redmax_mex('resume', h, opts, kbegin, q0, qdot0, tau, q, qdot)");
        rmx_scene* s = handle(prhs[1]);
        rmx_opts o = opts_from(prhs[2], 0);
        const int nr = rmx_scene_nr(s);
        const mwSize B = mxGetN(prhs[4]);
        if (B < 1 || mxGetM(prhs[4]) != (size_t)nr) mexErrMsgIdAndTxt("redmax:arg", "q0 must be nr x B");
        need_size(prhs[5], (size_t)nr * B, "qdot0");
        std::vector<int32_t> kb = to_i32(prhs[3]);
        if (kb.size() != (size_t)B) mexErrMsgIdAndTxt("redmax:resume", "kbegin needs one entry per rollout");
        const mxArray* tau = !mxIsEmpty(prhs[6]) ? prhs[6] : nullptr;
        o.tau_mode = !tau ? RMX_TAU_NONE : (mxGetNumberOfElements(tau) == (size_t)nr * B ? RMX_TAU_CONST : RMX_TAU_PER_STEP);
        if (o.tau_mode == RMX_TAU_PER_STEP) need_size(tau, (size_t)nr * o.nsteps * B, "tau (nr x nsteps x B)");
        need_size(prhs[7], (size_t)nr * o.nsteps * B, "q (nr x nsteps x B)");
        need_size(prhs[8], (size_t)nr * o.nsteps * B, "qdot (nr x nsteps x B)");
        dbl(prhs[7], "q");
        dbl(prhs[8], "qdot");
        plhs[0] = mxDuplicateArray(prhs[7]);
        mxArray* qd = mxDuplicateArray(prhs[8]);
        mxArray* st = mxCreateNumericMatrix(B, 1, mxINT32_CLASS, mxREAL);
        mxArray* it = mxCreateNumericMatrix(2, B, mxINT32_CLASS, mxREAL);
        check(rmx_rollout_resume(s, &o, (int64_t)B, kb.data(), nullptr, dbl(prhs[4], "q0"), dbl(prhs[5], "qdot0"),
                                 tau ? dbl(tau, "tau") : nullptr, mxGetDoubles(plhs[0]), mxGetDoubles(qd),
                                 (int32_t*)mxGetData(st), (int32_t*)mxGetData(it)),
              "rmx_rollout_resume");
        if (nlhs > 1) plhs[1] = qd; else mxDestroyArray(qd);
        if (nlhs > 2) plhs[2] = st; else mxDestroyArray(st);
        if (nlhs > 3) plhs[3] = it; else mxDestroyArray(it);
    } else if (!std::strcmp(cmd, "adjoint")) {
        need(nrhs, 8, "[P,dPdp,status,q] = redmax_mex('adjoint', h, opts, task, q0, qdot0, p, xtarget)");
        rmx_scene* s = handle(prhs[1]);
        rmx_opts o = opts_from(prhs[2], 1);
        const mxArray* t = prhs[3];
        rmx_task_pointpos tk;
        std::memset(&tk, 0, sizeof(tk));
        tk.body = (int32_t)scalar(t, "body", 1) - 1;  // MATLAB index -> 0-based
        need_size(field(t, "xlocal"), 3, "task.xlocal");
        std::memcpy(tk.xlocal, dfield(t, "xlocal"), 3 * sizeof(double));
        tk.t_target = scalar(t, "t", 0);
        tk.pscale = scalar(t, "pscale", 1);
        tk.wreg = scalar(t, "wreg", 1);
        tk.wpos = scalar(t, "wpos", 1);
        const int nr = rmx_scene_nr(s);
        const mwSize B = mxGetN(prhs[6]);
        if (B < 1 || mxGetM(prhs[6]) != (size_t)nr) mexErrMsgIdAndTxt("redmax:arg", "p must be nr x B");
        need_size(prhs[4], (size_t)nr * B, "q0");
        need_size(prhs[5], (size_t)nr * B, "qdot0");
        need_size(prhs[7], 3 * (size_t)B, "xtarget (3 x B)");
        plhs[0] = mxCreateDoubleMatrix(B, 1, mxREAL);
        mxArray* G = mxCreateDoubleMatrix(nr, B, mxREAL);
        mxArray* st = mxCreateNumericMatrix(B, 1, mxINT32_CLASS, mxREAL);
        mxArray* q = nullptr;
        if (nlhs > 3) {
            const mwSize dims[3] = {(mwSize)nr, (mwSize)o.nsteps, B};
            q = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        }
        check(rmx_rollout_adjoint(s, &o, &tk, (int64_t)B, dbl(prhs[4], "q0"), dbl(prhs[5], "qdot0"), dbl(prhs[6], "p"),
                                  dbl(prhs[7], "xtarget"), mxGetDoubles(plhs[0]), mxGetDoubles(G), q ? mxGetDoubles(q) : nullptr,
                                  (int32_t*)mxGetData(st)),
              "rmx_rollout_adjoint");
        if (nlhs > 1) plhs[1] = G; else mxDestroyArray(G);
        if (nlhs > 2) plhs[2] = st; else mxDestroyArray(st);
        if (nlhs > 3) plhs[3] = q;
    } else if (!std::strcmp(cmd, "energies")) {
        need(nrhs, 4, "[T,V] = redmax_mex('energies', h, q, qdot)");
        rmx_scene* s = handle(prhs[1]);
        const mwSize B = mxGetN(prhs[2]);
        if (B < 1 || mxGetM(prhs[2]) != (size_t)rmx_scene_nr(s)) mexErrMsgIdAndTxt("redmax:arg", "q must be nr x B");
        need_size(prhs[3], mxGetNumberOfElements(prhs[2]), "qdot");
        plhs[0] = mxCreateDoubleMatrix(B, 1, mxREAL);
        mxArray* V = mxCreateDoubleMatrix(B, 1, mxREAL);
        check(rmx_energies(s, (int64_t)B, dbl(prhs[2], "q"), dbl(prhs[3], "qdot"), mxGetDoubles(plhs[0]), mxGetDoubles(V)),
              "rmx_energies");
        if (nlhs > 1) plhs[1] = V; else mxDestroyArray(V);
    } else {
        mexErrMsgIdAndTxt("redmax:arg", "unknown command '%s'", cmd);
    }
}
#else
// Not a MEX build: nothing to compile (mex.h exists only where MATLAB is installed).
#endif
