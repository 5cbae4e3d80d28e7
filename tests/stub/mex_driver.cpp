// tests/stub/mex_driver.cpp -- drives mexFunction of matlab/redmax_mex.cpp through the stub mex.h (test infrastructure).
//   mex_driver create   : flattenScene-style struct of a 3-link chain (chart as int32 AND as double) -> create -> destroy;
//                         bad calls (missing arguments, wrong class, stale handle) must raise redmax:arg, not crash.  No GPU.
//   mex_driver rollout  : the same scene, 'rollout' (B = 4, constant tau for b > 1 too) and 'energies'; prints q(:,end,b).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mex.h"

static mxArray* dmat(std::vector<mwSize> dims, const std::vector<double>& v) {
    mxArray* a = mxCreateNumericArray(dims.size(), dims.data(), mxDOUBLE_CLASS, mxREAL);
    if (a->numel() != v.size()) { std::fprintf(stderr, "driver: bad dims\n"); std::exit(2); }
    std::memcpy(a->data.data(), v.data(), v.size() * 8);
    return a;
}
static mxArray* imat(std::vector<mwSize> dims, const std::vector<int32_t>& v) {
    mxArray* a = mxCreateNumericArray(dims.size(), dims.data(), mxINT32_CLASS, mxREAL);
    std::memcpy(a->data.data(), v.data(), v.size() * 4);
    return a;
}
static mxArray* str(const char* s) {
    mxArray* a = new mxArray();
    a->cls = mxCHAR_CLASS;
    a->str = s;
    a->dims = {1, std::strlen(s)};
    return a;
}
static mxArray* scal(double v) { return dmat({1, 1}, {v}); }

static mxArray* chain_desc(int n, bool chart_int32) {
    mxArray* d = new mxArray();
    d->cls = mxSTRUCT_CLASS;
    d->dims = {1, 1};
    std::vector<double> parent(n), jtype(n, 1.0), E0_pj(16 * n, 0.0), E0_ji(16 * n, 0.0), axis(3 * n, 0.0), axis2(3 * n, 0.0), I(6 * n),
        sides(3 * n), zeros(n, 0.0), qRest(6 * n, 0.0), lo(n, -1e8), hi(n, 1e8), k(n, 1e8);
    for (int j = 0; j < n; ++j) {
        parent[j] = j - 1;
        for (int i = 0; i < 4; ++i) E0_pj[16 * j + 5 * i] = E0_ji[16 * j + 5 * i] = 1.0;
        if (j > 0) E0_pj[16 * j + 12] = 10.0;  // trans([10 0 0]) (scenesRedMax.m scene 0 pattern)
        E0_ji[16 * j + 12] = 5.0;
        axis[3 * j + 1] = 1.0;
        axis2[3 * j + 1] = 1.0;
        const double s[3] = {10, 1, 1}, m = 10.0;
        for (int i = 0; i < 3; ++i) sides[3 * j + i] = s[i];
        I[6 * j + 0] = m / 12 * (s[1] * s[1] + s[2] * s[2]);
        I[6 * j + 1] = m / 12 * (s[0] * s[0] + s[2] * s[2]);
        I[6 * j + 2] = m / 12 * (s[0] * s[0] + s[1] * s[1]);
        I[6 * j + 3] = I[6 * j + 4] = I[6 * j + 5] = m;
    }
    const mwSize N = n;
    d->fields["parent"] = dmat({1, N}, parent);
    d->fields["jtype"] = dmat({1, N}, jtype);
    d->fields["E0_pj"] = dmat({4, 4, N}, E0_pj);
    d->fields["E0_ji"] = dmat({4, 4, N}, E0_ji);
    d->fields["axis"] = dmat({3, N}, axis);
    d->fields["axis2"] = dmat({3, N}, axis2);
    d->fields["I_i"] = dmat({6, N}, I);
    d->fields["sides"] = dmat({3, N}, sides);
    d->fields["stiffness"] = dmat({1, N}, zeros);
    d->fields["damping"] = dmat({1, N}, zeros);
    d->fields["qRest"] = dmat({6, N}, qRest);
    d->fields["qLimL"] = dmat({1, N}, lo);
    d->fields["qLimU"] = dmat({1, N}, hi);
    d->fields["qLimK"] = dmat({1, N}, k);
    d->fields["qLimD"] = dmat({1, N}, zeros);
    d->fields["grav"] = dmat({3, 1}, {0, 0, -980});
    d->fields["chart"] = chart_int32 ? imat({1, N}, std::vector<int32_t>(n, 0)) : dmat({1, N}, zeros);
    for (const char* e : {"pf_body1", "cable_npts", "ground_body"}) d->fields[e] = dmat({0, 0}, {});
    return d;
}

static int expect_error(const char* what, int nrhs, const mxArray** prhs) {
    mxArray* out[4] = {nullptr, nullptr, nullptr, nullptr};
    try {
        mexFunction(1, out, nrhs, prhs);
    } catch (const mex_error& e) {
        std::printf("ok: %s -> %s: %s\n", what, e.id.c_str(), e.what());
        return 0;
    }
    std::printf("FAIL: %s raised nothing\n", what);
    return 1;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "create";
    int bad = 0;
    try {
        mxArray* h[2] = {nullptr, nullptr};
        for (int v = 0; v < 2; ++v) {
            const mxArray* in[2] = {str("create"), chain_desc(3, v == 0)};
            mxArray* out[1] = {nullptr};
            mexFunction(1, out, 2, in);
            h[v] = out[0];
            std::printf("create (chart as %s): handle class uint64 = %d\n", v == 0 ? "int32" : "double", (int)mxIsUint64(h[v]));
        }
        {
            const mxArray* in1[1] = {str("create")};
            bad += expect_error("create without desc", 1, in1);
            const mxArray* in2[2] = {str("rollout"), h[0]};
            bad += expect_error("rollout without arguments", 2, in2);
            mxArray* d = chain_desc(3, false);
            d->fields["axis"] = imat({3, 3}, std::vector<int32_t>(9, 0));
            const mxArray* in3[2] = {str("create"), d};
            bad += expect_error("create with int32 axis", 2, in3);
            const mxArray* in4[2] = {str("destroy"), scal(12345.0)};
            bad += expect_error("destroy with a double handle", 2, in4);
        }
        if (mode == "rollout") {
            const int nr = 3, B = 4, ns = 10;
            mxArray* o = new mxArray();
            o->cls = mxSTRUCT_CLASS;
            o->dims = {1, 1};
            o->fields["scheme"] = scal(1);
            o->fields["nsteps"] = scal(ns);
            o->fields["h"] = scal(1e-2);
            std::vector<double> q0(nr * B), qd0(nr * B, 0.0), tau(nr * B);
            for (int i = 0; i < nr * B; ++i) {
                q0[i] = 0.1 * ((i * 7) % 5) - 0.2;
                tau[i] = 10.0 * (i % 3);
            }
            const mxArray* in[6] = {str("rollout"), h[1], o, dmat({(mwSize)nr, (mwSize)B}, q0), dmat({(mwSize)nr, (mwSize)B}, qd0),
                                    dmat({(mwSize)nr, (mwSize)B}, tau)};
            mxArray* out[4] = {nullptr, nullptr, nullptr, nullptr};
            mexFunction(4, out, 6, in);
            const double* q = mxGetDoubles(out[0]);
            const int32_t* st = mxGetInt32s(out[2]);
            for (int b = 0; b < B; ++b) {
                std::printf("rollout b=%d status=%d q_end =", b, st[b]);
                for (int r = 0; r < nr; ++r) {
                    const double v = q[(size_t)b * ns * nr + (size_t)(ns - 1) * nr + r];
                    std::printf(" %.17g", v);
                    if (!std::isfinite(v)) bad++;
                }
                std::printf("\n");
                if (st[b] != 0) bad++;
            }
            // int32 kbegin through 'resume': continue from step 5 and land on the same end state
            std::vector<int32_t> kb(B, 5);
            const mxArray* inr[9] = {str("resume"), h[1], o, imat({1, (mwSize)B}, kb), in[3], in[4], in[5], out[0], out[1]};
            mxArray* outr[4] = {nullptr, nullptr, nullptr, nullptr};
            mexFunction(4, outr, 9, inr);
            const double* q2 = mxGetDoubles(outr[0]);
            double dmax = 0;
            for (size_t i = 0; i < (size_t)nr * ns * B; ++i) dmax = std::fmax(dmax, std::fabs(q2[i] - q[i]));
            std::printf("resume from step 5 (int32 kbegin): max |dq| = %.3g\n", dmax);
            if (dmax != 0.0) bad++;
        }
        for (int v = 0; v < 2; ++v) {
            const mxArray* in[2] = {str("destroy"), h[v]};
            mexFunction(0, nullptr, 2, in);
        }
        {
            const mxArray* in[2] = {str("destroy"), h[0]};
            bad += expect_error("destroy twice (stale handle)", 2, in);
        }
    } catch (const std::exception& e) {
        std::printf("FAIL: %s\n", e.what());
        return 1;
    }
    std::printf(bad ? "FAIL\n" : "PASS\n");
    return bad ? 1 : 0;
}
