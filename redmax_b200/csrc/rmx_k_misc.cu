// rmx_k_misc.cu -- instances of the test-hook, energy and adjoint-backward kernels behind plain launchers (rmx_host.h)
#include "rmx_host.h"

using namespace rmx;

template <int NW, bool GROUND, int IMPL>
static int launch_eval_t(const EvalArgs& a, size_t smem) {
    int rc = rmx_set_smem(eval_kernel<NW, GROUND, IMPL>, smem);
    if (rc) return rc;
    eval_kernel<NW, GROUND, IMPL><<<1, 32 * NW, smem>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RMX_OK;
}

int rmx_launch_eval(int impl, int nw, bool gr, const EvalArgs& a, size_t smem) {
    if (impl == 2) {
        if (nw == 1) return gr ? launch_eval_t<1, true, 2>(a, smem) : launch_eval_t<1, false, 2>(a, smem);
        return gr ? launch_eval_t<2, true, 2>(a, smem) : launch_eval_t<2, false, 2>(a, smem);
    }
    if (nw == 1) return gr ? launch_eval_t<1, true, 1>(a, smem) : launch_eval_t<1, false, 1>(a, smem);
    if (nw == 2) return gr ? launch_eval_t<2, true, 1>(a, smem) : launch_eval_t<2, false, 1>(a, smem);
    return gr ? launch_eval_t<4, true, 1>(a, smem) : launch_eval_t<4, false, 1>(a, smem);
}

// rmx_eval_newton test hook: H and dx = -H\g through the forward kernel's own assembly + factorisation path
template <int NW, bool GROUND>
static int launch_eval_newton_t(const EvalArgs& a, double* dx, size_t smem) {
    int rc = rmx_set_smem(eval_newton_kernel<NW, GROUND>, smem);
    if (rc) return rc;
    eval_newton_kernel<NW, GROUND><<<1, 32 * NW, smem>>>(a, dx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RMX_OK;
}

int rmx_launch_eval_newton(int nw, bool gr, const EvalArgs& a, double* dx, size_t smem) {
    if (nw == 1) return gr ? launch_eval_newton_t<1, true>(a, dx, smem) : launch_eval_newton_t<1, false>(a, dx, smem);
    return gr ? launch_eval_newton_t<2, true>(a, dx, smem) : launch_eval_newton_t<2, false>(a, dx, smem);
}

template <int NW, bool GROUND, int IMPL>
static int launch_energy_t(const EnergyArgs& a, size_t smem) {
    int rc = rmx_set_smem(energies_kernel<NW, GROUND, IMPL>, smem);
    if (rc) return rc;
    energies_kernel<NW, GROUND, IMPL><<<(unsigned)a.B, 32 * NW, smem>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RMX_OK;
}

int rmx_launch_energy(int impl, int nw, bool g, const EnergyArgs& a, size_t smem) {
    if (impl == 2) {
        if (nw == 1) return g ? launch_energy_t<1, true, 2>(a, smem) : launch_energy_t<1, false, 2>(a, smem);
        return g ? launch_energy_t<2, true, 2>(a, smem) : launch_energy_t<2, false, 2>(a, smem);
    }
    if (nw == 1) return g ? launch_energy_t<1, true, 1>(a, smem) : launch_energy_t<1, false, 1>(a, smem);
    if (nw == 2) return g ? launch_energy_t<2, true, 1>(a, smem) : launch_energy_t<2, false, 1>(a, smem);
    return g ? launch_energy_t<4, true, 1>(a, smem) : launch_energy_t<4, false, 1>(a, smem);
}

template <int NW>
static int launch_bwd_t(const BwdArgs& a, cudaStream_t st) {
    const size_t vec = (size_t)(3 * a.nr + 32) * sizeof(double);
    size_t smem2 = (size_t)2 * a.tape.sza * sizeof(double) + vec;
    if (smem2 <= 200 * 1024) {
        int rc = rmx_set_smem(adjoint_bwd_kernel<NW, 2>, smem2);
        if (rc) return rc;
        adjoint_bwd_kernel<NW, 2><<<(unsigned)a.B, 32 * NW, smem2, st>>>(a);
    } else {
        size_t smem1 = (size_t)a.tape.sza * sizeof(double) + vec;
        int rc = rmx_set_smem(adjoint_bwd_kernel<NW, 1>, smem1);
        if (rc) return rc;
        adjoint_bwd_kernel<NW, 1><<<(unsigned)a.B, 32 * NW, smem1, st>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return RMX_OK;
}

int rmx_launch_bwd(int nw, const BwdArgs& a, cudaStream_t st) {
    if (nw == 1) return launch_bwd_t<1>(a, st);
    if (nw == 2) return launch_bwd_t<2>(a, st);
    return launch_bwd_t<4>(a, st);
}
