// rmx_k_misc.cu -- instances of the test-hook, energy and adjoint-backward kernels behind plain launchers (rmx_host.h)
#include "rmx_host.h"
#include "rmx_tc.cuh"

using namespace rmx;

template <int NW, int GROUND, int IMPL>
static int launch_eval_t(const EvalArgs& a, size_t smem) {
    int rc = rmx_set_smem(eval_kernel<NW, GROUND, IMPL>, smem);
    if (rc) return rc;
    eval_kernel<NW, GROUND, IMPL><<<1, 32 * NW, smem>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RMX_OK;
}

// gr: external-force level of the scene (0 none, 1 ground contact, 2 ground contact + point forces; composite kernels only)
int rmx_launch_eval(int impl, int nw, int gr, const EvalArgs& a, size_t smem) {
    if (impl == 2) {
        if (nw == 1) return gr == 2 ? launch_eval_t<1, 2, 2>(a, smem) : (gr ? launch_eval_t<1, 1, 2>(a, smem) : launch_eval_t<1, 0, 2>(a, smem));
        return gr == 2 ? launch_eval_t<2, 2, 2>(a, smem) : (gr ? launch_eval_t<2, 1, 2>(a, smem) : launch_eval_t<2, 0, 2>(a, smem));
    }
    if (nw == 1) return gr ? launch_eval_t<1, 1, 1>(a, smem) : launch_eval_t<1, 0, 1>(a, smem);
    if (nw == 2) return gr ? launch_eval_t<2, 1, 1>(a, smem) : launch_eval_t<2, 0, 1>(a, smem);
    return gr ? launch_eval_t<4, 1, 1>(a, smem) : launch_eval_t<4, 0, 1>(a, smem);
}

// rmx_eval_newton test hook: H and dx = -H\g through the forward kernel's own assembly + factorisation path
template <int NW, int GROUND>
static int launch_eval_newton_t(const EvalArgs& a, double* dx, size_t smem) {
    int rc = rmx_set_smem(eval_newton_kernel<NW, GROUND>, smem);
    if (rc) return rc;
    eval_newton_kernel<NW, GROUND><<<1, 32 * NW, smem>>>(a, dx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RMX_OK;
}

int rmx_launch_eval_newton(int nw, int gr, const EvalArgs& a, double* dx, size_t smem) {
    if (nw == 1)
        return gr == 2 ? launch_eval_newton_t<1, 2>(a, dx, smem)
                       : (gr ? launch_eval_newton_t<1, 1>(a, dx, smem) : launch_eval_newton_t<1, 0>(a, dx, smem));
    return gr == 2 ? launch_eval_newton_t<2, 2>(a, dx, smem)
                   : (gr ? launch_eval_newton_t<2, 1>(a, dx, smem) : launch_eval_newton_t<2, 0>(a, dx, smem));
}

// rmx_eval_krylov test hook: y = H x (matrix-free) and z = preconditioner^-1 x at one evaluation point
template <int NW, int GROUND>
static int launch_eval_krylov_t(const EvalArgs& a, const double* x, double* hx, double* pinvx, size_t smem) {
    int rc = rmx_set_smem(eval_krylov_kernel<NW, GROUND>, smem);
    if (rc) return rc;
    eval_krylov_kernel<NW, GROUND><<<1, 32 * NW, smem>>>(a, x, hx, pinvx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RMX_OK;
}

int rmx_launch_eval_krylov(int nw, int gr, const EvalArgs& a, const double* x, double* hx, double* pinvx, size_t smem) {
    if (nw == 1) return gr ? launch_eval_krylov_t<1, 1>(a, x, hx, pinvx, smem) : launch_eval_krylov_t<1, 0>(a, x, hx, pinvx, smem);
    return gr ? launch_eval_krylov_t<2, 1>(a, x, hx, pinvx, smem) : launch_eval_krylov_t<2, 0>(a, x, hx, pinvx, smem);
}

template <int NW, int GROUND, int IMPL>
static int launch_energy_t(const EnergyArgs& a, size_t smem) {
    int rc = rmx_set_smem(energies_kernel<NW, GROUND, IMPL>, smem);
    if (rc) return rc;
    energies_kernel<NW, GROUND, IMPL><<<(unsigned)a.B, 32 * NW, smem>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RMX_OK;
}

int rmx_launch_energy(int impl, int nw, int g, const EnergyArgs& a, size_t smem) {
    if (impl == 2) {
        if (nw == 1) return g == 2 ? launch_energy_t<1, 2, 2>(a, smem) : (g ? launch_energy_t<1, 1, 2>(a, smem) : launch_energy_t<1, 0, 2>(a, smem));
        return g == 2 ? launch_energy_t<2, 2, 2>(a, smem) : (g ? launch_energy_t<2, 1, 2>(a, smem) : launch_energy_t<2, 0, 2>(a, smem));
    }
    if (nw == 1) return g ? launch_energy_t<1, 1, 1>(a, smem) : launch_energy_t<1, 0, 1>(a, smem);
    if (nw == 2) return g ? launch_energy_t<2, 1, 1>(a, smem) : launch_energy_t<2, 0, 1>(a, smem);
    return g ? launch_energy_t<4, 1, 1>(a, smem) : launch_energy_t<4, 0, 1>(a, smem);
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

// ---------------------------------------------------------------------------------------------------
// rmx_fp64_probe: sustained DFMA and DMMA.8x8x4 rates of the whole chip (independent chains, 32 warps per SM)
// ---------------------------------------------------------------------------------------------------
template <int ILP>
__global__ void __launch_bounds__(1024) probe_dfma_kernel(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(1024) probe_dmma_kernel(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        c0[i] = i;
        c1[i] = -i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int rmx_fp64_probe(double* dfma_tflops, double* dmma_tflops) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return rmx_fail(RMX_ENOGPU, "rmx_fp64_probe: no CUDA device");
    }
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = 2 * sms, threads = 1024, iters = 4000;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms_fma = 0.f, ms_mma = 0.f;
    for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
        cudaEventRecord(e0);
        probe_dfma_kernel<8><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_fma, e0, e1);
        cudaEventRecord(e0);
        probe_dmma_kernel<4><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_mma, e0, e1);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CUDA_TRY(cudaGetLastError());
    const double lanes = (double)blocks * threads, warps = lanes / 32.0;
    if (dfma_tflops) *dfma_tflops = lanes * iters * 8.0 * 2.0 / (ms_fma * 1e-3) / 1e12;
    if (dmma_tflops) *dmma_tflops = warps * iters * 4.0 * 512.0 / (ms_mma * 1e-3) / 1e12;
    return RMX_OK;
}

extern "C" int rmx_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return rmx_fail(RMX_EINVAL, "rmx_host_register: null buffer");
    CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return RMX_OK;
}
extern "C" int rmx_host_unregister(void* p) {
    if (!p) return rmx_fail(RMX_EINVAL, "rmx_host_unregister: null buffer");
    CUDA_TRY(cudaHostUnregister(p));
    return RMX_OK;
}
