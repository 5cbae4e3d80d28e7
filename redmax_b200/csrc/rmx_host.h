// rmx_host.h -- host-side declarations shared by the translation units of libredmax_b200.so.
//
// The library is built from several .cu files compiled in parallel (one per kernel family, see __graft_entry__.py):
// rmx_api.cu holds the C ABI and all host logic, the rmx_k_*.cu files hold nothing but explicit kernel instantiations
// behind plain launcher functions (RMX_DEFINE_* below).  Nothing here is part of the public interface.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/redmax_b200.h"
#include "rmx_adjoint.cuh"
#include "rmx_rollout.cuh"

int rmx_fail(int code, const std::string& msg);  // records the message for rmx_last_error(), returns code

#define CUDA_TRY(x)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (x);                                                                                \
        if (e_ != cudaSuccess) {                                                                             \
            cudaGetLastError();                                                                              \
            return rmx_fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RMX_ENOGPU : RMX_ECUDA, \
                            std::string(#x) + ": " + cudaGetErrorString(e_));                                \
        }                                                                                                    \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
// Load-balancing plan of a forward launch (see RolloutArgs::seg): cached per device for the last (B, nsteps, slots)
struct SchedPlan {
    long long B = -1;
    int nsteps = -1;
    long long slots = -1;
    std::vector<int4> seg;
    std::vector<int> off;
};

struct DevCopy {
    SchedPlan plan;
    rmx::JointConst* jc = nullptr;
    int* ends = nullptr;
    int* anc = nullptr;
    rmx::PointForce* pf = nullptr;
    int* pf_ep = nullptr;
    unsigned long long* kry = nullptr;  // Krylov iteration counter
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // device-to-host copies of finished sub-batches (pageable outputs of rmx_rollout)
    cudaEvent_t chunk_done[4] = {nullptr, nullptr, nullptr, nullptr};
    long long slots_query = 0;           // co-resident blocks of the last forward-kernel occupancy query (launcher called with B <= 0)
    DevBuf buf[16];
    // page-locked, device-mapped staging for pageable outputs of rmx_rollout (q, qdot): the kernel writes them while it runs,
    // host threads copy finished sub-batches on to the caller's arrays
    void* stage[2] = {nullptr, nullptr};
    size_t stage_cap[2] = {0, 0};
    void* plan_dev = nullptr;  // device copy of `plan` currently in buf[13]/buf[14]
};

int rmx_dev_reserve(DevBuf& b, size_t bytes);
void rmx_build_plan(SchedPlan& p, long long B, int nsteps, long long slots);
bool rmx_sched_enabled();

// Dynamic shared memory of a kernel and the shared-memory / L1 split it runs with.  `threads` > 0 (the persistent forward
// kernels): ask for exactly the carve-out that the co-resident blocks need instead of the maximum -- the kernels spill a few
// hundred bytes per thread, and what the blocks do not use of the 256 KB array serves those spills as L1 (measured on B200,
// profiles/r02_carveout_ab.log: headline kernel 21.73 -> 21.24 ms at 196 KB instead of 228 KB).  The hardware rounds the
// request up to a configuration it has.  RMX_DEBUG_CARVEOUT=<percent> overrides.
template <typename K>
static inline int rmx_set_smem(K kernel, size_t bytes, int threads = 0) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    static const int forced = [] { const char* e = std::getenv("RMX_DEBUG_CARVEOUT"); return e ? std::atoi(e) : 0; }();
    int carve = (int)cudaSharedmemCarveoutMaxShared;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    if (forced > 0) {
        carve = forced > 100 ? 100 : forced;
    } else if (threads > 0) {
        int nb = 0, dev = 0, per_sm = 0;
        cudaFuncAttributes fa;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, bytes));
        CUDA_TRY(cudaFuncGetAttributes(&fa, kernel));
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        if (nb > 0 && per_sm > 0) {
            const size_t need = (size_t)nb * (bytes + fa.sharedSizeBytes + 1024);  // 1 KB per block is reserved by the system
            const int pct = (int)((need * 100 + (size_t)per_sm - 1) / (size_t)per_sm);
            if (pct < 100) carve = pct < 1 ? 1 : pct;
        }
    }
    if (carve != (int)cudaSharedmemCarveoutMaxShared)
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------------------
// launchers (defined once each in a rmx_k_*.cu file)
// ---------------------------------------------------------------------------------------------------
typedef int (*rmx_fwd_launcher)(const rmx::RolloutArgs& a, size_t smem, cudaStream_t st, DevCopy* dc);
typedef int (*rmx_eval_launcher)(const rmx::EvalArgs& a, size_t smem);
typedef int (*rmx_evaln_launcher)(const rmx::EvalArgs& a, double* dx, size_t smem);
typedef int (*rmx_energy_launcher)(const rmx::EnergyArgs& a, size_t smem);

// forward rollout kernels: IMPL (1 sweep / 2 composite), NW warps per rollout, GROUND (external forces: 0 none, 1 ground
// contact, 2 ground contact + point forces), ADJ (tape-writing), LIN (0 LU, 1 Krylov)
#define RMX_FWD_NAME(IMPL, NW, G, A, L) rmx_fwd_i##IMPL##_w##NW##_g##G##_a##A##_l##L
#define RMX_DECLARE_FWD(IMPL, NW, G, A, L) \
    int RMX_FWD_NAME(IMPL, NW, G, A, L)(const rmx::RolloutArgs& a, size_t smem, cudaStream_t st, DevCopy* dc);
#define RMX_DEFINE_FWD(IMPL, NW, G, A, L)                                                                    \
    int RMX_FWD_NAME(IMPL, NW, G, A, L)(const rmx::RolloutArgs& a, size_t smem, cudaStream_t st, DevCopy* dc) { \
        return rmx_launch_fwd_t<NW, G, A != 0, IMPL, L>(a, smem, st, dc);                                  \
    }

#define RMX_FWD_ALL(X)                                                                                      \
    X(2, 1, 0, 0, 0) X(2, 1, 1, 0, 0) X(2, 1, 0, 1, 0) X(2, 1, 1, 1, 0) X(2, 1, 2, 0, 0) X(2, 1, 2, 1, 0)    \
    X(2, 2, 0, 0, 0) X(2, 2, 1, 0, 0) X(2, 2, 0, 1, 0) X(2, 2, 1, 1, 0) X(2, 2, 2, 0, 0) X(2, 2, 2, 1, 0)    \
    X(2, 1, 0, 0, 1) X(2, 1, 1, 0, 1) X(2, 2, 0, 0, 1) X(2, 2, 1, 0, 1) X(2, 1, 2, 0, 1) X(2, 2, 2, 0, 1)    \
    X(1, 1, 0, 0, 0) X(1, 1, 1, 0, 0) X(1, 1, 0, 1, 0) X(1, 1, 1, 1, 0)                                      \
    X(1, 2, 0, 0, 0) X(1, 2, 1, 0, 0) X(1, 2, 0, 1, 0) X(1, 2, 1, 1, 0)                                      \
    X(1, 4, 0, 0, 0) X(1, 4, 1, 0, 0) X(1, 4, 0, 1, 0) X(1, 4, 1, 1, 0)
RMX_FWD_ALL(RMX_DECLARE_FWD)

// test hooks and diagnostics (rmx_k_misc.cu)
int rmx_launch_eval(int impl, int nw, int ground, const rmx::EvalArgs& a, size_t smem);
int rmx_launch_eval_newton(int nw, int ground, const rmx::EvalArgs& a, double* dx, size_t smem);
int rmx_launch_eval_krylov(int nw, int ground, const rmx::EvalArgs& a, const double* x, double* hx, double* pinvx, size_t smem);
int rmx_launch_energy(int impl, int nw, int ground, const rmx::EnergyArgs& a, size_t smem);
int rmx_launch_bwd(int nw, const rmx::BwdArgs& a, cudaStream_t st);
