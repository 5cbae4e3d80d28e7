// rmx_launch.cuh -- launch of one forward-kernel instance (included by the rmx_k_fwd_*.cu files only).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "rmx_host.h"

// One persistent launch of rollout_fwd_kernel<NW, GROUND, ADJ, IMPL, LIN>.  Forward launches whose batch is not a multiple of
// the co-resident blocks run the load-balanced segment schedule (rmx_build_plan); everything else one block per rollout.
template <int NW, int GROUND, bool ADJ, int IMPL, int LIN>
static int rmx_launch_fwd_t(const rmx::RolloutArgs& a0, size_t smem, cudaStream_t st, DevCopy* dc) {
    using namespace rmx;
    auto kernel = rollout_fwd_kernel<NW, GROUND, ADJ, IMPL, LIN>;
    int rc = 0;
    RolloutArgs a = a0;
    // lockstep groups: G rollouts (warps) per block, G regions of `smem` bytes
    const int G = (NW == 1 && !ADJ && LIN == 0 && a.group > 1) ? (a.group < RMX_MAX_GROUP ? a.group : RMX_MAX_GROUP) : 1;
    a.group = G;
    smem = (smem + 15) & ~(size_t)15;  // every slot's region starts 16-byte aligned (double2 pivot-row buffers)
    a.group_stride = (int)(smem / sizeof(double));
    smem *= (size_t)G;
    if (smem > 227 * 1024) return rmx_fail(RMX_ELIMIT, "lockstep group does not fit the shared memory of one SM");
    rc = rmx_set_smem(kernel, smem, 32 * NW * G);
    if (rc) return rc;
    if (a.B <= 0) {  // occupancy query only: how many blocks of this kernel are co-resident on the current device
        if (!dc) return RMX_OK;
        int nb = 0, dev = 0, sms = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 32 * NW * G, smem));
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        dc->slots_query = (long long)nb * sms * G;
        return RMX_OK;
    }
    long long grid = a.B;
    if (!ADJ && LIN == 0 && dc && a.qd_out && a.op.nsteps >= 2 && a.B < (1ll << 30) && rmx_sched_enabled()) {
        int nb = 0, dev = 0, sms = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 32 * NW * G, smem));
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        long long slots = (long long)nb * sms * G;
        // (developer knob: pretend more blocks are co-resident than are, to exercise the waits of cut rollouts and the hand-off
        // of segments whose owner has not started -- tests/test_gpu_long_chains.py)
        if (const char* e = std::getenv("RMX_DEBUG_SLOTS_SCALE")) slots *= std::max(1, std::atoi(e));
        if (slots > 0 && a.B > slots && a.B % slots != 0) {
            SchedPlan& p = dc->plan;
            const bool fresh = !(p.B == a.B && p.nsteps == a.op.nsteps && p.slots == slots);
            if (fresh) rmx_build_plan(p, a.B, a.op.nsteps, slots);
            const size_t sb = p.seg.size() * sizeof(int4), ob = p.off.size() * sizeof(int);
            const size_t fb = ((size_t)a.B + (size_t)slots) * sizeof(int);  // flags[B] | cursor[slots]
            if ((rc = rmx_dev_reserve(dc->buf[13], sb)) || (rc = rmx_dev_reserve(dc->buf[14], ob)) ||
                (rc = rmx_dev_reserve(dc->buf[15], fb)))
                return rc;
            if (fresh || dc->plan_dev != dc->buf[13].p) {
                CUDA_TRY(cudaMemcpyAsync(dc->buf[13].p, p.seg.data(), sb, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaMemcpyAsync(dc->buf[14].p, p.off.data(), ob, cudaMemcpyHostToDevice, st));
                dc->plan_dev = dc->buf[13].p;
            }
            CUDA_TRY(cudaMemsetAsync(dc->buf[15].p, 0, fb, st));
            CUDA_TRY(cudaMemsetAsync(a.status, 0, (size_t)a.B * sizeof(int), st));
            if (a.iters) CUDA_TRY(cudaMemsetAsync(a.iters, 0, 2 * (size_t)a.B * sizeof(int), st));
            a.seg = (const int4*)dc->buf[13].p;
            a.seg_off = (const int*)dc->buf[14].p;
            a.flags = (int*)dc->buf[15].p;
            a.cursor = a.flags + a.B;
            a.nlists = (int)slots;
            grid = slots;
        }
    }
    grid = (grid + G - 1) / G;  // slots -> blocks (slots beyond the batch find no work)
    kernel<<<(unsigned)grid, dim3(32 * NW, G), smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return RMX_OK;
}
