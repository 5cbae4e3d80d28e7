// Micro-benchmark (developer tool): latency and throughput of DMMA.8x8x4 (mma.sync m8n8k4 f64) against DFMA on sm_100a,
// to decide whether the warp LU / Newton-matrix assembly should use the FP64 tensor path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int ILP>
__global__ void k_dmma(double* out, int iters, long long* cyc) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b, c0[i], c1[i]);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, long long* cyc) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void k_shfl(double* out, int iters, long long* cyc) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i + threadIdx.x;
    int src = (threadIdx.x * 7 + 3) & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = __shfl_sync(0xffffffffu, c[i], src);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <class K>
void run(const char* name, K kern, int ilp, int blocks, int threads, int iters, double* out, long long* cyc) {
    kern<<<blocks, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(out, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double per = (double)c / ((double)iters * ilp);
    double warps = (double)blocks * threads / 32.0;
    printf("%-6s ilp=%2d blocks=%4d thr=%4d : %.2f cycles per instr per warp (block 0), %.3f ms, %.3e warp-instr/s total\n",
           name, ilp, blocks, threads, per, ms, warps * iters * ilp / (ms * 1e-3));
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 32 * 1024);
    cudaMalloc(&cyc, sizeof(long long));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("no device: %s\n", cudaGetErrorString(e)); return 1; }
    const int it = 20000;
    // latency (1 warp, dependent chain) and per-warp ILP scaling
    run("dmma", k_dmma<1>, 1, 1, 32, it, out, cyc);
    run("dmma", k_dmma<2>, 2, 1, 32, it, out, cyc);
    run("dmma", k_dmma<4>, 4, 1, 32, it, out, cyc);
    run("dmma", k_dmma<8>, 8, 1, 32, it, out, cyc);
    run("dfma", k_dfma<1>, 1, 1, 32, it, out, cyc);
    run("dfma", k_dfma<4>, 4, 1, 32, it, out, cyc);
    run("dfma", k_dfma<8>, 8, 1, 32, it, out, cyc);
    run("shfl", k_shfl<1>, 1, 1, 32, it, out, cyc);
    run("shfl", k_shfl<8>, 8, 1, 32, it, out, cyc);
    // throughput: whole chip, 8 warps/SM (the rollout kernel's residency) and 32 warps/SM
    run("dmma", k_dmma<4>, 4, 148, 256, it, out, cyc);
    run("dmma", k_dmma<4>, 4, 148, 1024, it, out, cyc);
    run("dfma", k_dfma<4>, 4, 148, 256, it, out, cyc);
    run("dfma", k_dfma<8>, 8, 148, 1024, it, out, cyc);
    run("shfl", k_shfl<8>, 8, 148, 256, it, out, cyc);
    run("shfl", k_shfl<8>, 8, 148, 1024, it, out, cyc);
    return 0;
}
