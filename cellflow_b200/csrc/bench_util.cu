// bench_util.cu — measurement helpers used by bench.py (not part of the simulation path):
// a live FP32 FMA-pipe peak (the roofline denominator of the pair-force kernel; the driver's
// MEASURED_PEAKS.json only holds HBM and bf16 tensor peaks) and an L2 flush.
#include <cuda_runtime.h>

#include "../../include/cellflow_b200.h"

// 16 independent FMA chains per thread: enough ILP to saturate both FMA pipes of an SMSP.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void flush_kernel(float4* buf, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_float4(v, v, v, v);
}

extern "C" int cf_bench_fp32_peak(int device, double* tflops, double* sm_mhz_effective) {
    if (!tflops) return CF_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CF_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CF_ERR_CUDA;
    float* out = nullptr;
    if (cudaMalloc(&out, 16) != cudaSuccess) return CF_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int blocks = prop.multiProcessorCount * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 8; rep++) {
        cudaEventRecord(e0);
        fp32_peak_kernel<<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16.0 * iters * 256.0 * blocks;
        if (rep >= 2 && ms > 0.f) best = flops / (ms * 1e-3) * 1e-12 > best ? flops / (ms * 1e-3) * 1e-12 : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess || best <= 0.0) return CF_ERR_CUDA;
    *tflops = best;
    // 128 FP32 lanes per SM, 2 flop per FMA
    if (sm_mhz_effective) *sm_mhz_effective = best * 1e12 / (2.0 * 128.0 * prop.multiProcessorCount) * 1e-6;
    return CF_OK;
}

static float4* g_flush = nullptr;
static size_t g_flush_n = 0;
extern "C" int cf_bench_flush_l2(int device, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return CF_ERR_CUDA;
    size_t n = bytes / sizeof(float4);
    if (n > g_flush_n) {
        cudaFree(g_flush);
        g_flush = nullptr;
        if (cudaMalloc(&g_flush, n * sizeof(float4)) != cudaSuccess) return CF_ERR_CUDA;
        g_flush_n = n;
    }
    static float v = 0.f;
    v += 1.0f;
    flush_kernel<<<1184, 256>>>(g_flush, n, v);
    return cudaDeviceSynchronize() == cudaSuccess ? CF_OK : CF_ERR_CUDA;
}

// Declared in engine.cu: the stream and device of a handle (bench helper only).
extern "C" int cf_internal_stream(cf_sim* sim, cudaStream_t* stream, int* device);
extern "C" int cf_bench_flush_l2_async(cf_sim* sim, size_t bytes) {
    cudaStream_t st = nullptr;
    int device = 0;
    if (cf_internal_stream(sim, &st, &device) != CF_OK) return CF_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CF_ERR_CUDA;
    size_t n = bytes / sizeof(float4);
    if (n > g_flush_n) {
        cudaDeviceSynchronize();
        cudaFree(g_flush);
        g_flush = nullptr;
        if (cudaMalloc(&g_flush, n * sizeof(float4)) != cudaSuccess) return CF_ERR_CUDA;
        g_flush_n = n;
    }
    static float v = 0.f;
    v += 1.0f;
    flush_kernel<<<1184, 256, 0, st>>>(g_flush, n, v);
    return cudaGetLastError() == cudaSuccess ? CF_OK : CF_ERR_CUDA;
}
