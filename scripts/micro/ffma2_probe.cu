// Probe: issue cost of packed FP32 (FFMA2) against scalar FFMA on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe scripts/micro/ffma2_probe.cu && /tmp/ffma2_probe
// Each thread runs kChains independent dependent-chains; the scalar kernel does 2*kChains FFMA per
// iteration, the packed kernel kChains FFMA2 (same flops).  Reports flops/s and instructions/s.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int kChains>
__global__ void scalar_kernel(float* out, float a, float b) {
    float x[2 * kChains];
#pragma unroll
    for (int k = 0; k < 2 * kChains; ++k) x[k] = threadIdx.x * 0.001f + k;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int k = 0; k < 2 * kChains; ++k) x[k] = fmaf(x[k], a, b);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 2 * kChains; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int kChains>
__global__ void packed_kernel(float* out, float a, float b) {
    float2 x[kChains];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = make_float2(threadIdx.x * 0.001f + k, threadIdx.x * 0.002f + k);
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) x[k] = __ffma2_rn(x[k], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k].x + x[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: packed FMA interleaved with integer work (does FFMA2 free issue slots for the ALU pipe?)
template <int kChains>
__global__ void packed_mixed_kernel(float* out, float a, float b, unsigned m) {
    float2 x[kChains];
    unsigned y[kChains];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int k = 0; k < kChains; ++k) { x[k] = make_float2(threadIdx.x * 0.001f + k, threadIdx.x * 0.002f + k); y[k] = threadIdx.x + k; }
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) { x[k] = __ffma2_rn(x[k], a2, b2); y[k] = __dp4a(y[k], m, y[k]); }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k].x + x[k].y + (float)y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int kChains>
__global__ void scalar_mixed_kernel(float* out, float a, float b, unsigned m) {
    float x[2 * kChains];
    unsigned y[kChains];
#pragma unroll
    for (int k = 0; k < 2 * kChains; ++k) x[k] = threadIdx.x * 0.001f + k;
#pragma unroll
    for (int k = 0; k < kChains; ++k) y[k] = threadIdx.x + k;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) { x[2 * k] = fmaf(x[2 * k], a, b); x[2 * k + 1] = fmaf(x[2 * k + 1], a, b); y[k] = __dp4a(y[k], m, y[k]); }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 2 * kChains; ++k) s += x[k];
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += (float)y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    constexpr int C = 8;
    const double flops = 2.0 * 2 * C * (double)kIters * blocks * threads;
    const double lane_cycles_per_sm_per_clk = 1.0;  // reference only
    (void)lane_cycles_per_sm_per_clk;
    float ms;
    ms = time_ms([&] { scalar_kernel<C><<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("scalar FFMA        : %.3f ms  %.1f TFLOP/s  fma-instr/clk/SM %.2f\n", ms, flops / ms * 1e-9,
           (2.0 * C * kIters * blocks * threads / 32) / (ms * 1e-3 * clk * 1e3 * sms));
    ms = time_ms([&] { packed_kernel<C><<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("packed FFMA2       : %.3f ms  %.1f TFLOP/s  ffma2-instr/clk/SM %.2f\n", ms, flops / ms * 1e-9,
           (1.0 * C * kIters * blocks * threads / 32) / (ms * 1e-3 * clk * 1e3 * sms));
    ms = time_ms([&] { scalar_mixed_kernel<C><<<blocks, threads>>>(out, 1.0001f, 0.5f, 0x01010101u); });
    printf("scalar FFMA + IDP  : %.3f ms  (2 FFMA + 1 IDP per slot) instr/clk/SM %.2f\n", ms,
           (3.0 * C * kIters * blocks * threads / 32) / (ms * 1e-3 * clk * 1e3 * sms));
    ms = time_ms([&] { packed_mixed_kernel<C><<<blocks, threads>>>(out, 1.0001f, 0.5f, 0x01010101u); });
    printf("packed FFMA2 + IDP : %.3f ms  (1 FFMA2 + 1 IDP per slot) instr/clk/SM %.2f\n", ms,
           (2.0 * C * kIters * blocks * threads / 32) / (ms * 1e-3 * clk * 1e3 * sms));
    printf("SMs %d, clock attr %d kHz\n", sms, clk);
    return 0;
}
