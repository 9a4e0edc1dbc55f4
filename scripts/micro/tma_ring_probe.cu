// Probe: what the strip ring of dwtsvd_tma.cu can move when the consumers do no arithmetic.
//   read : bulk-copy 15 KB strips global -> shared through a 3-slot ring, consumers only acknowledge
//   copy : the same plus a bulk store of every strip to a second buffer
// Gives the bandwidth ceiling of the access pattern itself (HBM + copy engine + barrier round trip),
// against which the extract / embed kernels are judged.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

constexpr int kStages = 3, kConsumerWarps = 4, kThreads = 160;

template <bool kCopy>
__global__ void __launch_bounds__(kThreads) ring_kernel(const uint8_t* src, uint8_t* dst, int total, unsigned strip_bytes, unsigned* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[2 * kStages];
    const unsigned ring = smem_u32(smem), full0 = smem_u32(&bars[0]), done0 = smem_u32(&bars[kStages]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, step = gridDim.x;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(done0 + 8 * s, kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int stage = 0; unsigned parity = 0;
    if (warp == kConsumerWarps) {
        if (lane != 0) return;
        const int pre = kCopy ? kStages - 1 : kStages;
        for (int s = 0; s < pre; ++s) {
            const int i = blockIdx.x + s * step;
            if (i < total) { mbar_expect(full0 + 8 * s, strip_bytes); bulk_load(ring + s * strip_bytes, src + (size_t)i * strip_bytes, strip_bytes, full0 + 8 * s); }
        }
        int refill = kStages - 1;
        for (int i = blockIdx.x; i < total; i += step) {
            mbar_wait(done0 + 8 * stage, parity);
            int nxt, slot;
            if (kCopy) {
                bulk_store(dst + (size_t)i * strip_bytes, ring + stage * strip_bytes, strip_bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                nxt = i + (kStages - 1) * step; slot = refill; refill = stage;
            } else { nxt = i + kStages * step; slot = stage; }
            if (nxt < total) { mbar_expect(full0 + 8 * slot, strip_bytes); bulk_load(ring + slot * strip_bytes, src + (size_t)nxt * strip_bytes, strip_bytes, full0 + 8 * slot); }
            if (++stage == kStages) { stage = 0; parity ^= 1u; }
        }
        if (kCopy) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        return;
    }
    unsigned acc = 0;
    for (int i = blockIdx.x; i < total; i += step) {
        mbar_wait(full0 + 8 * stage, parity);
        acc += smem[stage * strip_bytes + threadIdx.x * 8];      // one touch per thread, no arithmetic
        if (kCopy) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(done0 + 8 * stage);
        if (++stage == kStages) { stage = 0; parity ^= 1u; }
    }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}

int main() {
    const size_t bytes = (size_t)3000 * 135 * 15360;     // 6.2 GB, like the Y planes of the bench
    uint8_t *a, *b; unsigned* sink;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned strips[] = {7680, 15360, 30720};
    for (unsigned strip : strips) {
        const size_t smem = kStages * strip;
        cudaFuncSetAttribute(ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int per_sm = 1; per_sm <= 8; ++per_sm) {
            if ((smem + 1024) * per_sm > 227 * 1024) break;
            for (int mode = 0; mode < 2; ++mode) {
                const int total = (int)(bytes / strip), blocks = sms * per_sm;
                auto launch = [&] {
                    if (mode) ring_kernel<true><<<blocks, kThreads, smem>>>(a, b, total, strip, sink);
                    else ring_kernel<false><<<blocks, kThreads, smem>>>(a, b, total, strip, sink);
                };
                launch(); launch();
                cudaEventRecord(e0);
                for (int r = 0; r < 5; ++r) launch();
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
                printf("%s  strip %5u B  %d CTAs/SM (%3zu KB in the rings per SM): %.3f ms  %.0f GB/s %s\n", mode ? "copy" : "read", strip, per_sm,
                       smem * per_sm / 1024, ms, (mode ? 2.0 : 1.0) * bytes / ms * 1e-6, cudaGetLastError() == cudaSuccess ? "" : "ERROR");
            }
        }
    }
    // plain device-to-device memcpy for reference
    cudaMemcpy(b, a, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("cudaMemcpy D2D: %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * bytes / ms * 1e-6);
    return 0;
}
