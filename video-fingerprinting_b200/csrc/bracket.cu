// Colour-space bracket around the frame plugins (sm_100a).
//
// Replaces the cv2.cvtColor calls of src/offmark/video/embedder.py:33-39 and
// src/offmark/video/extractor.py:30-34:
//   uint8 H x W x 3 (c0,c1,c2 - FileDecoder's rgb24 taken as BGR)  ->  float32 YUV, and back with
//   clip(0,255) -> round-half-even -> uint8.
// OpenCV's float path (0.5 chroma offset) evaluates, with fused multiply-adds,
//   Y  = fma(c0, .114, fma(c1, .587, c2*.299));  U = fma(c0 - Y, .492, .5);  V = fma(c2 - Y, .877, .5)
//   c0 = fma(U-.5, 2.032, Y);  c1 = fma(V-.5, -.581, fma(U-.5, -.395, Y));  c2 = fma(V-.5, 1.14, Y)
// which is reproduced bit for bit (checked against cv2 through the oracle in tests/).
// Four pixels per thread: 3 x 32-bit loads / 3 x 128-bit stores (and the reverse).
#include "common.cuh"

namespace b200wm {

__device__ __forceinline__ void bgr_to_yuv(float c0, float c1, float c2, float& y, float& u, float& v) {
    y = fmaf(c0, 0.114f, fmaf(c1, 0.587f, c2 * 0.299f));
    u = fmaf(c0 - y, 0.492f, 0.5f);
    v = fmaf(c2 - y, 0.877f, 0.5f);
}

__device__ __forceinline__ void yuv_to_bgr(float y, float u, float v, float& c0, float& c1, float& c2) {
    const float du = u - 0.5f, dv = v - 0.5f;
    c0 = fmaf(du, 2.032f, y);
    c1 = fmaf(dv, -0.581f, fmaf(du, -0.395f, y));
    c2 = fmaf(dv, 1.14f, y);
}

__device__ __forceinline__ unsigned to_u8(float f) {
    return (unsigned)__float2int_rn(fminf(fmaxf(f, 0.0f), 255.0f));
}

__global__ void __launch_bounds__(256) bgr8_to_yuv32_kernel(const uint8_t* __restrict__ bgr, float* __restrict__ yuv,
                                                            long long n_pixels, int vec_ok) {
    const long long quad = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long p0 = quad * 4;
    if (p0 >= n_pixels) return;
    if (vec_ok && p0 + 4 <= n_pixels) {
        const uint3 w = *reinterpret_cast<const uint3*>(bgr + p0 * 3);   // 12 bytes, 4-byte aligned
        const unsigned words[3] = {w.x, w.y, w.z};
        float out[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float c[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int byte = 3 * k + ch;
                c[ch] = (float)((words[byte >> 2] >> (8 * (byte & 3))) & 0xFFu);
            }
            bgr_to_yuv(c[0], c[1], c[2], out[3 * k], out[3 * k + 1], out[3 * k + 2]);
        }
        float4* o = reinterpret_cast<float4*>(yuv + p0 * 3);
        o[0] = make_float4(out[0], out[1], out[2], out[3]);
        o[1] = make_float4(out[4], out[5], out[6], out[7]);
        o[2] = make_float4(out[8], out[9], out[10], out[11]);
    } else {
        for (long long p = p0; p < n_pixels && p < p0 + 4; ++p) {
            float y, u, v;
            bgr_to_yuv((float)bgr[3 * p], (float)bgr[3 * p + 1], (float)bgr[3 * p + 2], y, u, v);
            yuv[3 * p] = y; yuv[3 * p + 1] = u; yuv[3 * p + 2] = v;
        }
    }
}

__global__ void __launch_bounds__(256) yuv32_to_bgr8_kernel(const float* __restrict__ yuv, uint8_t* __restrict__ bgr,
                                                            long long n_pixels, int vec_ok) {
    const long long quad = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long p0 = quad * 4;
    if (p0 >= n_pixels) return;
    if (vec_ok && p0 + 4 <= n_pixels) {
        const float4* in = reinterpret_cast<const float4*>(yuv + p0 * 3);
        const float4 a = in[0], b = in[1], c = in[2];
        const float f[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        unsigned words[3] = {0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float c3[3];
            yuv_to_bgr(f[3 * k], f[3 * k + 1], f[3 * k + 2], c3[0], c3[1], c3[2]);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int byte = 3 * k + ch;
                words[byte >> 2] |= to_u8(c3[ch]) << (8 * (byte & 3));
            }
        }
        *reinterpret_cast<uint3*>(bgr + p0 * 3) = make_uint3(words[0], words[1], words[2]);
    } else {
        for (long long p = p0; p < n_pixels && p < p0 + 4; ++p) {
            float c0, c1, c2;
            yuv_to_bgr(yuv[3 * p], yuv[3 * p + 1], yuv[3 * p + 2], c0, c1, c2);
            bgr[3 * p] = (uint8_t)to_u8(c0); bgr[3 * p + 1] = (uint8_t)to_u8(c1); bgr[3 * p + 2] = (uint8_t)to_u8(c2);
        }
    }
}

int launch_bgr8_to_yuv32(const uint8_t* bgr, float* yuv, long long n_pixels, cudaStream_t stream) {
    if (!bgr || !yuv || n_pixels < 0) return B200WM_ERR_INVALID;
    if (n_pixels == 0) return B200WM_OK;
    const int vec_ok = ((uintptr_t)bgr % 4 == 0) && ((uintptr_t)yuv % 16 == 0);
    const long long quads = (n_pixels + 3) / 4;
    bgr8_to_yuv32_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, stream>>>(bgr, yuv, n_pixels, vec_ok);
    B200WM_LAUNCH_CHECK("bgr8_to_yuv32_kernel");
    return B200WM_OK;
}

int launch_yuv32_to_bgr8(const float* yuv, uint8_t* bgr, long long n_pixels, cudaStream_t stream) {
    if (!bgr || !yuv || n_pixels < 0) return B200WM_ERR_INVALID;
    if (n_pixels == 0) return B200WM_OK;
    const int vec_ok = ((uintptr_t)bgr % 4 == 0) && ((uintptr_t)yuv % 16 == 0);
    const long long quads = (n_pixels + 3) / 4;
    yuv32_to_bgr8_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, stream>>>(yuv, bgr, n_pixels, vec_ok);
    B200WM_LAUNCH_CHECK("yuv32_to_bgr8_kernel");
    return B200WM_OK;
}

}  // namespace b200wm
