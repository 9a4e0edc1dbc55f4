// Distortion "channel" kernels for the robustness study (BASELINE.json config 5).
//
// The reference has no attack module (its only distortions are a JPEG round trip, tests/test.py:99,111,
// and ffmpeg transcodes, src/offmark/video/frame_writer.py:31-37); SURVEY.md §8d defines synthetic
// stand-ins so that bit-error rates can be measured without ffmpeg.  Two of them run here, on uint8
// planes already resident in HBM:
//   * JPEG-like requantisation: per 8x8 block, DCT of (x - 128), quantise/dequantise with the libjpeg
//     luminance table scaled to a quality, IDCT, round, clip - register-resident forward and inverse
//     8-point butterflies (dct8.cuh).
//   * additive noise: x + noise (a caller-supplied float32 field, so that CPU and GPU runs can use the
//     very same samples), round, clip.
// The CPU definitions are in oracle/attacks.py.
#include "common.cuh"
#include "dct8.cuh"

namespace b200wm {

struct QuantTable {
    float q[64];
};

__global__ void __launch_bounds__(128) jpeg_requant_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           long long frame_stride, unsigned pitch, int bx, int nb,
                                                           unsigned long long div_magic, QuantTable t, int frame0) {
    __shared__ float s_q[64];
    if (threadIdx.x < 64) s_q[threadIdx.x] = t.q[threadIdx.x];
    __syncthreads();
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * 128 + threadIdx.x;
    if (c >= (unsigned)nb) return;
    const unsigned by = (unsigned)(((unsigned long long)c * div_magic) >> 40);
    const unsigned bxi = c - by * bx;
    const long long off = frame * frame_stride + (unsigned long long)(by * 8) * pitch + bxi * 8;
    float b[64];
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        const uint8_t* r = src + off + (unsigned long long)y * pitch;
#pragma unroll
        for (int x = 0; x < 8; ++x) b[8 * y + x] = (float)r[x] - 128.0f;
    }
    dct8x8(b);
#pragma unroll
    for (int k = 0; k < 64; ++k) b[k] = rintf(b[k] / s_q[k]) * s_q[k];
    idct8x8(b);
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        uint8_t* r = dst + off + (unsigned long long)y * pitch;
#pragma unroll
        for (int x = 0; x < 8; ++x) r[x] = (uint8_t)__float2int_rn(fminf(fmaxf(rintf(b[8 * y + x] + 128.0f), 0.0f), 255.0f));
    }
}

__global__ void __launch_bounds__(256) add_noise_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                        const float* __restrict__ noise, long long frame_stride,
                                                        unsigned pitch, int height, int width, int n_frames) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_frame = (long long)height * width;
    if (i >= per_frame * n_frames) return;
    const long long f = i / per_frame, p = i - f * per_frame;
    const int y = (int)(p / width), x = (int)(p - (long long)y * width);
    const long long o = f * frame_stride + (unsigned long long)y * pitch + x;
    dst[o] = (uint8_t)__float2int_rn(fminf(fmaxf(rintf((float)src[o] + noise[i]), 0.0f), 255.0f));
}

int validate_plane(const b200wm_plane* pl);

int launch_attack_jpeg(const void* src, void* dst, const b200wm_plane* pl, int quality, cudaStream_t stream) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (!src || !dst || quality < 1 || quality > 100) return B200WM_ERR_INVALID;
    if (pl->dtype != B200WM_U8 || pl->elem_stride != 1) return B200WM_ERR_UNSUPPORTED;
    static const int base[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                                 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;     // libjpeg jpeg_quality_scaling
    QuantTable t;
    for (int k = 0; k < 64; ++k) {
        int q = (base[k] * scale + 50) / 100;
        q = q < 1 ? 1 : (q > 255 ? 255 : q);
        t.q[k] = (float)q;
    }
    const int bx = pl->width / 8, by = pl->height / 8, nb = bx * by;
    if (nb == 0 || pl->n_frames == 0) return B200WM_OK;
    const unsigned long long magic = (1ull << 40) / (unsigned long long)bx + 1ull;
    for (int f0 = 0; f0 < pl->n_frames; f0 += 65535) {
        const dim3 grid((nb + 127) / 128, (unsigned)((pl->n_frames - f0) < 65535 ? (pl->n_frames - f0) : 65535));
        jpeg_requant_kernel<<<grid, 128, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, pl->frame_stride_bytes,
                                                      (unsigned)pl->pitch_bytes, bx, nb, magic, t, f0);
        B200WM_LAUNCH_CHECK("jpeg_requant_kernel");
    }
    return B200WM_OK;
}

int launch_attack_noise(const void* src, void* dst, const b200wm_plane* pl, const float* noise, cudaStream_t stream) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (!src || !dst || !noise) return B200WM_ERR_INVALID;
    if (pl->dtype != B200WM_U8 || pl->elem_stride != 1) return B200WM_ERR_UNSUPPORTED;
    const long long n = (long long)pl->height * pl->width * pl->n_frames;
    if (n == 0) return B200WM_OK;
    add_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, noise,
                                                                      pl->frame_stride_bytes, (unsigned)pl->pitch_bytes,
                                                                      pl->height, pl->width, pl->n_frames);
    B200WM_LAUNCH_CHECK("add_noise_kernel");
    return B200WM_OK;
}

}  // namespace b200wm
