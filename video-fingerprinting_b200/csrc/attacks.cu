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
//   * resize: cv2.resize of uint8 planes, INTER_AREA (down-scaling) and INTER_LINEAR, bit for bit as
//     OpenCV computes them (float area weights accumulated in table order; 11-bit fixed-point bilinear
//     weights with the two-stage truncation of its vertical pass) - the 1080p -> 720p -> 1080p round trip.
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

// ---- cv2.resize ------------------------------------------------------------------------------------
// INTER_AREA, general (non-integer) ratio.  OpenCV builds a table of (source cell, weight) per
// destination column/row - a partial cell on the left when more than 1e-3 of it is covered, the fully
// covered cells, a partial cell on the right - with the geometry in double and the weights rounded
// to float, then accumulates  buf = sum_x S*alpha  and  sum = sum_y beta*buf  in table order with
// separate float multiplies and adds.  Each thread rebuilds the few table entries of its own pixel.
struct AreaSpan {
    int s1, s2;            // fully covered cells [s1, s2)
    float left, mid, right;  // weights; left/right < 0 when the partial cell is absent
};

__device__ __forceinline__ AreaSpan area_span(int d, double scale, int ssize) {
    const double f1 = d * scale, f2 = f1 + scale;
    const double cell = fmin(scale, (double)ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    AreaSpan a;
    a.s1 = s1;
    a.s2 = s2;
    a.left = (s1 - f1 > 1e-3) ? (float)((s1 - f1) / cell) : -1.0f;
    a.mid = (float)(1.0 / cell);
    a.right = (f2 - s2 > 1e-3) ? (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell) : -1.0f;
    return a;
}

__device__ __forceinline__ float area_row(const uint8_t* __restrict__ row, const AreaSpan& ax) {
    float buf = 0.0f;
    if (ax.left >= 0.0f) buf = __fadd_rn(buf, __fmul_rn((float)row[ax.s1 - 1], ax.left));
    for (int sx = ax.s1; sx < ax.s2; ++sx) buf = __fadd_rn(buf, __fmul_rn((float)row[sx], ax.mid));
    if (ax.right >= 0.0f) buf = __fadd_rn(buf, __fmul_rn((float)row[ax.s2], ax.right));
    return buf;
}

__global__ void __launch_bounds__(256) resize_area_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          long long s_frame, unsigned s_pitch, int sh, int sw,
                                                          long long d_frame, unsigned d_pitch, int dh, int dw,
                                                          double scale_x, double scale_y) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dw) return;
    const uint8_t* s = src + blockIdx.z * s_frame;
    const AreaSpan ax = area_span(dx, scale_x, sw), ay = area_span(dy, scale_y, sh);
    float sum = 0.0f;
    bool first = true;
    auto add_row = [&](int sy, float beta) {
        const float v = __fmul_rn(beta, area_row(s + (unsigned long long)sy * s_pitch, ax));
        sum = first ? v : __fadd_rn(sum, v);
        first = false;
    };
    if (ay.left >= 0.0f) add_row(ay.s1 - 1, ay.left);
    for (int sy = ay.s1; sy < ay.s2; ++sy) add_row(sy, ay.mid);
    if (ay.right >= 0.0f) add_row(ay.s2, ay.right);
    dst[blockIdx.z * d_frame + (unsigned long long)dy * d_pitch + dx] =
        (uint8_t)min(max(__float2int_rn(sum), 0), 255);
}

// INTER_AREA with integer ratios: OpenCV sums the ix*iy cell in integers and either shifts
// ((sum + 2) >> 2 for 2x2) or multiplies by float(1/area) and rounds half to even.
__global__ void __launch_bounds__(256) resize_area_int_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                              long long s_frame, unsigned s_pitch, long long d_frame,
                                                              unsigned d_pitch, int dw, int ix, int iy) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dw) return;
    const uint8_t* s = src + blockIdx.z * s_frame + (unsigned long long)(dy * iy) * s_pitch + dx * ix;
    int sum = 0;
    for (int y = 0; y < iy; ++y)
        for (int x = 0; x < ix; ++x) sum += s[(unsigned long long)y * s_pitch + x];
    int v;
    if (ix == 2 && iy == 2) v = (sum + 2) >> 2;
    else v = __float2int_rn(__fmul_rn((float)sum, 1.0f / (float)(ix * iy)));
    dst[blockIdx.z * d_frame + (unsigned long long)dy * d_pitch + dx] = (uint8_t)min(max(v, 0), 255);
}

// INTER_LINEAR on uint8: source coordinate (d + 0.5) * scale - 0.5 in double, rounded to float; weights
// as 11-bit fixed point (round half to even).  Columns clamp the coordinate (weight 0 on the missing
// neighbour); rows keep their weights and clamp the row INDEX instead.  Horizontal pass in exact
// integers, vertical pass  ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
__global__ void __launch_bounds__(256) resize_linear_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                            long long s_frame, unsigned s_pitch, int sh, int sw,
                                                            long long d_frame, unsigned d_pitch, int dh, int dw,
                                                            double scale_x, double scale_y) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dw) return;
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.0f; sx = 0; }
    if (sx >= sw - 1) { fx = 0.0f; sx = sw - 1; }
    const int a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, fx), 2048.0f)), a1 = __float2int_rn(__fmul_rn(fx, 2048.0f));
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, fy), 2048.0f)), b1 = __float2int_rn(__fmul_rn(fy, 2048.0f));
    const uint8_t* s = src + blockIdx.z * s_frame;
    const uint8_t* r0 = s + (unsigned long long)min(max(sy, 0), sh - 1) * s_pitch;
    const uint8_t* r1 = s + (unsigned long long)min(max(sy + 1, 0), sh - 1) * s_pitch;
    const int sx1 = min(sx + 1, sw - 1);
    const int h0 = r0[sx] * a0 + r0[sx1] * a1, h1 = r1[sx] * a0 + r1[sx1] * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    dst[blockIdx.z * d_frame + (unsigned long long)dy * d_pitch + dx] = (uint8_t)min(max(v, 0), 255);
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

int launch_attack_resize(const void* src, const b200wm_plane* sp, void* dst, const b200wm_plane* dp, int interpolation,
                         cudaStream_t stream) {
    int rc = validate_plane(sp);
    if (rc) return rc;
    rc = validate_plane(dp);
    if (rc) return rc;
    if (!src || !dst || src == dst || sp->n_frames != dp->n_frames) return B200WM_ERR_INVALID;
    if (sp->dtype != B200WM_U8 || sp->elem_stride != 1 || dp->dtype != B200WM_U8 || dp->elem_stride != 1) return B200WM_ERR_UNSUPPORTED;
    if (interpolation != B200WM_INTER_LINEAR && interpolation != B200WM_INTER_AREA) return B200WM_ERR_UNSUPPORTED;
    if (sp->n_frames == 0 || dp->height == 0 || dp->width == 0) return B200WM_OK;
    if (sp->height == 0 || sp->width == 0 || dp->height > 65535 || sp->n_frames > 65535) return B200WM_ERR_INVALID;
    const double scale_x = (double)sp->width / dp->width, scale_y = (double)sp->height / dp->height;
    const dim3 grid((dp->width + 255) / 256, dp->height, sp->n_frames);
    const uint8_t* s = (const uint8_t*)src;
    uint8_t* d = (uint8_t*)dst;
    if (interpolation == B200WM_INTER_AREA) {
        // OpenCV treats INTER_AREA as a (modified) bilinear filter when enlarging: not provided here
        if (dp->width > sp->width || dp->height > sp->height) return B200WM_ERR_UNSUPPORTED;
        const int ix = sp->width / dp->width, iy = sp->height / dp->height;
        if (ix * dp->width == sp->width && iy * dp->height == sp->height) {
            resize_area_int_kernel<<<grid, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes,
                                                             dp->frame_stride_bytes, (unsigned)dp->pitch_bytes, dp->width, ix, iy);
            B200WM_LAUNCH_CHECK("resize_area_int_kernel");
        } else {
            resize_area_kernel<<<grid, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes, sp->height,
                                                         sp->width, dp->frame_stride_bytes, (unsigned)dp->pitch_bytes,
                                                         dp->height, dp->width, scale_x, scale_y);
            B200WM_LAUNCH_CHECK("resize_area_kernel");
        }
        return B200WM_OK;
    }
    // cv2 silently turns an exact 2x2 bilinear reduction into INTER_AREA
    if (sp->width == 2 * dp->width && sp->height == 2 * dp->height) {
        resize_area_int_kernel<<<grid, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes,
                                                         dp->frame_stride_bytes, (unsigned)dp->pitch_bytes, dp->width, 2, 2);
        B200WM_LAUNCH_CHECK("resize_area_int_kernel");
        return B200WM_OK;
    }
    resize_linear_kernel<<<grid, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes, sp->height, sp->width,
                                                   dp->frame_stride_bytes, (unsigned)dp->pitch_bytes, dp->height, dp->width,
                                                   scale_x, scale_y);
    B200WM_LAUNCH_CHECK("resize_linear_kernel");
    return B200WM_OK;
}

}  // namespace b200wm
