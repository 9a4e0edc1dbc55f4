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

// kVec (planar rows 8-byte aligned): eight 64-bit loads, bytes into the mantissa of 2^15 by PRMT (no conversion
// instruction; the row transform removes the bias AND the JPEG level shift of 128 from the exact integer DC sum,
// dct8.cuh), rounding to the quantiser lattice by adding and removing 1.5 * 2^23 (round half to even, like np.rint),
// output bytes rounded by the same bias and clamped / packed in 16-bit lanes, eight 64-bit stores.
template <bool kVec>
__global__ void __launch_bounds__(128) jpeg_requant_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           long long frame_stride, unsigned pitch, int bx, int nb,
                                                           unsigned long long div_magic, QuantTable t, int frame0) {
    __shared__ float s_q[64], s_qinv[64];
    if (threadIdx.x < 64) {
        s_q[threadIdx.x] = t.q[threadIdx.x];
        s_qinv[threadIdx.x] = 1.0f / t.q[threadIdx.x];
    }
    __syncthreads();
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * 128 + threadIdx.x;
    if (c >= (unsigned)nb) return;
    const unsigned by = (unsigned)(((unsigned long long)c * div_magic) >> 40);
    const unsigned bxi = c - by * bx;
    const long long off = frame * frame_stride + (unsigned long long)(by * 8) * pitch + bxi * 8;
    float b[64];
    if (kVec) {
        uint2 rows[8];
        load_rows_u8(src + off, pitch, rows);
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            b[8 * y + 0] = mid_biased_byte<0>(rows[y].x); b[8 * y + 1] = mid_biased_byte<1>(rows[y].x);
            b[8 * y + 2] = mid_biased_byte<2>(rows[y].x); b[8 * y + 3] = mid_biased_byte<3>(rows[y].x);
            b[8 * y + 4] = mid_biased_byte<0>(rows[y].y); b[8 * y + 5] = mid_biased_byte<1>(rows[y].y);
            b[8 * y + 6] = mid_biased_byte<2>(rows[y].y); b[8 * y + 7] = mid_biased_byte<3>(rows[y].y);
        }
        dct8x8<8 * (32768 + 128)>(b);
    } else {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const uint8_t* r = src + off + (unsigned long long)y * pitch;
#pragma unroll
            for (int x = 0; x < 8; ++x) b[8 * y + x] = (float)r[x] - 128.0f;
        }
        dct8x8(b);
    }
    // np.rint(c / q) * q.  c / q = c * (1/q) plus one Newton step (correctly rounded but for rare last-bit cases, which can
    // only matter when c / q sits on a rounding tie anyway; 64 inline IEEE divisions were 1,300 instructions per block and
    // did not fit the instruction cache); round half to even by adding and removing 1.5 * 2^23 (|c / q| < 2^22).
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        const float q0 = b[k] * s_qinv[k];
        const float q1 = fmaf(fmaf(-q0, s_q[k], b[k]), s_qinv[k], q0);
        b[k] = ((q1 + kBias) - kBias) * s_q[k];
    }
    idct8x8(b);
    if (kVec) {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            float f[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) f[x] = (b[8 * y + x] + 128.0f) + kBias;     // two roundings, like np.rint(res + 128.0)
            stg_stream_u2(dst + off + (unsigned long long)y * pitch, pack_biased_row(f));
        }
    } else {
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            uint8_t* r = dst + off + (unsigned long long)y * pitch;
#pragma unroll
            for (int x = 0; x < 8; ++x) r[x] = (uint8_t)__float2int_rn(fminf(fmaxf(rintf(b[8 * y + x] + 128.0f), 0.0f), 255.0f));
        }
    }
}

// kVec4: width, pitch, frame stride and both base addresses multiples of 4: four samples per thread (one 32-bit load, one
// 128-bit noise load, one 32-bit store); round half to even by the 1.5 * 2^23 bias, clamp in 16-bit lanes.
template <bool kVec4>
__global__ void __launch_bounds__(256) add_noise_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                        const float* __restrict__ noise, long long frame_stride,
                                                        unsigned pitch, int height, int width, int n_frames) {
    const long long per_frame = (long long)height * width;
    if (kVec4) {
        const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (i >= per_frame * n_frames) return;
        const long long f = i / per_frame, p = i - f * per_frame;
        const int y = (int)(p / width), x = (int)(p - (long long)y * width);
        const long long o = f * frame_stride + (unsigned long long)y * pitch + x;
        const unsigned w = *reinterpret_cast<const unsigned*>(src + o);
        const float4 nz = *reinterpret_cast<const float4*>(noise + i);
        // (float)sample + noise, one rounding; kept inside the 16-bit lanes of the clamp below whatever the noise is
        const float b0 = fminf(fmaxf((biased_byte<0>(w) - kBias) + nz.x, -1024.0f), 1024.0f);
        const float b1 = fminf(fmaxf((biased_byte<1>(w) - kBias) + nz.y, -1024.0f), 1024.0f);
        const float b2 = fminf(fmaxf((biased_byte<2>(w) - kBias) + nz.z, -1024.0f), 1024.0f);
        const float b3 = fminf(fmaxf((biased_byte<3>(w) - kBias) + nz.w, -1024.0f), 1024.0f);
        const unsigned e = __viaddmin_s16x2_relu(__byte_perm(__float_as_uint(b0 + kBias), __float_as_uint(b2 + kBias), 0x5410), 0u, 0x00FF00FFu);
        const unsigned od = __viaddmin_s16x2_relu(__byte_perm(__float_as_uint(b1 + kBias), __float_as_uint(b3 + kBias), 0x5410), 0u, 0x00FF00FFu);
        *reinterpret_cast<unsigned*>(dst + o) = __byte_perm(e, od, 0x6240);
        return;
    }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_frame * n_frames) return;
    const long long f = i / per_frame, p = i - f * per_frame;
    const int y = (int)(p / width), x = (int)(p - (long long)y * width);
    const long long o = f * frame_stride + (unsigned long long)y * pitch + x;
    dst[o] = (uint8_t)__float2int_rn(fminf(fmaxf(rintf((float)src[o] + noise[i]), 0.0f), 255.0f));
}

// ---- cv2.resize ------------------------------------------------------------------------------------
// Both filters are separable and their per-column / per-row coefficients depend on the geometry only, so a tiny
// first kernel builds the two tables (dw + dh entries, the double-precision geometry of OpenCV evaluated once per
// column and row instead of once per pixel) and the resampling kernels read them: four output pixels per thread,
// one 32-bit store.  Scratch for the tables comes from the stream-ordered allocator.
//
// INTER_AREA, general (non-integer) ratio.  OpenCV builds a table of (source cell, weight) per
// destination column/row - a partial cell on the left when more than 1e-3 of it is covered, the fully
// covered cells, a partial cell on the right - with the geometry in double and the weights rounded
// to float, then accumulates  buf = sum_x S*alpha  and  sum = sum_y beta*buf  in table order with
// separate float multiplies and adds.
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

// INTER_LINEAR on uint8: source coordinate (d + 0.5) * scale - 0.5 in double, rounded to float; weights
// as 11-bit fixed point (round half to even).  Columns clamp the coordinate (weight 0 on the missing
// neighbour); rows keep their weights and clamp the row INDEX instead.
struct LinearTap {
    int i0, i1;            // source indices of the two taps, clamped
    int w0, w1;            // 11-bit weights
};

__device__ __forceinline__ LinearTap linear_tap(int d, double scale, int ssize, bool clamp_coordinate) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    LinearTap t;
    if (clamp_coordinate) {
        if (s < 0) { f = 0.0f; s = 0; }
        if (s >= ssize - 1) { f = 0.0f; s = ssize - 1; }
        t.i0 = s;
        t.i1 = min(s + 1, ssize - 1);
    } else {
        t.i0 = min(max(s, 0), ssize - 1);
        t.i1 = min(max(s + 1, 0), ssize - 1);
    }
    t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    t.w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    return t;
}

// The same span as a fixed-length tap list (<= 4 taps, padded with zero weights: adding 0 * sample leaves a float
// sum unchanged, so the padded list accumulates to the very value of the variable-length one).  With the tap count
// known per launch the resampling loops have no data-dependent branches.
struct AreaTaps {
    int s[4];
    float w[4];
};

__host__ __device__ inline int area_tap_count(int d, double scale, int ssize) {
    const double f1 = d * scale, f2 = f1 + scale;
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = s2 < ssize - 1 ? s2 : ssize - 1;
    s1 = s1 < s2 ? s1 : s2;
    return (s1 - f1 > 1e-3 ? 1 : 0) + (s2 - s1) + (f2 - s2 > 1e-3 ? 1 : 0);
}

__device__ __forceinline__ AreaTaps area_taps(int d, double scale, int ssize) {
    const AreaSpan a = area_span(d, scale, ssize);
    AreaTaps t;
    int n = 0;
    if (a.left >= 0.0f) { t.s[n] = a.s1 - 1; t.w[n++] = a.left; }
    for (int sx = a.s1; sx < a.s2 && n < 4; ++sx) { t.s[n] = sx; t.w[n++] = a.mid; }
    if (a.right >= 0.0f && n < 4) { t.s[n] = a.s2; t.w[n++] = a.right; }
    const int last = n ? t.s[n - 1] : 0;
    for (; n < 4; ++n) { t.s[n] = last; t.w[n] = 0.0f; }
    return t;
}

__global__ void __launch_bounds__(256) resize_tables_kernel(int area, void* __restrict__ tab_x, void* __restrict__ tab_y, int sw, int sh,
                                                            int dw, int dh, double scale_x, double scale_y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dw + dh) return;
    const bool is_x = i < dw;
    const int d = is_x ? i : i - dw;
    if (area == 2) (is_x ? (AreaTaps*)tab_x : (AreaTaps*)tab_y)[d] = area_taps(d, is_x ? scale_x : scale_y, is_x ? sw : sh);
    else if (area) (is_x ? (AreaSpan*)tab_x : (AreaSpan*)tab_y)[d] = area_span(d, is_x ? scale_x : scale_y, is_x ? sw : sh);
    else (is_x ? (LinearTap*)tab_x : (LinearTap*)tab_y)[d] = linear_tap(d, is_x ? scale_x : scale_y, is_x ? sw : sh, is_x);
}

__device__ __forceinline__ float area_row(const uint8_t* __restrict__ row, const AreaSpan& ax) {
    float buf = 0.0f;
    if (ax.left >= 0.0f) buf = __fadd_rn(buf, __fmul_rn((float)row[ax.s1 - 1], ax.left));
    for (int sx = ax.s1; sx < ax.s2; ++sx) buf = __fadd_rn(buf, __fmul_rn((float)row[sx], ax.mid));
    if (ax.right >= 0.0f) buf = __fadd_rn(buf, __fmul_rn((float)row[ax.s2], ax.right));
    return buf;
}

// four output bytes of one row: a 32-bit store when the address allows it
__device__ __forceinline__ void store_px4(uint8_t* p, const int (&v)[4], int n) {
    if (n == 4 && ((uintptr_t)p & 3) == 0) {
        *reinterpret_cast<unsigned*>(p) = (unsigned)v[0] | ((unsigned)v[1] << 8) | ((unsigned)v[2] << 16) | ((unsigned)v[3] << 24);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < n) p[k] = (uint8_t)v[k];
    }
}

// A CTA owns 1024 output columns x kResizeRows output rows: its slice of the column table is staged in shared memory
// once (table reads from global memory per pixel outweighed the pixels themselves) and reused for every row.
constexpr int kResizeRows = 16;

__global__ void __launch_bounds__(256) resize_area_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          long long s_frame, unsigned s_pitch, long long d_frame, unsigned d_pitch,
                                                          int dw, int dh, const AreaSpan* __restrict__ tab_x, const AreaSpan* __restrict__ tab_y) {
    __shared__ AreaSpan s_tab[1024];
    const int col0 = blockIdx.x * 1024;
    for (int i = threadIdx.x; i < 1024; i += 256) s_tab[i] = tab_x[min(col0 + i, dw - 1)];
    __syncthreads();
    const int dx0 = col0 + threadIdx.x * 4;
    if (dx0 >= dw) return;
    const int n = min(4, dw - dx0);
    const uint8_t* s = src + blockIdx.z * s_frame;
    AreaSpan ax[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ax[k] = s_tab[threadIdx.x * 4 + k];
    const int dy_end = min(dh, (int)(blockIdx.y + 1) * kResizeRows);
    for (int dy = blockIdx.y * kResizeRows; dy < dy_end; ++dy) {
        const AreaSpan ay = tab_y[dy];
        float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        bool first = true;
        auto add_row = [&](int sy, float beta) {
            const uint8_t* row = s + (unsigned long long)sy * s_pitch;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float v = __fmul_rn(beta, area_row(row, ax[k]));
                sum[k] = first ? v : __fadd_rn(sum[k], v);
            }
            first = false;
        };
        if (ay.left >= 0.0f) add_row(ay.s1 - 1, ay.left);
        for (int sy = ay.s1; sy < ay.s2; ++sy) add_row(sy, ay.mid);
        if (ay.right >= 0.0f) add_row(ay.s2, ay.right);
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = min(max(__float2int_rn(sum[k]), 0), 255);
        store_px4(dst + blockIdx.z * d_frame + (unsigned long long)dy * d_pitch + dx0, v, n);
    }
}

// Fixed tap counts (the usual reductions, e.g. 1080p -> 720p needs two taps per axis): straight-line inner loops.
template <int TX, int TY>
__global__ void __launch_bounds__(256) resize_area_taps_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                               long long s_frame, unsigned s_pitch, long long d_frame, unsigned d_pitch,
                                                               int dw, int dh, const AreaTaps* __restrict__ tab_x,
                                                               const AreaTaps* __restrict__ tab_y) {
    __shared__ AreaTaps s_tab[1024];
    const int col0 = blockIdx.x * 1024;
    for (int i = threadIdx.x; i < 1024; i += 256) s_tab[i] = tab_x[min(col0 + i, dw - 1)];
    __syncthreads();
    const int dx0 = col0 + threadIdx.x * 4;
    if (dx0 >= dw) return;
    const int n = min(4, dw - dx0);
    const uint8_t* s = src + blockIdx.z * s_frame;
    int sx[4][TX];
    float wx[4][TX];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < TX; ++j) { sx[k][j] = s_tab[threadIdx.x * 4 + k].s[j]; wx[k][j] = s_tab[threadIdx.x * 4 + k].w[j]; }
    const int dy_end = min(dh, (int)(blockIdx.y + 1) * kResizeRows);
    for (int dy = blockIdx.y * kResizeRows; dy < dy_end; ++dy) {
        const AreaTaps ay = tab_y[dy];
        float sum[4];
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const uint8_t* row = s + (unsigned long long)ay.s[r] * s_pitch;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float buf = 0.0f;
#pragma unroll
                for (int j = 0; j < TX; ++j)      // bytes to float through the mantissa of 2^23: no conversion-pipe instruction
                    buf = __fadd_rn(buf, __fmul_rn(__uint_as_float(0x4B000000u | row[sx[k][j]]) - 8388608.0f, wx[k][j]));
                const float v = __fmul_rn(ay.w[r], buf);
                sum[k] = r == 0 ? v : __fadd_rn(sum[k], v);
            }
        }
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = min(max(__float2int_rn(sum[k]), 0), 255);
        store_px4(dst + blockIdx.z * d_frame + (unsigned long long)dy * d_pitch + dx0, v, n);
    }
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

// INTER_LINEAR (taps: LinearTap above).  Horizontal pass in exact integers, vertical pass
// ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
__global__ void __launch_bounds__(256) resize_linear_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                            long long s_frame, unsigned s_pitch, long long d_frame, unsigned d_pitch,
                                                            int dw, int dh, const LinearTap* __restrict__ tab_x, const LinearTap* __restrict__ tab_y) {
    __shared__ LinearTap s_tab[1024];
    const int col0 = blockIdx.x * 1024;
    for (int i = threadIdx.x; i < 1024; i += 256) s_tab[i] = tab_x[min(col0 + i, dw - 1)];
    __syncthreads();
    const int dx0 = col0 + threadIdx.x * 4;
    if (dx0 >= dw) return;
    const int n = min(4, dw - dx0);
    LinearTap tx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tx[k] = s_tab[threadIdx.x * 4 + k];
    const uint8_t* s = src + blockIdx.z * s_frame;
    const int dy_end = min(dh, (int)(blockIdx.y + 1) * kResizeRows);
    for (int dy = blockIdx.y * kResizeRows; dy < dy_end; ++dy) {
        const LinearTap ty = tab_y[dy];
        const uint8_t* r0 = s + (unsigned long long)ty.i0 * s_pitch;
        const uint8_t* r1 = s + (unsigned long long)ty.i1 * s_pitch;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h0 = r0[tx[k].i0] * tx[k].w0 + r0[tx[k].i1] * tx[k].w1, h1 = r1[tx[k].i0] * tx[k].w0 + r1[tx[k].i1] * tx[k].w1;
            v[k] = min(max((((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2, 0), 255);
        }
        store_px4(dst + blockIdx.z * d_frame + (unsigned long long)dy * d_pitch + dx0, v, n);
    }
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
        const bool vec = (pl->pitch_bytes % 8) == 0 && (pl->frame_stride_bytes % 8) == 0 && ((uintptr_t)src % 8) == 0 && ((uintptr_t)dst % 8) == 0;
        if (vec) jpeg_requant_kernel<true><<<grid, 128, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, pl->frame_stride_bytes,
                                                                    (unsigned)pl->pitch_bytes, bx, nb, magic, t, f0);
        else jpeg_requant_kernel<false><<<grid, 128, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, pl->frame_stride_bytes,
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
    const bool vec4 = pl->width % 4 == 0 && pl->pitch_bytes % 4 == 0 && pl->frame_stride_bytes % 4 == 0 && (uintptr_t)src % 4 == 0 &&
                      (uintptr_t)dst % 4 == 0 && (uintptr_t)noise % 16 == 0;
    if (vec4)
        add_noise_kernel<true><<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, noise,
                                                                                 pl->frame_stride_bytes, (unsigned)pl->pitch_bytes,
                                                                                 pl->height, pl->width, pl->n_frames);
    else
        add_noise_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const uint8_t*)src, (uint8_t*)dst, noise,
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
    const dim3 grid((dp->width + 255) / 256, dp->height, sp->n_frames);          // one output pixel per thread
    const dim3 grid4((dp->width + 1023) / 1024, (dp->height + kResizeRows - 1) / kResizeRows, sp->n_frames);   // 4 pixels x kResizeRows rows per thread
    const uint8_t* s = (const uint8_t*)src;
    uint8_t* d = (uint8_t*)dst;
    auto with_tables = [&](int mode, size_t entry) -> int {      // mode: 0 linear, 1 area (spans), 2 + 16 tx + 64 ty area with fixed tap counts
        const int area = mode & 3, taps_x = (mode >> 4) & 3, taps_y = (mode >> 6) & 3;
        void* tab = nullptr;
        const int rc_pool = retain_async_pool();
        if (rc_pool) return rc_pool;
        B200WM_CUDA_TRY(cudaMallocAsync(&tab, entry * (size_t)(dp->width + dp->height), stream));
        void* tab_y = (uint8_t*)tab + entry * (size_t)dp->width;
        resize_tables_kernel<<<(dp->width + dp->height + 255) / 256, 256, 0, stream>>>(area, tab, tab_y, sp->width, sp->height, dp->width,
                                                                                      dp->height, scale_x, scale_y);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) {
            count_launch();
            if (area == 2) {
                const AreaTaps* ax = (const AreaTaps*)tab;
                const AreaTaps* ay = (const AreaTaps*)tab_y;
#define B200WM_AREA_TAPS(TX, TY)                                                                                                   \
    if (taps_x == TX && taps_y == TY)                                                                                              \
        resize_area_taps_kernel<TX, TY><<<grid4, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes,         \
                                                                   dp->frame_stride_bytes, (unsigned)dp->pitch_bytes, dp->width,   \
                                                                   dp->height, ax, ay);
                B200WM_AREA_TAPS(1, 1) B200WM_AREA_TAPS(1, 2) B200WM_AREA_TAPS(1, 3) B200WM_AREA_TAPS(2, 1) B200WM_AREA_TAPS(2, 2)
                B200WM_AREA_TAPS(2, 3) B200WM_AREA_TAPS(3, 1) B200WM_AREA_TAPS(3, 2) B200WM_AREA_TAPS(3, 3)
#undef B200WM_AREA_TAPS
            } else if (area)
                resize_area_kernel<<<grid4, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes, dp->frame_stride_bytes,
                                                              (unsigned)dp->pitch_bytes, dp->width, dp->height, (const AreaSpan*)tab, (const AreaSpan*)tab_y);
            else
                resize_linear_kernel<<<grid4, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes, dp->frame_stride_bytes,
                                                                (unsigned)dp->pitch_bytes, dp->width, dp->height, (const LinearTap*)tab, (const LinearTap*)tab_y);
            e = cudaGetLastError();
        }
        cudaFreeAsync(tab, stream);
        if (e != cudaSuccess) {
            set_cuda_error(e, area ? "resize_area_kernel" : "resize_linear_kernel");
            return B200WM_ERR_CUDA;
        }
        count_launch();
        return B200WM_OK;
    };
    if (interpolation == B200WM_INTER_AREA) {
        // OpenCV treats INTER_AREA as a (modified) bilinear filter when enlarging: not provided here
        if (dp->width > sp->width || dp->height > sp->height) return B200WM_ERR_UNSUPPORTED;
        const int ix = sp->width / dp->width, iy = sp->height / dp->height;
        if (ix * dp->width == sp->width && iy * dp->height == sp->height) {
            resize_area_int_kernel<<<grid, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes,
                                                             dp->frame_stride_bytes, (unsigned)dp->pitch_bytes, dp->width, ix, iy);
            B200WM_LAUNCH_CHECK("resize_area_int_kernel");
            return B200WM_OK;
        }
        // exact tap counts per axis (the same double arithmetic as the table kernel, once per column and row on the host)
        int tx = 0, ty = 0;
        for (int c = 0; c < dp->width; ++c) { const int k = area_tap_count(c, scale_x, sp->width); tx = k > tx ? k : tx; }
        for (int r = 0; r < dp->height; ++r) { const int k = area_tap_count(r, scale_y, sp->height); ty = k > ty ? k : ty; }
        if (tx >= 1 && tx <= 3 && ty >= 1 && ty <= 3) return with_tables(2 + 16 * tx + 64 * ty, sizeof(AreaTaps));
        return with_tables(1, sizeof(AreaSpan));
    }
    // cv2 silently turns an exact 2x2 bilinear reduction into INTER_AREA
    if (sp->width == 2 * dp->width && sp->height == 2 * dp->height) {
        resize_area_int_kernel<<<grid, 256, 0, stream>>>(s, d, sp->frame_stride_bytes, (unsigned)sp->pitch_bytes,
                                                         dp->frame_stride_bytes, (unsigned)dp->pitch_bytes, dp->width, 2, 2);
        B200WM_LAUNCH_CHECK("resize_area_int_kernel");
        return B200WM_OK;
    }
    return with_tables(0, sizeof(LinearTap));
}

}  // namespace b200wm
