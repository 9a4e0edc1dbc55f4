// Colour bracket fused into the DWT/SVD embed and extract.
//
// The reference's video drivers wrap every frame plugin call in
//   uint8 HxWx3 -> float32 -> cv2.cvtColor(BGR2YUV) -> encode/decode -> cv2.cvtColor(YUV2BGR) -> clip -> around -> uint8
// (src/offmark/video/embedder.py:33-39, src/offmark/video/extractor.py:30-34).  Run as separate
// kernels that is 3 + 12 + (strided) 24 + 12 + 3 bytes of traffic per pixel; fused it is 3 bytes in
// and 3 bytes out: the thread that owns an 8x8 tile converts its 64 pixels on the fly, builds the
// 2x2 sums of the marked channel(s), runs the same SVD quantiser as dwtsvd.cu, adds the increments
// in YUV space and converts back.  Same OpenCV float formulas (fused multiply-adds) as bracket.cu.
#include "common.cuh"
#include "svd4.cuh"
#include "dwtsvd_tile.cuh"

namespace b200wm {

struct RgbArgs {
    const uint8_t* src;
    uint8_t* dst;
    long long frame_stride;
    unsigned pitch;            // bytes per row (3*width when tight)
    float scale[3];            // per YUV channel; <= 0: channel untouched (reference: scales=[0,15,0])
};

__device__ __forceinline__ void px_to_yuv(float c0, float c1, float c2, float& y, float& u, float& v) {
    y = fmaf(c0, 0.114f, fmaf(c1, 0.587f, c2 * 0.299f));
    u = fmaf(c0 - y, 0.492f, 0.5f);
    v = fmaf(c2 - y, 0.877f, 0.5f);
}

// Flat-tile probe (dwtsvd_tile.cuh) for packed rgb frames: all 64 pixels equal -> the value of YUV channel
// `channel` of that colour, NaN otherwise.
static __device__ __noinline__ float flat_probe_rgb(const uint8_t* p, unsigned pitch, int channel) {
    const uint8_t c0 = p[0], c1 = p[1], c2 = p[2];
    bool flat = true;
#pragma unroll 1
    for (int y = 0; y < 8; ++y) {
        const uint8_t* r = p + (unsigned long long)y * pitch;
#pragma unroll 1
        for (int x = 0; x < 8; ++x) flat &= (r[3 * x] == c0) & (r[3 * x + 1] == c1) & (r[3 * x + 2] == c2);
    }
    float y, u, v;
    px_to_yuv((float)c0, (float)c1, (float)c2, y, u, v);
    return flat ? (channel == 0 ? y : (channel == 1 ? u : v)) : __int_as_float(0x7FC00000);
}

// bytes of one tile row: 8 pixels x 3 channels = 24 bytes = 6 words
template <bool kAligned>
__device__ __forceinline__ void load_row24(const uint8_t* p, unsigned (&w)[6]) {
    if (kAligned) {
        const uint2* q = reinterpret_cast<const uint2*>(p);
        const uint2 a = q[0], b = q[1], c = q[2];
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y;
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            w[k] = (unsigned)p[4 * k] | ((unsigned)p[4 * k + 1] << 8) | ((unsigned)p[4 * k + 2] << 16) | ((unsigned)p[4 * k + 3] << 24);
    }
}

__device__ __forceinline__ float byte_of(const unsigned (&w)[6], int i) {
    return (float)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);       // compiles to one I2F with a byte selector
}

// 24 colour values of a row -> 24 bytes: clip(., 0, 255) then round-half-even (video/embedder.py:37-38) equals
// round-half-even then clamp; adding 1.5 * 2^23 rounds and leaves the integer in the low mantissa bits, which
// are packed as int16 lanes, clamped by one DPX instruction per two values and narrowed to bytes.
__device__ __forceinline__ void pack_row24(const float (&c)[24], unsigned (&out)[6]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const unsigned b0 = __float_as_uint(c[4 * k] + 12582912.0f), b1 = __float_as_uint(c[4 * k + 1] + 12582912.0f);
        const unsigned b2 = __float_as_uint(c[4 * k + 2] + 12582912.0f), b3 = __float_as_uint(c[4 * k + 3] + 12582912.0f);
        const unsigned e = __viaddmin_s16x2_relu(__byte_perm(b0, b2, 0x5410), 0u, 0x00FF00FFu);
        const unsigned o = __viaddmin_s16x2_relu(__byte_perm(b1, b3, 0x5410), 0u, 0x00FF00FFu);
        out[k] = __byte_perm(e, o, 0x6240);
    }
}

// kMask: bit c set = YUV channel c is marked (scale[c] > 0); compile-time so that unmarked channels
// cost neither registers nor instructions (the reference default, scales=[0,15,0], is kMask = 2).
template <bool kAligned, int kMask>
__global__ void __launch_bounds__(128, 4) dwtsvd_embed_rgb8_kernel(RgbArgs a, EmbedArgs em, TileGeom g, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * 128 + threadIdx.x;
    if (c >= (unsigned)g.n_tiles) return;
    const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
    const unsigned tx = c - ty * g.tiles_x;
    const int row = em.frame_row ? clamp_row(em.frame_row[frame], em.n_rows) : 0;
    const int bit = (em.wm[(long long)row * em.wm_words + (c >> 5)] >> (c & 31)) & 1;
    const long long off = frame * a.frame_stride + (unsigned long long)(ty * 8) * a.pitch + tx * 24;

    // Pass A: 2x2 sums of the marked channels.  The 192 source bytes are NOT kept in registers across
    // the eigen-iteration; pass B re-reads them row by row (L1 hits).
    float S[3][16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned r0[6], r1[6];
        load_row24<kAligned>(a.src + off + (unsigned long long)(2 * i) * a.pitch, r0);
        load_row24<kAligned>(a.src + off + (unsigned long long)(2 * i + 1) * a.pitch, r1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float y[4], u[4], v[4];
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int px = 2 * j + dx;
                px_to_yuv(byte_of(r0, 3 * px), byte_of(r0, 3 * px + 1), byte_of(r0, 3 * px + 2), y[dx], u[dx], v[dx]);
                px_to_yuv(byte_of(r1, 3 * px), byte_of(r1, 3 * px + 1), byte_of(r1, 3 * px + 2), y[2 + dx], u[2 + dx], v[2 + dx]);
            }
            if (kMask & 1) S[0][4 * i + j] = (y[0] + y[1]) + (y[2] + y[3]);
            if (kMask & 2) S[1][4 * i + j] = (u[0] + u[1]) + (u[2] + u[3]);
            if (kMask & 4) S[2][4 * i + j] = (v[0] + v[1]) + (v[2] + v[3]);
        }
    }
    float D[3][16];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        if (kMask & (1 << ch))
            embed_deltas<false>(S[ch], bit, a.scale[ch], 1.0f / a.scale[ch], 0.0f, D[ch], nullptr,
                                [&]() { return flat_probe_rgb(a.src + off, a.pitch, ch); });

    // Pass B: convert again, add the increments in YUV space, convert back, clip, round, store.
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        unsigned raw[6];
        load_row24<kAligned>(a.src + off + (unsigned long long)r * a.pitch, raw);
        float cv[24];
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            float y, u, v;
            px_to_yuv(byte_of(raw, 3 * px), byte_of(raw, 3 * px + 1), byte_of(raw, 3 * px + 2), y, u, v);
            const int k = 4 * (r >> 1) + (px >> 1);
            if (kMask & 1) y += D[0][k];
            if (kMask & 2) u += D[1][k];
            if (kMask & 4) v += D[2][k];
            const float du = u - 0.5f, dv = v - 0.5f;
            cv[3 * px] = fmaf(du, 2.032f, y);
            cv[3 * px + 1] = fmaf(dv, -0.581f, fmaf(du, -0.395f, y));
            cv[3 * px + 2] = fmaf(dv, 1.14f, y);
        }
        unsigned out[6];
        pack_row24(cv, out);
        uint8_t* o = a.dst + off + (unsigned long long)r * a.pitch;
        if (kAligned) {
            uint2* q = reinterpret_cast<uint2*>(o);
            q[0] = make_uint2(out[0], out[1]); q[1] = make_uint2(out[2], out[3]); q[2] = make_uint2(out[4], out[5]);
        } else {
#pragma unroll
            for (int k = 0; k < 24; ++k) o[k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
        }
    }
}

template <bool kAligned>
__global__ void __launch_bounds__(128) dwtsvd_extract_rgb8_kernel(RgbArgs a, int channel, ExtractArgs ex, TileGeom g, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * 128 + threadIdx.x;
    const unsigned word = c >> 5;
    const bool live = word < (unsigned)g.words;
    int bit = 0;
    if (c < (unsigned)g.n_tiles) {
        const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
        const unsigned tx = c - ty * g.tiles_x;
        const long long off = frame * a.frame_stride + (unsigned long long)(ty * 8) * a.pitch + tx * 24;
        float S[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned r0[6], r1[6];
            load_row24<kAligned>(a.src + off + (unsigned long long)(2 * i) * a.pitch, r0);
            load_row24<kAligned>(a.src + off + (unsigned long long)(2 * i + 1) * a.pitch, r1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float q[4];
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int px = 2 * j + dx;
                    float y, u, v;
                    px_to_yuv(byte_of(r0, 3 * px), byte_of(r0, 3 * px + 1), byte_of(r0, 3 * px + 2), y, u, v);
                    q[dx] = channel == 0 ? y : (channel == 1 ? u : v);
                    px_to_yuv(byte_of(r1, 3 * px), byte_of(r1, 3 * px + 1), byte_of(r1, 3 * px + 2), y, u, v);
                    q[2 + dx] = channel == 0 ? y : (channel == 1 ? u : v);
                }
                S[4 * i + j] = (q[0] + q[1]) + (q[2] + q[3]);
            }
        }
        float sigma;
        bit = extract_bit(S, ex.scale, ex.inv_scale, sigma, [&]() { return flat_probe_rgb(a.src + off, a.pitch, channel); });
    }
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, bit);
    const unsigned lane = threadIdx.x & 31;
    if (lane == 0 && live) ex.raw_bits[(long long)frame * g.words + word] = ballot;
    if (ex.pos_counts && (int)lane < ex.payload_len) {
        // straight to global memory, one reduction per warp and payload position: a CTA-wide barrier here made
        // every warp wait for the slowest eigen-solve of the CTA (8.6 stall cycles per issue in ncu)
        const int n = __popc(ballot & (ex.every << lane));
        if (n) atomicAdd(&ex.pos_counts[(long long)frame * ex.payload_len + lane], n);
    }
}

int launch_vote_counts(const uint32_t* raw_bits, int n_frames, int words_per_frame, long long block_num,
                       int payload_len, int32_t* pos_counts, cudaStream_t stream);

static int check_rgb(const void* src, int n_frames, int height, int width, long long pitch, long long frame_stride) {
    if (!src || n_frames < 0 || height <= 0 || width <= 0) return B200WM_ERR_INVALID;
    if (pitch < 3ll * width || pitch >= (1ll << 31)) return B200WM_ERR_INVALID;
    if (n_frames > 1 && frame_stride < pitch * height) return B200WM_ERR_INVALID;
    if ((long long)height * width / 64 >= (1ll << 26)) return B200WM_ERR_UNSUPPORTED;
    return B200WM_OK;
}

int launch_embed_rgb8(const uint8_t* src, uint8_t* dst, int n_frames, int height, int width, long long pitch,
                      long long frame_stride, const float* scales, const uint32_t* wm, int n_wm_rows, int wm_words, long long wm_len,
                      const int32_t* frame_row, cudaStream_t stream) {
    int rc = check_rgb(src, n_frames, height, width, pitch, frame_stride);
    if (rc) return rc;
    if (!dst || !scales || !wm || n_wm_rows <= 0 || wm_words <= 0) return B200WM_ERR_INVALID;
    const TileGeom g = make_geom(height, width);
    const int mask = (scales[0] > 0.0f ? 1 : 0) | (scales[1] > 0.0f ? 2 : 0) | (scales[2] > 0.0f ? 4 : 0);
    if (mask == 0) return B200WM_OK;             // nothing to mark: the reference's loop skips every channel
    if (wm_len < g.n_tiles || (long long)wm_words * 32 < g.n_tiles) return B200WM_ERR_SHORT_WM;
    if (g.n_tiles == 0 || n_frames == 0) return B200WM_OK;
    RgbArgs a{src, dst, frame_stride, (unsigned)pitch, {scales[0], scales[1], scales[2]}};
    EmbedArgs ea{wm, frame_row, n_wm_rows, wm_words, 0.0f, 0.0f};
    const bool aligned = ((uintptr_t)src % 8) == 0 && ((uintptr_t)dst % 8) == 0 && pitch % 8 == 0 && frame_stride % 8 == 0;
    const unsigned gx = (g.n_tiles + 127) / 128;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const dim3 grid(gx, (unsigned)((n_frames - f0) < 65535 ? (n_frames - f0) : 65535));
#define B200WM_RGB_CASE(M)                                                                          \
    case M:                                                                                         \
        if (aligned) dwtsvd_embed_rgb8_kernel<true, M><<<grid, 128, 0, stream>>>(a, ea, g, f0);      \
        else dwtsvd_embed_rgb8_kernel<false, M><<<grid, 128, 0, stream>>>(a, ea, g, f0);             \
        break;
        switch (mask) {
            B200WM_RGB_CASE(1) B200WM_RGB_CASE(2) B200WM_RGB_CASE(3) B200WM_RGB_CASE(4)
            B200WM_RGB_CASE(5) B200WM_RGB_CASE(6) B200WM_RGB_CASE(7)
        }
#undef B200WM_RGB_CASE
        B200WM_LAUNCH_CHECK("dwtsvd_embed_rgb8_kernel");
    }
    return B200WM_OK;
}

int launch_extract_rgb8(const uint8_t* src, int n_frames, int height, int width, long long pitch, long long frame_stride,
                        int channel, float scale, uint32_t* raw_bits, int words_per_frame, int payload_len,
                        int32_t* pos_counts, cudaStream_t stream) {
    int rc = check_rgb(src, n_frames, height, width, pitch, frame_stride);
    if (rc) return rc;
    if (!raw_bits || channel < 0 || channel > 2 || !(scale > 0.0f)) return B200WM_ERR_INVALID;
    if (pos_counts && payload_len <= 0) return B200WM_ERR_INVALID;
    const TileGeom g = make_geom(height, width);
    if (words_per_frame != g.words) return B200WM_ERR_INVALID;
    if (n_frames == 0) return B200WM_OK;
    const bool fused = pos_counts && payload_len <= 32 && (32 % payload_len) == 0;
    if (fused) B200WM_CUDA_TRY(cudaMemsetAsync(pos_counts, 0, sizeof(int32_t) * (size_t)n_frames * payload_len, stream));
    if (g.words > 0) {
        RgbArgs a{src, nullptr, frame_stride, (unsigned)pitch, {0.0f, 0.0f, 0.0f}};
        ExtractArgs xa{raw_bits, fused ? pos_counts : nullptr, nullptr, payload_len, fused ? every_mask(payload_len) : 0u, scale,
                       1.0f / scale};
        const bool aligned = ((uintptr_t)src % 8) == 0 && pitch % 8 == 0 && frame_stride % 8 == 0;
        const unsigned gx = ((unsigned)g.words * 32 + 127) / 128;
        for (int f0 = 0; f0 < n_frames; f0 += 65535) {
            const dim3 grid(gx, (unsigned)((n_frames - f0) < 65535 ? (n_frames - f0) : 65535));
            if (aligned) dwtsvd_extract_rgb8_kernel<true><<<grid, 128, 0, stream>>>(a, channel, xa, g, f0);
            else dwtsvd_extract_rgb8_kernel<false><<<grid, 128, 0, stream>>>(a, channel, xa, g, f0);
            B200WM_LAUNCH_CHECK("dwtsvd_extract_rgb8_kernel");
        }
    }
    if (pos_counts && !fused)
        return launch_vote_counts(raw_bits, n_frames, words_per_frame, g.block_num, payload_len, pos_counts, stream);
    return B200WM_OK;
}

}  // namespace b200wm
