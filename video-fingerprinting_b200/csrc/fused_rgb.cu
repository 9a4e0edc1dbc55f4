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

// byte i of a row as the float 2^23 + byte: PRMT drops it into the mantissa of 2^23 (exact); the 2^23 is removed by a
// packed FADD2 on two pixels at once.  An I2F per byte would put 384 conversions per tile on the quarter-rate
// conversion pipe.
__device__ __forceinline__ float biased_of(const unsigned (&w)[6], int i) {
    return __uint_as_float(__byte_perm(w[i >> 2], 0x4B000000u, 0x7650u | (unsigned)(i & 3)));
}
// Two pixels per instruction (FFMA2 / FADD2 / FMUL2): lane .x and lane .y are different pixels, every operation is the
// scalar formula's operation in the scalar formula's order, so the values are OpenCV's float path bit for bit.
struct Yuv2 { f2 y, u, v; };
__device__ __forceinline__ Yuv2 px_to_yuv2(f2 b0, f2 b1, f2 b2) {        // b*: 2^23 + byte
    const f2 unbias = bc2(-8388608.0f);
    const f2 c0 = add2(b0, unbias), c1 = add2(b1, unbias), c2 = add2(b2, unbias);
    Yuv2 o;
    o.y = fma2(c0, bc2(0.114f), fma2(c1, bc2(0.587f), mul2(c2, bc2(0.299f))));
    o.u = fma2(fma2(o.y, bc2(-1.0f), c0), bc2(0.492f), bc2(0.5f));      // (c0 - y) * 0.492 + 0.5; y * -1 + c0 == c0 - y exactly
    o.v = fma2(fma2(o.y, bc2(-1.0f), c2), bc2(0.877f), bc2(0.5f));
    return o;
}

// 24 colour values of a row -> 24 bytes: clip(., 0, 255) then round-half-even (video/embedder.py:37-38) equals
// round-half-even then clamp; adding 1.5 * 2^23 rounds and leaves the integer in the low mantissa bits, which
// are packed as int16 lanes, clamped by one DPX instruction per two values and narrowed to bytes.
__device__ __forceinline__ void pack_row24(const float (&c)[24], unsigned (&out)[6]) {
    const f2 bias = bc2(12582912.0f);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const f2 lo = add2(make_float2(c[4 * k], c[4 * k + 1]), bias), hi = add2(make_float2(c[4 * k + 2], c[4 * k + 3]), bias);
        const unsigned e = __viaddmin_s16x2_relu(__byte_perm(__float_as_uint(lo.x), __float_as_uint(hi.x), 0x5410), 0u, 0x00FF00FFu);
        const unsigned o = __viaddmin_s16x2_relu(__byte_perm(__float_as_uint(lo.y), __float_as_uint(hi.y), 0x5410), 0u, 0x00FF00FFu);
        out[k] = __byte_perm(e, o, 0x6240);
    }
}

// The four 2x2 sums of one LL row (pixel rows 2i and 2i+1 of the tile) for the channels in kMask.  Packed lanes are
// (row 2i, row 2i+1) of the same pixel column, so a 2x2 is (left + right) in both lanes, then lane .x + lane .y:
// (top-left + top-right) + (bottom-left + bottom-right) like the scalar loader of dwtsvd_tile.cuh.
template <bool kAligned, int kMask>
__device__ __forceinline__ void ll_row_sums(const uint8_t* p, unsigned pitch, float (&s)[3][4]) {
    unsigned r0[6], r1[6];
    load_row24<kAligned>(p, r0);
    load_row24<kAligned>(p + pitch, r1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        Yuv2 q[2];
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int px = 2 * j + dx;
            q[dx] = px_to_yuv2(make_float2(biased_of(r0, 3 * px), biased_of(r1, 3 * px)),
                               make_float2(biased_of(r0, 3 * px + 1), biased_of(r1, 3 * px + 1)),
                               make_float2(biased_of(r0, 3 * px + 2), biased_of(r1, 3 * px + 2)));
        }
        if (kMask & 1) { const f2 t = add2(q[0].y, q[1].y); s[0][j] = t.x + t.y; }
        if (kMask & 2) { const f2 t = add2(q[0].u, q[1].u); s[1][j] = t.x + t.y; }
        if (kMask & 4) { const f2 t = add2(q[0].v, q[1].v); s[2][j] = t.x + t.y; }
    }
}

// G += s s^T for one row of the block, packed upper-triangular: row 0 first, then rows 1..3 - the accumulation order
// of gram4() (fma(a, b, 0) rounds like a * b), so the eigen-solver sees the Gram matrix it would get from the whole block.
__device__ __forceinline__ void gram_add_row(const float (&s)[4], float (&G)[10]) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j, ++k) G[k] = fmaf(s[i], s[j], G[k]);
}

// What embed_deltas() of dwtsvd_tile.cuh derives from a block, without the block: the increment of the 2x2 (i, j)
// is  flat ? flat_inc : (t * (S_row_i . v)) * v[j].
struct EmbedFactors {
    float v[4], t, flat_inc;
    bool flat;              // all-zero block: svd(0) = (I, 0, I), the same increment everywhere
};

template <typename Probe>
__device__ __forceinline__ void embed_factors(float (&G)[10], int bit, float scale, float inv_scale, EmbedFactors& f, Probe probe) {
    bool maybe_flat;
    float sigma = 0.5f * top_singular_gram<true, 3>(G, f.v, f.flat, maybe_flat);
    float q, rem;
    floor_divmod(sigma, scale, inv_scale, q, rem);
    if (maybe_flat && on_boundary<false>(sigma, rem, scale)) {
        const float flat = probe();
        if (flat == flat) {
            sigma = flat_sigma_ref(flat);
            floor_divmod(sigma, scale, inv_scale, q, rem);
        }
    }
    const float target = (q + 0.25f + 0.5f * (float)bit) * scale;
    f.flat_inc = target * 0.125f;
    f.t = 0.25f * ((target - sigma) / sigma);
}

// kMask: bit c set = YUV channel c is marked (scale[c] > 0); compile-time so that unmarked channels
// cost neither registers nor instructions (the reference default, scales=[0,15,0], is kMask = 2).
//
// Two rolled passes over the four LL rows of the tile (code that fits the instruction cache: the fully unrolled
// predecessor was 107 KB of SASS and stalled on instruction fetch more than on anything else, profiles/r02_fused_rgb.md):
//   pass A  convert two pixel rows, form the row's four 2x2 sums per marked channel, park them in shared memory and
//           add their outer product to the Gram matrix;
//   solve   sigma_0, v_0 and the quantisation target per marked channel from the Gram matrix alone;
//   pass B  convert the two pixel rows again (L1 hits), add the row's increments in YUV space, convert back, clip,
//           round, store.
constexpr int kRgbThreads = 128;

template <bool kAligned, int kMask>
__global__ void __launch_bounds__(kRgbThreads, 4) dwtsvd_embed_rgb8_kernel(RgbArgs a, EmbedArgs em, TileGeom g, int frame0) {
    constexpr int kCh = (kMask & 1) + ((kMask >> 1) & 1) + ((kMask >> 2) & 1);
    __shared__ float4 s_rows[kCh][4][kRgbThreads];
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * kRgbThreads + threadIdx.x;
    if (c >= (unsigned)g.n_tiles) return;
    const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
    const unsigned tx = c - ty * g.tiles_x;
    const int row = em.frame_row ? clamp_row(em.frame_row[frame], em.n_rows) : 0;
    const int bit = (em.wm[(long long)row * em.wm_words + (c >> 5)] >> (c & 31)) & 1;
    const long long off = frame * a.frame_stride + (unsigned long long)(ty * 8) * a.pitch + tx * 24;

    float G[3][10];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int k = 0; k < 10; ++k) G[ch][k] = 0.0f;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        float s[3][4];
        ll_row_sums<kAligned, kMask>(a.src + off + (unsigned long long)(2 * i) * a.pitch, a.pitch, s);
        int slot = 0;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            if (kMask & (1 << ch)) {
                gram_add_row(s[ch], G[ch]);
                s_rows[slot++][i][threadIdx.x] = make_float4(s[ch][0], s[ch][1], s[ch][2], s[ch][3]);
            }
    }
    EmbedFactors f[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        if (kMask & (1 << ch))
            embed_factors(G[ch], bit, a.scale[ch], 1.0f / a.scale[ch], f[ch], [&]() { return flat_probe_rgb(a.src + off, a.pitch, ch); });

#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        float d[3][4];                       // increments of the four 2x2 of this LL row
        int slot = 0;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            if (kMask & (1 << ch)) {
                const float4 sr = s_rows[slot++][i][threadIdx.x];
                const float sv = fmaf(sr.w, f[ch].v[3], fmaf(sr.z, f[ch].v[2], fmaf(sr.y, f[ch].v[1], sr.x * f[ch].v[0])));
                const float ai = f[ch].t * sv;
#pragma unroll
                for (int j = 0; j < 4; ++j) d[ch][j] = f[ch].flat ? f[ch].flat_inc : ai * f[ch].v[j];
            }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const unsigned long long ro = (unsigned long long)(2 * i + rr) * a.pitch;
            unsigned raw[6];
            load_row24<kAligned>(a.src + off + ro, raw);
            float cv[24];
#pragma unroll
            for (int j = 0; j < 4; ++j) {         // lanes = pixels 2j and 2j+1 of this row: they share the increment of their 2x2
                const int p0 = 2 * j, p1 = 2 * j + 1;
                Yuv2 q = px_to_yuv2(make_float2(biased_of(raw, 3 * p0), biased_of(raw, 3 * p1)),
                                    make_float2(biased_of(raw, 3 * p0 + 1), biased_of(raw, 3 * p1 + 1)),
                                    make_float2(biased_of(raw, 3 * p0 + 2), biased_of(raw, 3 * p1 + 2)));
                if (kMask & 1) q.y = add2(q.y, bc2(d[0][j]));
                if (kMask & 2) q.u = add2(q.u, bc2(d[1][j]));
                if (kMask & 4) q.v = add2(q.v, bc2(d[2][j]));
                const f2 du = add2(q.u, bc2(-0.5f)), dv = add2(q.v, bc2(-0.5f));
                const f2 o0 = fma2(du, bc2(2.032f), q.y);
                const f2 o1 = fma2(dv, bc2(-0.581f), fma2(du, bc2(-0.395f), q.y));
                const f2 o2 = fma2(dv, bc2(1.14f), q.y);
                cv[3 * p0] = o0.x; cv[3 * p0 + 1] = o1.x; cv[3 * p0 + 2] = o2.x;
                cv[3 * p1] = o0.y; cv[3 * p1 + 1] = o1.y; cv[3 * p1 + 2] = o2.y;
            }
            unsigned out[6];
            pack_row24(cv, out);
            uint8_t* o = a.dst + off + ro;
            if (kAligned) {
                uint2* q = reinterpret_cast<uint2*>(o);
                q[0] = make_uint2(out[0], out[1]); q[1] = make_uint2(out[2], out[3]); q[2] = make_uint2(out[4], out[5]);
            } else {
#pragma unroll
                for (int k = 0; k < 24; ++k) o[k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
}

template <bool kAligned, int kChannel>      // kChannel: the YUV channel that is read (compile time: the two others cost nothing)
__global__ void __launch_bounds__(kRgbThreads) dwtsvd_extract_rgb8_kernel(RgbArgs a, ExtractArgs ex, TileGeom g, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * kRgbThreads + threadIdx.x;
    const unsigned word = c >> 5;
    const bool live = word < (unsigned)g.words;
    int bit = 0;
    if (c < (unsigned)g.n_tiles) {
        const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
        const unsigned tx = c - ty * g.tiles_x;
        const long long off = frame * a.frame_stride + (unsigned long long)(ty * 8) * a.pitch + tx * 24;
        float G[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) G[k] = 0.0f;
#pragma unroll      // 4 x 330 instructions: small enough for the instruction cache, and the loads of all rows are in flight together
        for (int i = 0; i < 4; ++i) {
            float s[3][4];
            ll_row_sums<kAligned, 1 << kChannel>(a.src + off + (unsigned long long)(2 * i) * a.pitch, a.pitch, s);
            gram_add_row(s[kChannel], G);
        }
        float v[4];
        bool zero, maybe_flat;
        float sigma = 0.5f * top_singular_gram<false, 3>(G, v, zero, maybe_flat), q, rem;
        floor_divmod(sigma, ex.scale, ex.inv_scale, q, rem);
        if (maybe_flat && on_boundary<true>(sigma, rem, ex.scale)) {
            const float flat = flat_probe_rgb(a.src + off, a.pitch, kChannel);
            if (flat == flat) {
                sigma = flat_sigma_ref(flat);
                floor_divmod(sigma, ex.scale, ex.inv_scale, q, rem);
            }
        }
        bit = rem > 0.5f * ex.scale ? 1 : 0;
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
        if (aligned) dwtsvd_embed_rgb8_kernel<true, M><<<grid, kRgbThreads, 0, stream>>>(a, ea, g, f0);  \
        else dwtsvd_embed_rgb8_kernel<false, M><<<grid, kRgbThreads, 0, stream>>>(a, ea, g, f0);         \
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
#define B200WM_RGB_X(CH)                                                                              \
    case CH:                                                                                         \
        if (aligned) dwtsvd_extract_rgb8_kernel<true, CH><<<grid, 128, 0, stream>>>(a, xa, g, f0);    \
        else dwtsvd_extract_rgb8_kernel<false, CH><<<grid, 128, 0, stream>>>(a, xa, g, f0);           \
        break;
            switch (channel) { B200WM_RGB_X(0) B200WM_RGB_X(1) B200WM_RGB_X(2) }
#undef B200WM_RGB_X
            B200WM_LAUNCH_CHECK("dwtsvd_extract_rgb8_kernel");
        }
    }
    if (pos_counts && !fused)
        return launch_vote_counts(raw_bits, n_frames, words_per_frame, g.block_num, payload_len, pos_counts, stream);
    return B200WM_OK;
}

}  // namespace b200wm
