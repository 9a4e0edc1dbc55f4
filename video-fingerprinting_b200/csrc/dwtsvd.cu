// Haar-DWT / 4x4-block SVD quantisation-index embed and extract kernels (sm_100a).
//
// Replaces the per-block Python loops of the reference:
//   embed   src/offmark/embed/dwt_dct_svd_encoder.py:19-45
//   extract src/offmark/extract/dwt_dct_svd_decoder.py:12-37
//   counts  src/offmark/degenerator/de_shuffler.py:17-18 (the sums behind the means)
//
// One thread owns one 8x8-sample tile = one 4x4 block of the Haar LL band.
//   LL = (a+b+c+d)/2 of each 2x2 (pywt.dwt2 'haar'); the detail bands are never modified and
//   idwt2 is linear, so writing the new LL back is  sample += (LL' - LL)/2  on each 2x2.
//   The kernels work on S = 2*LL (the plain 2x2 sums, exact for uint8): sigma(LL) = sigma(S)/2.
//   sigma_0' = (floor(sigma_0/scale) + 0.25 + 0.5*bit)*scale, and since only sigma_0 changes,
//   S' = S + (sigma_0'/sigma_0 - 1) (S v0) v0^T,   sample += (S' - S)/4.
// A warp owns 32 consecutive block indices, so a ballot is one word of the packed raw bits and
// (when payload_len divides 32) the per-position vote counts come from popc on masked ballots.
//
// Memory: every sample of the plane is read exactly once (extract) or read once and written
// once (embed) with 64-bit accesses that a warp coalesces into full 256-byte row segments;
// algorithmic bytes per frame are W*H (extract) and 2*W*H (embed), see DESIGN.md.
#include "common.cuh"
#include "svd4.cuh"

namespace b200wm {

constexpr int kThreads = 128;

struct PlaneArgs {
    const uint8_t* src;
    uint8_t* dst;
    long long frame_stride;   // bytes
    unsigned pitch;           // bytes (< 2^31: row addresses are one IMAD.WIDE each)
    int elem_stride;          // samples
};

__device__ __forceinline__ const uint8_t* row_ptr(const uint8_t* p, unsigned r, unsigned pitch) {
    return p + (unsigned long long)r * pitch;
}
__device__ __forceinline__ uint8_t* row_ptr(uint8_t* p, unsigned r, unsigned pitch) {
    return p + (unsigned long long)r * pitch;
}

struct EmbedArgs {
    const uint32_t* wm;       // [rows, wm_words]
    const int32_t* frame_row; // nullable
    int wm_words;
    float scale, inv_scale;
};

struct ExtractArgs {
    uint32_t* raw_bits;       // [n_frames, words]
    int32_t* pos_counts;      // nullable, [n_frames, payload_len]; only when 32 % payload_len == 0
    float* sigma;             // nullable debug output [n_frames, n_tiles]
    int payload_len;
    float scale, inv_scale;
};

// ------------------------------------------------------------------------------------------
// tile loaders: produce S[16] (2x2 sums, row-major over the 4x4 block)
// ------------------------------------------------------------------------------------------
// Fast path: planar uint8, 8-byte aligned rows.  rows[r] keeps the raw bytes for the embed.
template <bool kReadOnly>
__device__ __forceinline__ void load_tile_u8(const uint8_t* p, unsigned pitch, uint2 (&rows)[8], float (&S)[16]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) rows[r] = kReadOnly ? ldg_nc_u2(row_ptr(p, r, pitch)) : ldg_stream_u2(row_ptr(p, r, pitch));
    // float(2^23 + n) has n in its low mantissa bits: accumulate the four bytes of a 2x2 with
    // dp4a straight into that bit pattern, then one FADD removes the 2^23.
    constexpr unsigned kMagic = 0x4B000000u;
    constexpr float kMagicF = 8388608.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint2 a = rows[2 * i], b = rows[2 * i + 1];
        unsigned s0 = __dp4a(a.x, 0x00000101u, kMagic); s0 = __dp4a(b.x, 0x00000101u, s0);
        unsigned s1 = __dp4a(a.x, 0x01010000u, kMagic); s1 = __dp4a(b.x, 0x01010000u, s1);
        unsigned s2 = __dp4a(a.y, 0x00000101u, kMagic); s2 = __dp4a(b.y, 0x00000101u, s2);
        unsigned s3 = __dp4a(a.y, 0x01010000u, kMagic); s3 = __dp4a(b.y, 0x01010000u, s3);
        S[4 * i + 0] = __uint_as_float(s0) - kMagicF;
        S[4 * i + 1] = __uint_as_float(s1) - kMagicF;
        S[4 * i + 2] = __uint_as_float(s2) - kMagicF;
        S[4 * i + 3] = __uint_as_float(s3) - kMagicF;
    }
}

// Generic path: any dtype / stride / alignment.
template <typename T>
__device__ __forceinline__ void load_tile_generic(const uint8_t* p, unsigned pitch, int es, float (&S)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T* r0 = reinterpret_cast<const T*>(row_ptr(p, 2 * i, pitch));
        const T* r1 = reinterpret_cast<const T*>(row_ptr(p, 2 * i + 1, pitch));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = (float)r0[(2 * j) * es], b = (float)r0[(2 * j + 1) * es];
            const float c = (float)r1[(2 * j) * es], d = (float)r1[(2 * j + 1) * es];
            S[4 * i + j] = (a + b) + (c + d);
        }
    }
}

// ------------------------------------------------------------------------------------------
// per-block quantisation
// ------------------------------------------------------------------------------------------
// Per-sample increment of each 2x2 of the tile for watermark bit `bit`: D[4*i+j] = (S'-S)[i][j]/4.
// kStash: park S in shared memory while the eigen-iteration runs (4 STS.128 + 4 LDS.128 per
// thread) instead of letting the compiler rebuild it from the pixel bytes under register pressure.
template <bool kStash>
__device__ __forceinline__ void embed_deltas(float (&S)[16], int bit, float scale, float inv_scale,
                                             float bias, float (&D)[16], float4* stash) {
    float v[4];
    bool zero;
    if (kStash) {
#pragma unroll
        for (int i = 0; i < 4; ++i)   // asm: the compiler must not forward these stores to the loads below
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(stash + i * kThreads)),
                         "f"(S[4 * i]), "f"(S[4 * i + 1]), "f"(S[4 * i + 2]), "f"(S[4 * i + 3]) : "memory");
    }
    const float sigma = 0.5f * top_singular<true>(S, v, zero);   // sigma_0 of the LL block
    float q, rem;
    floor_divmod(sigma, scale, inv_scale, q, rem);
    const float target = (q + 0.25f + 0.5f * (float)bit) * scale;
    if (zero) {
        // svd(0) = (I, 0, I): the reference puts sigma_0' on DCT coefficient [0][0], i.e. a flat
        // block target/4 in the LL band -> target/8 on every sample.
#pragma unroll
        for (int k = 0; k < 16; ++k) D[k] = fmaf(target, 0.125f, bias);
        return;
    }
    const float t = 0.25f * ((target - sigma) / sigma);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s0 = S[4 * i], s1 = S[4 * i + 1], s2 = S[4 * i + 2], s3 = S[4 * i + 3];
        if (kStash) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(s0), "=f"(s1), "=f"(s2), "=f"(s3)
                         : "r"((unsigned)__cvta_generic_to_shared(stash + i * kThreads)) : "memory");
        }
        const float sv = fmaf(s3, v[3], fmaf(s2, v[2], fmaf(s1, v[1], s0 * v[0])));
        const float a = t * sv;
#pragma unroll
        for (int j = 0; j < 4; ++j) D[4 * i + j] = fmaf(a, v[j], bias);
    }
}

__device__ __forceinline__ int extract_bit(const float (&S)[16], float scale, float inv_scale, float& sigma) {
    float v[4];
    bool zero;
    sigma = 0.5f * top_singular<false>(S, v, zero);
    float q, rem;
    floor_divmod(sigma, scale, inv_scale, q, rem);
    return rem > 0.5f * scale ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// grid = (ceil(words / warps_per_cta), frames in this launch); one warp = one word of raw bits.
template <int kMode>   // 0: u8 fast, 1: u8 generic, 2: f32 generic
__global__ void __launch_bounds__(kThreads) dwtsvd_embed_kernel(PlaneArgs pl, EmbedArgs em, TileGeom g, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= (unsigned)g.n_tiles) return;
    const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
    const unsigned tx = c - ty * g.tiles_x;
    const int row = em.frame_row ? em.frame_row[frame] : 0;
    const int bit = (em.wm[(long long)row * em.wm_words + (c >> 5)] >> (c & 31)) & 1;

    const long long off = frame * pl.frame_stride + (unsigned long long)(ty * 8) * pl.pitch;
    float S[16], D[16];
    if (kMode == 0) {
        const uint8_t* p = pl.src + off + tx * 8;
        uint8_t* o = pl.dst + off + tx * 8;
        uint2 rows[8];
        __shared__ float4 s_stash[4 * kThreads];
        load_tile_u8<false>(p, pl.pitch, rows, S);
        // bias 1.5*2^23: the FMA rounds the increment to the nearest integer (ties to even) and
        // leaves it, two's complement, in the low mantissa bits.
        embed_deltas<true>(S, bit, em.scale, em.inv_scale, 12582912.0f, D, s_stash + threadIdx.x);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // 16-bit lanes: lo = increment of the left 2x2 of this word, hi = the right one
            const unsigned d01 = __byte_perm(__float_as_uint(D[4 * i + 0]), __float_as_uint(D[4 * i + 1]), 0x5410);
            const unsigned d23 = __byte_perm(__float_as_uint(D[4 * i + 2]), __float_as_uint(D[4 * i + 3]), 0x5410);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint2 w = rows[2 * i + rr];
                uint2 out;
                {
                    const unsigned e = __byte_perm(w.x, 0u, 0x4240), od = __byte_perm(w.x, 0u, 0x4341);
                    const unsigned e2 = __viaddmin_s16x2_relu(e, d01, 0x00FF00FFu);
                    const unsigned o2 = __viaddmin_s16x2_relu(od, d01, 0x00FF00FFu);
                    out.x = __byte_perm(e2, o2, 0x6240);
                }
                {
                    const unsigned e = __byte_perm(w.y, 0u, 0x4240), od = __byte_perm(w.y, 0u, 0x4341);
                    const unsigned e2 = __viaddmin_s16x2_relu(e, d23, 0x00FF00FFu);
                    const unsigned o2 = __viaddmin_s16x2_relu(od, d23, 0x00FF00FFu);
                    out.y = __byte_perm(e2, o2, 0x6240);
                }
                stg_stream_u2(row_ptr(o, 2 * i + rr, pl.pitch), out);
            }
        }
    } else if (kMode == 1) {
        const int es = pl.elem_stride;
        const uint8_t* p = pl.src + off + (long long)tx * 8 * es;
        uint8_t* o = pl.dst + off + (long long)tx * 8 * es;
        load_tile_generic<uint8_t>(p, pl.pitch, es, S);
        embed_deltas<false>(S, bit, em.scale, em.inv_scale, 0.0f, D, nullptr);
#pragma unroll
        for (int y = 0; y < 8; ++y)
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const float f = (float)row_ptr(p, y, pl.pitch)[x * es] + D[4 * (y >> 1) + (x >> 1)];
                row_ptr(o, y, pl.pitch)[x * es] = (uint8_t)__float2int_rn(fminf(fmaxf(f, 0.0f), 255.0f));
            }
    } else {
        const int es = pl.elem_stride;
        const uint8_t* p = pl.src + off + (long long)tx * 8 * es * 4;
        uint8_t* o = pl.dst + off + (long long)tx * 8 * es * 4;
        load_tile_generic<float>(p, pl.pitch, es, S);
        embed_deltas<false>(S, bit, em.scale, em.inv_scale, 0.0f, D, nullptr);
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const float* pr = reinterpret_cast<const float*>(row_ptr(p, y, pl.pitch));
            float* orow = reinterpret_cast<float*>(row_ptr(o, y, pl.pitch));
#pragma unroll
            for (int x = 0; x < 8; ++x) orow[x * es] = pr[x * es] + D[4 * (y >> 1) + (x >> 1)];
        }
    }
}

template <int kMode>
__global__ void __launch_bounds__(kThreads) dwtsvd_extract_kernel(PlaneArgs pl, ExtractArgs ex, TileGeom g, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * kThreads + threadIdx.x;
    const unsigned word = c >> 5;
    const bool live = word < (unsigned)g.words;     // warp-uniform; no early return (barrier below)
    int bit = 0;
    if (c < (unsigned)g.n_tiles) {
        const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
        const unsigned tx = c - ty * g.tiles_x;
        const long long off = frame * pl.frame_stride + (unsigned long long)(ty * 8) * pl.pitch;
        float S[16];
        if (kMode == 0) {
            uint2 rows[8];
            load_tile_u8<true>(pl.src + off + tx * 8, pl.pitch, rows, S);
        } else if (kMode == 1) {
            load_tile_generic<uint8_t>(pl.src + off + (long long)tx * 8 * pl.elem_stride, pl.pitch, pl.elem_stride, S);
        } else {
            load_tile_generic<float>(pl.src + off + (long long)tx * 8 * pl.elem_stride * 4, pl.pitch, pl.elem_stride, S);
        }
        float sigma;
        bit = extract_bit(S, ex.scale, ex.inv_scale, sigma);
        if (ex.sigma) ex.sigma[(long long)frame * g.n_tiles + c] = sigma;
    }
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, bit);
    const unsigned lane = threadIdx.x & 31;
    if (lane == 0 && live) ex.raw_bits[(long long)frame * g.words + word] = ballot;

    if (ex.pos_counts) {
        // payload_len divides 32 and every warp starts at a multiple of 32, so position i of the
        // payload owns ballot bits i, i+L, i+2L, ...
        __shared__ int cta_counts[32];
        const int L = ex.payload_len;
        if (threadIdx.x < 32) cta_counts[threadIdx.x] = 0;
        __syncthreads();
        if ((int)lane < L) {
            const unsigned every = L == 32 ? 1u : (0xFFFFFFFFu / ((1u << L) - 1u));
            const int n = __popc(ballot & (every << lane));
            if (n) atomicAdd(&cta_counts[lane], n);
        }
        __syncthreads();
        if ((int)threadIdx.x < L && cta_counts[threadIdx.x])
            atomicAdd(&ex.pos_counts[(long long)frame * L + threadIdx.x], cta_counts[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int plane_mode(const void* a, const void* b, const b200wm_plane* pl) {
    if (pl->dtype == B200WM_F32) return 2;
    const bool aligned = pl->elem_stride == 1 && (pl->pitch_bytes % 8) == 0 && (pl->frame_stride_bytes % 8) == 0 &&
                         ((uintptr_t)a % 8) == 0 && ((uintptr_t)b % 8) == 0;
    return aligned ? 0 : 1;
}

int validate_plane(const b200wm_plane* pl) {
    if (!pl) return B200WM_ERR_INVALID;
    if (pl->dtype != B200WM_U8 && pl->dtype != B200WM_F32) return B200WM_ERR_INVALID;
    if (pl->n_frames < 0 || pl->height <= 0 || pl->width <= 0 || pl->elem_stride <= 0) return B200WM_ERR_INVALID;
    const long long esz = pl->dtype == B200WM_U8 ? 1 : 4;
    if (pl->pitch_bytes < (long long)pl->width * pl->elem_stride * esz - (pl->elem_stride - 1) * esz) return B200WM_ERR_INVALID;
    if (pl->dtype == B200WM_F32 && (pl->pitch_bytes % 4 || pl->frame_stride_bytes % 4)) return B200WM_ERR_INVALID;
    if ((long long)pl->height * pl->width / 64 >= (1ll << 26) || pl->pitch_bytes >= (1ll << 31)) return B200WM_ERR_UNSUPPORTED;
    return B200WM_OK;
}

int launch_dwtsvd_embed(const void* src, void* dst, const b200wm_plane* pl, const uint32_t* wm, int wm_words,
                        long long wm_len, const int32_t* frame_row, float scale, cudaStream_t stream) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (!src || !dst || !wm || wm_words <= 0 || !(scale > 0.0f)) return B200WM_ERR_INVALID;
    const TileGeom g = make_geom(pl->height, pl->width);
    if (wm_len < g.n_tiles || (long long)wm_words * 32 < g.n_tiles) return B200WM_ERR_SHORT_WM;
    if (g.n_tiles == 0 || pl->n_frames == 0) return B200WM_OK;
    if ((uintptr_t)src % 4 && pl->dtype == B200WM_F32) return B200WM_ERR_INVALID;
    PlaneArgs pa{(const uint8_t*)src, (uint8_t*)dst, pl->frame_stride_bytes, (unsigned)pl->pitch_bytes, pl->elem_stride};
    EmbedArgs ea{wm, frame_row, wm_words, scale, 1.0f / scale};
    const int mode = plane_mode(src, dst, pl);
    const unsigned gx = (g.n_tiles + kThreads - 1) / kThreads;
    for (int f0 = 0; f0 < pl->n_frames; f0 += 65535) {
        const dim3 grid(gx, (unsigned)((pl->n_frames - f0) < 65535 ? (pl->n_frames - f0) : 65535));
        if (mode == 0) dwtsvd_embed_kernel<0><<<grid, kThreads, 0, stream>>>(pa, ea, g, f0);
        else if (mode == 1) dwtsvd_embed_kernel<1><<<grid, kThreads, 0, stream>>>(pa, ea, g, f0);
        else dwtsvd_embed_kernel<2><<<grid, kThreads, 0, stream>>>(pa, ea, g, f0);
        B200WM_LAUNCH_CHECK("dwtsvd_embed_kernel");
    }
    return B200WM_OK;
}

int launch_vote_counts(const uint32_t* raw_bits, int n_frames, int words_per_frame, long long block_num,
                       int payload_len, int32_t* pos_counts, cudaStream_t stream);

int launch_dwtsvd_extract(const void* src, const b200wm_plane* pl, float scale, uint32_t* raw_bits,
                          int words_per_frame, int payload_len, int32_t* pos_counts, float* sigma,
                          cudaStream_t stream) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (!src || !raw_bits || !(scale > 0.0f)) return B200WM_ERR_INVALID;
    if (pos_counts && payload_len <= 0) return B200WM_ERR_INVALID;
    const TileGeom g = make_geom(pl->height, pl->width);
    if (words_per_frame != g.words) return B200WM_ERR_INVALID;
    if (pl->n_frames == 0) return B200WM_OK;
    if ((uintptr_t)src % 4 && pl->dtype == B200WM_F32) return B200WM_ERR_INVALID;
    const bool fused = pos_counts && payload_len <= 32 && (32 % payload_len) == 0;
    if (pos_counts && fused)
        B200WM_CUDA_TRY(cudaMemsetAsync(pos_counts, 0, sizeof(int32_t) * (size_t)pl->n_frames * payload_len, stream));
    if (g.words > 0) {
        PlaneArgs pa{(const uint8_t*)src, nullptr, pl->frame_stride_bytes, (unsigned)pl->pitch_bytes, pl->elem_stride};
        ExtractArgs xa{raw_bits, fused ? pos_counts : nullptr, sigma, payload_len, scale, 1.0f / scale};
        const int mode = plane_mode(src, src, pl);
        const unsigned gx = ((unsigned)g.words * 32 + kThreads - 1) / kThreads;
        for (int f0 = 0; f0 < pl->n_frames; f0 += 65535) {
            const dim3 grid(gx, (unsigned)((pl->n_frames - f0) < 65535 ? (pl->n_frames - f0) : 65535));
            if (mode == 0) dwtsvd_extract_kernel<0><<<grid, kThreads, 0, stream>>>(pa, xa, g, f0);
            else if (mode == 1) dwtsvd_extract_kernel<1><<<grid, kThreads, 0, stream>>>(pa, xa, g, f0);
            else dwtsvd_extract_kernel<2><<<grid, kThreads, 0, stream>>>(pa, xa, g, f0);
            B200WM_LAUNCH_CHECK("dwtsvd_extract_kernel");
        }
    }
    if (pos_counts && !fused)
        return launch_vote_counts(raw_bits, pl->n_frames, words_per_frame, g.block_num, payload_len, pos_counts, stream);
    return B200WM_OK;
}

}  // namespace b200wm
