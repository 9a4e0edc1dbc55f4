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
#include <atomic>
#include "common.cuh"
#include "svd4.cuh"
#include "dwtsvd_tile.cuh"

namespace b200wm {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// grid = (ceil(words / warps_per_cta), frames in this launch); one warp = one word of raw bits.
// mode 0 is capped at 64 registers (8 CTAs of 128 threads per SM): measured 2.29 ms per 3000 1080p frames against
// 2.46 ms at 80 and 2.63 ms uncapped; the generic modes need their registers
template <int kMode>   // 0: u8 fast, 1: u8 generic, 2: f32 generic
__global__ void __launch_bounds__(kThreads, kMode == 0 ? 8 : 1) dwtsvd_embed_kernel(PlaneArgs pl, EmbedArgs em, TileGeom g, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= (unsigned)g.n_tiles) return;
    const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
    const unsigned tx = c - ty * g.tiles_x;
    const int row = em.frame_row ? clamp_row(em.frame_row[frame], em.n_rows) : 0;
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
        embed_deltas<true>(S, bit, em.scale, em.inv_scale, 12582912.0f, D, s_stash + threadIdx.x,
                           [&]() { return flat_probe_global<uint8_t>(p, pl.pitch, 1); });
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // 16-bit lanes: lo = increment of the left 2x2 of this word, hi = the right one
            const unsigned d01 = __byte_perm(__float_as_uint(D[4 * i + 0]), __float_as_uint(D[4 * i + 1]), 0x5410);
            const unsigned d23 = __byte_perm(__float_as_uint(D[4 * i + 2]), __float_as_uint(D[4 * i + 3]), 0x5410);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint2 out = add_clamp_row(rows[2 * i + rr], d01, d23);
                stg_stream_u2(row_ptr(o, 2 * i + rr, pl.pitch), out);
            }
        }
    } else if (kMode == 1) {
        const int es = pl.elem_stride;
        const uint8_t* p = pl.src + off + (long long)tx * 8 * es;
        uint8_t* o = pl.dst + off + (long long)tx * 8 * es;
        load_tile_generic<uint8_t>(p, pl.pitch, es, S);
        embed_deltas<false>(S, bit, em.scale, em.inv_scale, 0.0f, D, nullptr, [&]() { return flat_probe_global<uint8_t>(p, pl.pitch, es); });
#pragma unroll
        for (int y = 0; y < 8; ++y)
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                // sample + round-half-even(increment), like the vector path (see add_clamp_row)
                const float f = (float)row_ptr(p, y, pl.pitch)[x * es] + rintf(D[4 * (y >> 1) + (x >> 1)]);
                row_ptr(o, y, pl.pitch)[x * es] = (uint8_t)__float2int_rn(fminf(fmaxf(f, 0.0f), 255.0f));
            }
    } else {
        const int es = pl.elem_stride;
        const uint8_t* p = pl.src + off + (long long)tx * 8 * es * 4;
        uint8_t* o = pl.dst + off + (long long)tx * 8 * es * 4;
        load_tile_generic<float>(p, pl.pitch, es, S);
        embed_deltas<false, 3>(S, bit, em.scale, em.inv_scale, 0.0f, D, nullptr, [&]() { return flat_probe_global<float>(p, pl.pitch, es); });
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
        bit = extract_bit<kMode == 2 ? 3 : 0>(S, ex.scale, ex.inv_scale, sigma, [&]() {
            if (kMode == 2) return flat_probe_global<float>(pl.src + off + (long long)tx * 8 * pl.elem_stride * 4, pl.pitch, pl.elem_stride);
            return flat_probe_global<uint8_t>(pl.src + off + (long long)tx * 8 * (kMode == 0 ? 1 : pl.elem_stride), pl.pitch,
                                              kMode == 0 ? 1 : pl.elem_stride);
        });
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
            const int n = __popc(ballot & (ex.every << lane));
            if (n) atomicAdd(&cta_counts[lane], n);
        }
        __syncthreads();
        if ((int)threadIdx.x < L && cta_counts[threadIdx.x])
            atomicAdd(&ex.pos_counts[(long long)frame * L + threadIdx.x], cta_counts[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------
// packed variant of the fast extract path (planar uint8, 8-byte aligned rows): two tiles per thread in
// FFMA2 / FMUL2 / FADD2 (svd4x2.cuh), bit-identical to the scalar kernel above.  A warp owns 64
// consecutive tiles, lane l the tiles base + l and base + 32 + l, so its two ballots are two
// consecutive words of the packed raw bits.  Used for every plane the TMA kernels do not take
// (720p and portrait 1080p rows, rows that are not 16-byte aligned, small planes): 1.05 ms per 3000
// 1080p frames against 1.22 ms for the scalar kernel.  (The same treatment of the embed kernel needed 168
// registers with the source rows re-read from L2 and ran at half the scalar kernel's speed: not kept.)
// ------------------------------------------------------------------------------------------
struct TileAddr {
    long long off;     // byte offset of the tile's first sample inside the plane buffer
    bool live;
};
__device__ __forceinline__ TileAddr tile_addr(unsigned c, int frame, const PlaneArgs& pl, const TileGeom& g) {
    TileAddr a;
    a.live = c < (unsigned)g.n_tiles;
    const unsigned cc = a.live ? c : 0u;                 // dead lanes read tile 0 of the frame and are masked later
    const unsigned ty = (unsigned)(((unsigned long long)cc * g.div_magic) >> 40);
    const unsigned tx = cc - ty * g.tiles_x;
    a.off = frame * pl.frame_stride + (unsigned long long)(ty * 8) * pl.pitch + tx * 8;
    return a;
}

#ifndef B200WM_LDG_EXTRACT_MIN_CTAS
#define B200WM_LDG_EXTRACT_MIN_CTAS 6      // 80 registers: 1.03 ms per 3000 1080p frames (1.06 uncapped, 1.07 at 64)
#endif
__global__ void __launch_bounds__(kThreads, B200WM_LDG_EXTRACT_MIN_CTAS) dwtsvd_extract_x2_kernel(PlaneArgs pl, ExtractArgs ex, TileGeom g, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned base = (blockIdx.x * (kThreads / 32) + warp) * 64;
    const TileAddr lo = tile_addr(base + lane, frame, pl, g), hi = tile_addr(base + 32 + lane, frame, pl, g);
    unsigned bits;
    {
        f2 S[16];
        {
            uint2 ra[8], rb[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                ra[r] = ldg_nc_u2(row_ptr(pl.src + lo.off, r, pl.pitch));
                rb[r] = ldg_nc_u2(row_ptr(pl.src + hi.off, r, pl.pitch));
            }
            sums_from_rows_x2(ra, rb, S);
        }
        bits = extract_bits_x2(S, ex.scale, ex.inv_scale,
                               [&](int which) { return flat_probe_global<uint8_t>(pl.src + (which ? hi.off : lo.off), pl.pitch, 1); });
    }
    const unsigned ballot_lo = __ballot_sync(0xFFFFFFFFu, lo.live && (bits & 1u));
    const unsigned ballot_hi = __ballot_sync(0xFFFFFFFFu, hi.live && (bits & 2u));
    const unsigned word = base >> 5;
    if (lane == 0) {
        uint32_t* w = ex.raw_bits + (long long)frame * g.words;
        if (word < (unsigned)g.words) w[word] = ballot_lo;
        if (word + 1 < (unsigned)g.words) w[word + 1] = ballot_hi;
    }
    if (ex.pos_counts && (int)lane < ex.payload_len) {
        // payload_len divides 32 and both halves start at a multiple of 32: one reduction per warp and position
        const int n = __popc(ballot_lo & (ex.every << lane)) + __popc(ballot_hi & (ex.every << lane));
        if (n) atomicAdd(&ex.pos_counts[(long long)frame * ex.payload_len + lane], n);
    }
}

// Validation kernel: sigma_0 computed the way the reference writes it, WITH the 4x4 DCT
// (svd(cv2.dct(block)), extract/dwt_dct_svd_decoder.py:35): orthonormal 4-point DCT-II butterflies on the
// rows and columns of the LL block, then the same top-singular routine.  The production kernels drop
// the DCT because it is orthogonal (sigma is invariant); tests compare the two on the GPU.
__device__ __forceinline__ void dct4_1d(float& x0, float& x1, float& x2, float& x3) {
    const float s0 = x0 + x3, s1 = x1 + x2, d0 = x0 - x3, d1 = x1 - x2;
    x0 = (s0 + s1) * 0.5f;
    x2 = (s0 - s1) * 0.5f;
    x1 = fmaf(d0, 0.65328148243818826f, d1 * 0.27059805007309849f);       // cos(pi/8)/sqrt(2), cos(3pi/8)/sqrt(2)
    x3 = fmaf(d0, 0.27059805007309849f, d1 * -0.65328148243818826f);
}

template <int kMode>
__global__ void __launch_bounds__(kThreads) dwtsvd_sigma_dct_kernel(PlaneArgs pl, TileGeom g, float* __restrict__ sigma, int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= (unsigned)g.n_tiles) return;
    const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
    const unsigned tx = c - ty * g.tiles_x;
    const long long off = frame * pl.frame_stride + (unsigned long long)(ty * 8) * pl.pitch;
    float B[16];
    if (kMode == 2) load_tile_generic<float>(pl.src + off + (long long)tx * 8 * pl.elem_stride * 4, pl.pitch, pl.elem_stride, B);
    else load_tile_generic<uint8_t>(pl.src + off + (long long)tx * 8 * pl.elem_stride, pl.pitch, pl.elem_stride, B);
#pragma unroll
    for (int k = 0; k < 16; ++k) B[k] *= 0.5f;                              // Haar LL = 2x2 sum / 2
#pragma unroll
    for (int i = 0; i < 4; ++i) dct4_1d(B[4 * i], B[4 * i + 1], B[4 * i + 2], B[4 * i + 3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) dct4_1d(B[j], B[4 + j], B[8 + j], B[12 + j]);
    float v[4];
    bool zero, maybe_flat;
    sigma[(long long)frame * g.n_tiles + c] = top_singular<false>(B, v, zero, maybe_flat);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
bool tma_eligible(const void* a, const void* b, const b200wm_plane* pl, const TileGeom& g);
int launch_dwtsvd_extract_tma(const void* src, const b200wm_plane* pl, const TileGeom& g, ExtractArgs xa, cudaStream_t stream);
int launch_dwtsvd_embed_tma(const void* src, void* dst, const b200wm_plane* pl, const TileGeom& g, EmbedArgs ea,
                            cudaStream_t stream);

// Path selection: 0 = automatic (TMA when the plane qualifies), 1 = force the LDG kernels.
// Set through b200wm_set_path() for A/B measurements; results are identical either way.
static std::atomic<int> g_path{0};      // a measurement switch, read once per launch; results do not depend on it
void set_path(int p) { g_path.store(p, std::memory_order_relaxed); }
int get_path() { return g_path.load(std::memory_order_relaxed); }

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

int launch_dwtsvd_embed(const void* src, void* dst, const b200wm_plane* pl, const uint32_t* wm, int n_wm_rows, int wm_words,
                        long long wm_len, const int32_t* frame_row, float scale, cudaStream_t stream) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (!src || !dst || !wm || n_wm_rows <= 0 || wm_words <= 0 || !(scale > 0.0f)) return B200WM_ERR_INVALID;
    const TileGeom g = make_geom(pl->height, pl->width);
    if (wm_len < g.n_tiles || (long long)wm_words * 32 < g.n_tiles) return B200WM_ERR_SHORT_WM;
    if (g.n_tiles == 0 || pl->n_frames == 0) return B200WM_OK;
    if ((uintptr_t)src % 4 && pl->dtype == B200WM_F32) return B200WM_ERR_INVALID;
    PlaneArgs pa{(const uint8_t*)src, (uint8_t*)dst, pl->frame_stride_bytes, (unsigned)pl->pitch_bytes, pl->elem_stride};
    EmbedArgs ea{wm, frame_row, n_wm_rows, wm_words, scale, 1.0f / scale};
    if (get_path() == 0 && tma_eligible(src, dst, pl, g)) return launch_dwtsvd_embed_tma(src, dst, pl, g, ea, stream);
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

int launch_dwtsvd_sigma_dct(const void* src, const b200wm_plane* pl, float* sigma, cudaStream_t stream) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (!src || !sigma) return B200WM_ERR_INVALID;
    const TileGeom g = make_geom(pl->height, pl->width);
    if (g.n_tiles == 0 || pl->n_frames == 0) return B200WM_OK;
    if ((uintptr_t)src % 4 && pl->dtype == B200WM_F32) return B200WM_ERR_INVALID;
    PlaneArgs pa{(const uint8_t*)src, nullptr, pl->frame_stride_bytes, (unsigned)pl->pitch_bytes, pl->elem_stride};
    const unsigned gx = (g.n_tiles + kThreads - 1) / kThreads;
    for (int f0 = 0; f0 < pl->n_frames; f0 += 65535) {
        const dim3 grid(gx, (unsigned)((pl->n_frames - f0) < 65535 ? (pl->n_frames - f0) : 65535));
        if (pl->dtype == B200WM_F32) dwtsvd_sigma_dct_kernel<2><<<grid, kThreads, 0, stream>>>(pa, g, sigma, f0);
        else dwtsvd_sigma_dct_kernel<1><<<grid, kThreads, 0, stream>>>(pa, g, sigma, f0);
        B200WM_LAUNCH_CHECK("dwtsvd_sigma_dct_kernel");
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
    if (g.words > 0 && g.n_tiles == 0) {
        // a plane with a dimension below 8 (e.g. 4 x 64) has raw-bit words but no tile: nothing to read
        B200WM_CUDA_TRY(cudaMemsetAsync(raw_bits, 0, sizeof(uint32_t) * (size_t)pl->n_frames * g.words, stream));
        if (pos_counts && !fused)
            B200WM_CUDA_TRY(cudaMemsetAsync(pos_counts, 0, sizeof(int32_t) * (size_t)pl->n_frames * payload_len, stream));
    } else if (g.words > 0) {
        PlaneArgs pa{(const uint8_t*)src, nullptr, pl->frame_stride_bytes, (unsigned)pl->pitch_bytes, pl->elem_stride};
        ExtractArgs xa{raw_bits, fused ? pos_counts : nullptr, sigma, payload_len, fused ? every_mask(payload_len) : 0u, scale, 1.0f / scale};
        if (get_path() == 0 && !sigma && tma_eligible(src, src, pl, g)) {
            // strips OR their bits into the packed words; surplus words must read 0
            B200WM_CUDA_TRY(cudaMemsetAsync(raw_bits, 0, sizeof(uint32_t) * (size_t)pl->n_frames * g.words, stream));
            rc = launch_dwtsvd_extract_tma(src, pl, g, xa, stream);
            if (rc) return rc;
            if (pos_counts && !fused)
                return launch_vote_counts(raw_bits, pl->n_frames, words_per_frame, g.block_num, payload_len, pos_counts, stream);
            return B200WM_OK;
        }
        const int mode = plane_mode(src, src, pl);
        const unsigned gx = ((unsigned)g.words * 32 + kThreads - 1) / kThreads;
        for (int f0 = 0; f0 < pl->n_frames; f0 += 65535) {
            const dim3 grid(gx, (unsigned)((pl->n_frames - f0) < 65535 ? (pl->n_frames - f0) : 65535));
            if (mode == 0 && !sigma) {
                const dim3 grid2(((unsigned)g.words * 32 + 2 * kThreads - 1) / (2 * kThreads), grid.y);
                dwtsvd_extract_x2_kernel<<<grid2, kThreads, 0, stream>>>(pa, xa, g, f0);
            } else if (mode == 0) dwtsvd_extract_kernel<0><<<grid, kThreads, 0, stream>>>(pa, xa, g, f0);
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
