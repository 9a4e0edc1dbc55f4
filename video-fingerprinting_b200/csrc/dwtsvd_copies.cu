// One read, N marked copies: the "N watermarked copies of each segment" step of the reference's
// fingerprinting flow (tests/mark_video_to_hls.py:330-354 runs the whole embedder once per copy over
// the same source segment, with payload = 4-bit segment || 4-bit copy index, :27-43).
//
// The source tile, its 2x2 sums, sigma_0 and the singular pair do not depend on the payload, so each
// thread loads and decomposes its 8x8 tile ONCE and then writes one marked tile per copy: only the
// quantisation target and the rank-1 increment are per copy (embed_copy_deltas, the same arithmetic
// as the single embed, so copy c equals b200wm_dwtsvd_embed with payload row c bit for bit).
// Algorithmic bytes per frame: (1 + N) * W * H instead of 2 * N * W * H, and one eigen-solve instead
// of N.  Stores are 64-bit, coalesced into 256-byte row segments per warp like the single embed.
#include "common.cuh"
#include "svd4.cuh"
#include "dwtsvd_tile.cuh"

namespace b200wm {

struct CopyArgs {
    uint8_t* dst;               // copy c, frame f at dst + c * copy_stride + f * frame_stride
    long long copy_stride;
    const int32_t* copy_row;    // nullable [n_frames, n_copies]: watermark row of (frame, copy); NULL -> row = copy
    int n_copies;
};

template <int kMode>   // 0: planar uint8, 8-byte aligned rows; 1: any uint8 layout
__global__ void __launch_bounds__(kThreads) dwtsvd_embed_copies_kernel(PlaneArgs pl, CopyArgs cp, EmbedArgs em, TileGeom g,
                                                                      int frame0) {
    const int frame = frame0 + blockIdx.y;
    const unsigned c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= (unsigned)g.n_tiles) return;
    const unsigned ty = (unsigned)(((unsigned long long)c * g.div_magic) >> 40);
    const unsigned tx = c - ty * g.tiles_x;
    const long long off = frame * pl.frame_stride + (unsigned long long)(ty * 8) * pl.pitch;
    const int32_t* rows_of_frame = cp.copy_row ? cp.copy_row + (long long)frame * cp.n_copies : nullptr;
    const uint32_t* wm_word = em.wm + (c >> 5);
    float D[16];
    BlockPair bp;
    if (kMode == 0) {
        const long long o0 = off + tx * 8;
        uint2 rows[8];
        {
            float S[16];
            load_tile_u8<true>(pl.src + o0, pl.pitch, rows, S);
            embed_prepare(S, em.scale, em.inv_scale, bp, [&]() { return flat_probe_global<uint8_t>(pl.src + o0, pl.pitch, 1); });
        }
#pragma unroll 1
        for (int k = 0; k < cp.n_copies; ++k) {
            const int row = clamp_row(rows_of_frame ? rows_of_frame[k] : k, em.n_rows);
            const int bit = (wm_word[(long long)row * em.wm_words] >> (c & 31)) & 1;
            embed_copy_deltas(bp, bit, em.scale, 12582912.0f, D);
            uint8_t* o = cp.dst + k * cp.copy_stride + o0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned d01 = __byte_perm(__float_as_uint(D[4 * i + 0]), __float_as_uint(D[4 * i + 1]), 0x5410);
                const unsigned d23 = __byte_perm(__float_as_uint(D[4 * i + 2]), __float_as_uint(D[4 * i + 3]), 0x5410);
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
                    stg_stream_u2(row_ptr(o, 2 * i + rr, pl.pitch), add_clamp_row(rows[2 * i + rr], d01, d23));
            }
        }
    } else {
        const int es = pl.elem_stride;
        const uint8_t* p = pl.src + off + (long long)tx * 8 * es;
        {
            float S[16];
            load_tile_generic<uint8_t>(p, pl.pitch, es, S);
            embed_prepare(S, em.scale, em.inv_scale, bp, [&]() { return flat_probe_global<uint8_t>(p, pl.pitch, es); });
        }
#pragma unroll 1
        for (int k = 0; k < cp.n_copies; ++k) {
            const int row = clamp_row(rows_of_frame ? rows_of_frame[k] : k, em.n_rows);
            const int bit = (wm_word[(long long)row * em.wm_words] >> (c & 31)) & 1;
            embed_copy_deltas(bp, bit, em.scale, 0.0f, D);
            uint8_t* o = cp.dst + k * cp.copy_stride + off + (long long)tx * 8 * es;
#pragma unroll
            for (int y = 0; y < 8; ++y)
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const float f = (float)row_ptr(p, y, pl.pitch)[x * es] + rintf(D[4 * (y >> 1) + (x >> 1)]);
                    row_ptr(o, y, pl.pitch)[x * es] = (uint8_t)__float2int_rn(fminf(fmaxf(f, 0.0f), 255.0f));
                }
        }
    }
}

int validate_plane(const b200wm_plane* pl);

int launch_dwtsvd_embed_copies(const void* src, const b200wm_plane* pl, void* dst, long long copy_stride, int n_copies,
                               const uint32_t* wm, int n_wm_rows, int wm_words, long long wm_len, const int32_t* copy_row,
                               float scale, cudaStream_t stream) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (!src || !dst || !wm || wm_words <= 0 || n_copies < 0 || n_wm_rows <= 0 || !(scale > 0.0f)) return B200WM_ERR_INVALID;
    if (pl->dtype != B200WM_U8) return B200WM_ERR_UNSUPPORTED;
    if (!copy_row && n_copies > n_wm_rows) return B200WM_ERR_INVALID;
    const TileGeom g = make_geom(pl->height, pl->width);
    if (wm_len < g.n_tiles || (long long)wm_words * 32 < g.n_tiles) return B200WM_ERR_SHORT_WM;
    if (g.n_tiles == 0 || pl->n_frames == 0 || n_copies == 0) return B200WM_OK;
    const bool aligned = pl->elem_stride == 1 && (pl->pitch_bytes % 8) == 0 && (pl->frame_stride_bytes % 8) == 0 &&
                         (copy_stride % 8) == 0 && ((uintptr_t)src % 8) == 0 && ((uintptr_t)dst % 8) == 0;
    PlaneArgs pa{(const uint8_t*)src, nullptr, pl->frame_stride_bytes, (unsigned)pl->pitch_bytes, pl->elem_stride};
    CopyArgs ca{(uint8_t*)dst, copy_stride, copy_row, n_copies};
    EmbedArgs ea{wm, nullptr, n_wm_rows, wm_words, scale, 1.0f / scale};
    const unsigned gx = (g.n_tiles + kThreads - 1) / kThreads;
    for (int f0 = 0; f0 < pl->n_frames; f0 += 65535) {
        const dim3 grid(gx, (unsigned)((pl->n_frames - f0) < 65535 ? (pl->n_frames - f0) : 65535));
        if (aligned) dwtsvd_embed_copies_kernel<0><<<grid, kThreads, 0, stream>>>(pa, ca, ea, g, f0);
        else dwtsvd_embed_copies_kernel<1><<<grid, kThreads, 0, stream>>>(pa, ca, ea, g, f0);
        B200WM_LAUNCH_CHECK("dwtsvd_embed_copies_kernel");
    }
    return B200WM_OK;
}

}  // namespace b200wm
