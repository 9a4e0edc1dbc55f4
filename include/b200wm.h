/*
 * b200wm.h - C ABI of libb200wm.so: the B200 (sm_100a) watermark hot path.
 *
 * This is the drop-in boundary for the per-frame embed / extract / vote path of
 * offmark-py (vikasdimaniya/video-fingerprinting).  The reference is pure Python
 * and has no FFI of its own; the entry points below are what a ctypes binding
 * placed behind the reference's frame plugins would call (INTEGRATION.md shows
 * that stub).  Each function names the reference code it replaces, cited as
 * file:line relative to the reference repository root.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - Every function returns 0 (B200WM_OK) or a negative b200wm_status; nothing
 *    throws.  b200wm_strerror() names a status; b200wm_last_cuda_error() gives
 *    the CUDA text behind B200WM_ERR_CUDA for the calling thread.
 *  - Unless a function says "host", every data pointer is a DEVICE pointer on the
 *    current CUDA device and the work is enqueued on `stream` (a cudaStream_t
 *    passed as void*; NULL = legacy default stream) without synchronising.
 *  - The library never allocates caller-visible memory.
 *  - There is no CPU fallback: with no usable device the calls return
 *    B200WM_ERR_CUDA / B200WM_ERR_NO_DEVICE.
 *
 * Plane geometry ("b200wm_plane")
 *  One watermark plane per frame: `height` x `width` samples of `dtype`, sample
 *  (y, x) of frame f at byte address
 *      base + f*frame_stride_bytes + y*pitch_bytes + x*elem_stride*sizeof(dtype).
 *  elem_stride = 1 for a planar plane (the Y plane of I420 frames),
 *  elem_stride = 3 with base pointing at channel c for an interleaved H x W x 3
 *  frame (the reference's float32 YUV frames, c = 1).
 *  The fast path (128-bit/64-bit vector loads) needs dtype U8, elem_stride 1 and
 *  base, pitch and frame stride multiples of 8 bytes; anything else takes the
 *  generic path with identical results.
 *
 * Block order
 *  As the reference walks the Haar LL band (embed/dwt_dct_svd_encoder.py:31-40):
 *  the plane is cut to multiples of 4, each 8x8-sample tile is one 4x4 LL block,
 *  tiles are numbered row-major, tiles_x = ((width/4*4)/2)/4, tiles_y likewise.
 *  block_num = height*width/64 (extract/dwt_dct_svd_decoder.py:14) can exceed
 *  tiles_x*tiles_y; the surplus raw bits are zero, exactly as in the reference.
 *
 * Bit packing
 *  Watermark bits and raw extracted bits are little-endian bit arrays of
 *  uint32 words: bit c lives in word c>>5 at position c&31.
 */
#ifndef B200WM_H
#define B200WM_H

#include <stdint.h>

#if defined(__GNUC__)
#define B200WM_API __attribute__((visibility("default")))
#else
#define B200WM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define B200WM_VERSION_MAJOR 0
#define B200WM_VERSION_MINOR 1

typedef enum b200wm_status {
    B200WM_OK = 0,
    B200WM_ERR_INVALID = -1,      /* bad argument (NULL pointer, size <= 0, unsupported combination) */
    B200WM_ERR_SHORT_WM = -2,     /* watermark shorter than the number of blocks: the reference raises IndexError (dwt_dct_svd_encoder.py:36) */
    B200WM_ERR_CUDA = -3,         /* a CUDA call failed; see b200wm_last_cuda_error() */
    B200WM_ERR_NO_DEVICE = -4,    /* no CUDA device / not an sm_100 device */
    B200WM_ERR_UNSUPPORTED = -5   /* e.g. blk != 4, payload_len > limit for this call */
} b200wm_status;

typedef enum b200wm_dtype { B200WM_U8 = 0, B200WM_F32 = 1 } b200wm_dtype;

typedef struct b200wm_plane {
    int32_t dtype;               /* b200wm_dtype */
    int32_t n_frames;
    int32_t height;
    int32_t width;
    int64_t pitch_bytes;
    int64_t frame_stride_bytes;
    int32_t elem_stride;         /* in samples */
    int32_t reserved;
} b200wm_plane;

/* ---- library ------------------------------------------------------------------ */
B200WM_API int         b200wm_version(void);                 /* major*1000 + minor */
B200WM_API const char* b200wm_strerror(int status);
B200WM_API const char* b200wm_last_cuda_error(void);         /* thread-local; "" if none */
B200WM_API int         b200wm_device_ok(void);               /* 0 if the current device can run the kernels */
B200WM_API int         b200wm_kernel_launches(void);         /* kernels launched by this library since load (process-wide) */
/*
 * Kernel path for the DWT/SVD pair: 0 = automatic (TMA-staged persistent kernels when the plane
 * is planar uint8 with base, pitch and frame stride multiples of 16 bytes and >= 256 columns;
 * vectorised-load kernels otherwise), 1 = always the vectorised-load kernels.  Results are
 * identical; the switch exists so both paths can be measured and tested.  Process-wide.
 */
B200WM_API int         b200wm_set_path(int path);
B200WM_API int         b200wm_get_path(void);

/* ---- geometry helpers (host, pure arithmetic) ------------------------------------ */
/* DwtDctSvdEncoder.wm_capacity (embed/dwt_dct_svd_encoder.py:14-17): height*width/64. */
B200WM_API int64_t b200wm_block_num(int height, int width);
/* Number of blocks the reference actually walks (dwt_dct_svd_encoder.py:24,31-34). */
B200WM_API int64_t b200wm_tile_count(int height, int width);
/* uint32 words needed per frame for block_num raw bits. */
B200WM_API int32_t b200wm_words_per_frame(int height, int width);

/* ---- Haar-DWT / block-SVD quantisation-index pair -------------------------------- */
/*
 * Replaces DwtDctSvdEncoder.encode for one channel of a batch of frames
 * (embed/dwt_dct_svd_encoder.py:19-45: dwt2 'haar' -> per 4x4 LL block
 * dct/svd -> s[0] = (s[0]//scale + 0.25 + 0.5*bit)*scale -> inverse chain).
 * blk is fixed at 4.  `dst` may equal `src` (in place, like the reference) or
 * be another buffer of identical geometry that already holds the frames; only
 * the samples of the walked tiles are written.
 * U8 planes are written back as clip(sample + round-half-even(increment), 0, 255), the caller
 * bracket of video/embedder.py:37-38 (np.clip, np.around, uint8) up to exact rounding ties: the
 * reference rounds sample + increment, which differs by 1 LSB only when an increment is exactly
 * k + 0.5 (observed on 4e-6 of the samples).  F32 planes are written unrounded.
 * wm_packed: [n_wm_rows, wm_words] packed bits; frame f uses row
 * frame_wm_row[f] (NULL -> row 0 for every frame; entries are clamped into [0, n_wm_rows) on the device, so a
 * bad table can never read outside wm_packed - validate tables on the host if a wrong row must be an error).
 * wm_len = bits per row.
 */
B200WM_API int b200wm_dwtsvd_embed(const void* src, void* dst, const b200wm_plane* plane,
                        const uint32_t* wm_packed, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                        const int32_t* frame_wm_row, float scale, void* stream);

/*
 * One read, N marked copies of every frame: the "N watermarked copies of each segment" step of the
 * reference's fingerprinting flow (tests/mark_video_to_hls.py:330-354 runs the embedder once per copy
 * over the same source; payload = 4-bit segment || 4-bit copy, :27-43).  Copy c of frame f is written
 * at dst + c*copy_stride_bytes + f*plane->frame_stride_bytes with the pitch of `plane`, carries
 * watermark row copy_wm_row[f*n_copies + c] (NULL -> row c) and is bit-identical to
 * b200wm_dwtsvd_embed with that row.  dst must not overlap src and must already hold whatever the
 * caller wants outside the walked tiles (plane edges not covered by 8x8 tiles are not written).
 * uint8 planes only.  Algorithmic bytes per frame: (1 + n_copies) * W * H.
 */
B200WM_API int b200wm_dwtsvd_embed_copies(const void* src, const b200wm_plane* plane, void* dst, int64_t copy_stride_bytes,
                              int32_t n_copies, const uint32_t* wm_packed, int32_t n_wm_rows, int32_t wm_words,
                              int64_t wm_len, const int32_t* copy_wm_row, float scale, void* stream);

/*
 * Replaces DwtDctSvdDecoder.decode for one channel (extract/dwt_dct_svd_decoder.py:12-37:
 * bit = (sigma_0 % scale) > scale/2 per block) and the counting half of
 * DeShuffler.degenerate (degenerator/de_shuffler.py:17-18).
 * raw_bits  [n_frames, words_per_frame] packed bits, fully written (surplus bits 0).
 * pos_counts[n_frames, payload_len] (nullable): number of set raw bits at block
 *           indices congruent to i modulo payload_len.  Zeroed by the call.
 */
B200WM_API int b200wm_dwtsvd_extract(const void* src, const b200wm_plane* plane, float scale,
                          uint32_t* raw_bits, int32_t words_per_frame,
                          int32_t payload_len, int32_t* pos_counts, void* stream);

/* Debug/validation: sigma_0 of every walked block as float32 [n_frames, tile_count]. */
B200WM_API int b200wm_dwtsvd_sigma(const void* src, const b200wm_plane* plane, float* sigma, void* stream);
/*
 * The same quantity computed as the reference writes it, svd(cv2.dct(block)).s[0]
 * (extract/dwt_dct_svd_decoder.py:35), with orthonormal 4-point DCT-II butterflies on the LL block.
 * The production kernels drop the DCT (it is orthogonal, sigma is invariant); this entry point exists
 * so that tests can check that claim on the device.  Same output as b200wm_dwtsvd_sigma.
 */
B200WM_API int b200wm_dwtsvd_sigma_dct(const void* src, const b200wm_plane* plane, float* sigma, void* stream);

/* ---- 8x8 block-DCT quantisation-index pair --------------------------------------- */
/*
 * Replaces DctEncoder.luminance_mask + texture_mask (embed/dct_encoder.py:41-102,
 * duplicated at extract/dct_decoder.py:29-89).  `lum` is the luminance plane.
 * block_dc   [n_frames, by*bx] float32  DC coefficient / 8 of each 8x8 block (block mean)
 * tex_mask   [n_frames, by*bx] float32  texture mask
 * frame_sum  [n_frames] float64         sum of block means (for the frame-global mean); zeroed by the call
 * by = height/8, bx = width/8.
 */
B200WM_API int b200wm_dct8_masks(const void* lum, const b200wm_plane* lum_plane,
                      float* block_mean, float* tex_mask, double* frame_sum, void* stream);
/*
 * Replaces the block loop of DctEncoder.encode (embed/dct_encoder.py:24-38): QIM of
 * DCT coefficient [2][1] of every 8x8 block of the chroma plane with step
 * alpha * tex_mask * lum_mask, lum_mask derived from block_mean and frame_sum.
 */
B200WM_API int b200wm_dct8_embed(const void* src, void* dst, const b200wm_plane* plane,
                      const float* block_mean, const float* tex_mask, const double* frame_sum,
                      const uint32_t* wm_packed, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                      const int32_t* frame_wm_row, float alpha, void* stream);
/* Replaces the block loop of DctDecoder.decode (extract/dct_decoder.py:17-27). */
B200WM_API int b200wm_dct8_extract(const void* src, const b200wm_plane* plane,
                        const float* block_mean, const float* tex_mask, const double* frame_sum,
                        float alpha, uint32_t* raw_bits, int32_t words_per_frame,
                        int32_t payload_len, int32_t* pos_counts, void* stream);

/*
 * The pair as the reference calls it: DctEncoder.encode (embed/dct_encoder.py:18-39) and DctDecoder.decode
 * (extract/dct_decoder.py:10-27) each compute BOTH masks of the luminance channel and then walk the chroma blocks.
 * One call each, no caller-visible mask arrays: the masks kernel and the quantiser kernel back to back, the 8 bytes per
 * block between them in stream-ordered scratch.  `lum` and `src` are the two channels of the same frames: same dtype,
 * height, width and frame count, any pitches / strides (channel 0 and channel 1 of the reference's interleaved float32
 * frames, or the Y and a full-resolution chroma plane of planar uint8 4:4:4).  frame_sum [n_frames] float64 is scratch
 * that the call fills (the sum of the block means per frame).  Same results as b200wm_dct8_masks + b200wm_dct8_embed /
 * _extract, which remain for callers that reuse the masks of one luminance plane for embed AND extract.
 */
B200WM_API int b200wm_dct8_encode(const void* lum, const b200wm_plane* lum_plane, const void* src, void* dst, const b200wm_plane* plane,
                      double* frame_sum, const uint32_t* wm_packed, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                      const int32_t* frame_wm_row, float alpha, void* stream);
B200WM_API int b200wm_dct8_decode(const void* lum, const b200wm_plane* lum_plane, const void* src, const b200wm_plane* plane,
                      double* frame_sum, float alpha, uint32_t* raw_bits, int32_t words_per_frame, int32_t payload_len,
                      int32_t* pos_counts, void* stream);

/* ---- votes ------------------------------------------------------------------------ */
/*
 * Counting half of DeShuffler.degenerate for any payload_len (de_shuffler.py:17-18)
 * from packed raw bits.  pos_counts [n_frames, payload_len] is zeroed by the call.
 */
B200WM_API int b200wm_vote_counts(const uint32_t* raw_bits, int32_t n_frames, int32_t words_per_frame,
                       int64_t block_num, int32_t payload_len, int32_t* pos_counts, void* stream);
/*
 * Finishing half of DeShuffler.degenerate (de_shuffler.py:18-22) per frame, in the
 * reference's float64 expression: m_i = count_i / n_i, scatter through perm
 * (payload[perm[i]] = m_i), threshold 0.5*(max+min), strict '>'.
 * patterns [n_frames, payload_len] uint8 (the array degenerate() returns);
 * packed   [n_frames] uint64 (nullable; payload_len <= 64; payload bit j at bit position payload_len-1-j,
 *          i.e. the integer whose binary string is the pattern string of
 *          tests/segment_mark_detect_hls.py:145).
 * perm [payload_len] must be a permutation of 0..payload_len-1; an entry outside that range is skipped (nothing is
 * written for it), the host-buffer entry points reject such a table with B200WM_ERR_INVALID.
 */
B200WM_API int b200wm_vote_finish(const int32_t* pos_counts, int32_t n_frames, int32_t payload_len,
                       int64_t block_num, const int32_t* perm,
                       uint8_t* patterns, uint64_t* packed, void* stream);
/*
 * Cross-frame pattern vote, device half (tests/segment_mark_detect_hls.py:144-155):
 * histogram of per-frame patterns per segment plus what is needed to reproduce
 * Counter.most_common(1) exactly (ties -> pattern seen first) after an
 * exchange across GPUs (an all-gather, b200wm/vote.py).  payload_len <= 16.
 * frame_segment [n_frames] (nullable -> segment 0); frame_order [n_frames]
 * (nullable -> order_offset + f): global position of each frame in its segment.
 * hist       [n_segments, 2^payload_len] int32, accumulated (caller zeroes)
 * first_seen [n_segments, 2^payload_len] int32, min-accumulated (caller fills with INT32_MAX)
 * bit_votes  [n_segments, payload_len]   int32, accumulated: frames voting 1 per payload bit
 * seg_frames [n_segments]                int32, accumulated: frames per segment
 * Frames whose segment is outside [0, n_segments) or whose packed pattern has bits above payload_len are ignored.
 */
B200WM_API int b200wm_pattern_hist(const uint64_t* packed, const int32_t* frame_segment,
                        const int32_t* frame_order, int32_t order_offset,
                        int32_t n_frames, int32_t payload_len, int32_t n_segments,
                        int32_t* hist, int32_t* first_seen, int32_t* bit_votes,
                        int32_t* seg_frames, void* stream);

/*
 * b200wm_pattern_hist fused with the exchange of the finished state over NVLink / NVSwitch, for whole segments dealt to
 * ranks (BASELINE config 4).  Every rank keeps ONE buffer [world][block_len] int32 + `world` uint32 arrival flags (zero at
 * start) in peer-mapped ("symmetric") memory; `peer_states` is a DEVICE array of the `world` base pointers of these
 * buffers as mapped in this process (entry `rank` = the local buffer).  hist / first_seen / bit_votes / seg_frames point
 * into this rank's own block, exactly as for b200wm_pattern_hist.  The kernel accumulates the frames, then its last CTA
 * stores the block into slot `rank` of every peer's buffer (128-bit stores) and raises flag[rank] = epoch on every peer
 * (epochs increase by one per call, same value on every rank).  One-sided: the call never waits for a peer.
 * b200wm_vote_exchange_wait enqueues the wait - until every local flag shows `epoch` - for whoever reads the peers' blocks;
 * status: one int32, set to 1 if a peer did not arrive within two seconds (the wait never hangs the GPU).
 * ticket: two zeroed uint32 of local scratch.  At most 128 - 4 (world - 1) histogram CTAs (256 frames each) per call.  block_len must be a multiple of 4.
 */
B200WM_API int b200wm_pattern_hist_publish(const uint64_t* packed, const int32_t* frame_segment, const int32_t* frame_order,
                               int32_t order_offset, int32_t n_frames, int32_t payload_len, int32_t n_segments, int32_t* hist,
                               int32_t* first_seen, int32_t* bit_votes, int32_t* seg_frames, void* const* peer_states,
                               int64_t block_len, int32_t world, int32_t rank, uint32_t epoch, uint32_t* ticket, int32_t* status,
                               void* stream);

B200WM_API int b200wm_vote_exchange_wait(void* const* peer_states, int64_t block_len, int32_t world, int32_t rank, uint32_t epoch,
                             int32_t* status, void* stream);

/*
 * Reset of a vote state that lives in ONE int32 run [counters (hist | bit_votes | seg_frames) | first_seen]: the
 * first n_zero entries become 0, the rest INT32_MAX ("no patterns collected",
 * tests/segment_mark_detect_hls.py:140-142), in one launch - so that a state can be kept and reused per batch.
 */
B200WM_API int b200wm_vote_state_reset(int32_t* state, int64_t n_zero, int64_t n_total, void* stream);

/* ---- colour bracket (video/embedder.py:33-39, video/extractor.py:30-34) ------------- */
/*
 * uint8 H x W x 3 interleaved frames (what FileDecoder.read returns,
 * video/frame_reader.py:59-63) -> float32 H x W x 3 YUV with OpenCV's float
 * BGR2YUV coefficients, and back with clip / round-half-even / uint8.
 */
B200WM_API int b200wm_bgr8_to_yuv32(const uint8_t* bgr, float* yuv, int64_t n_pixels, void* stream);
B200WM_API int b200wm_yuv32_to_bgr8(const float* yuv, uint8_t* bgr, int64_t n_pixels, void* stream);

/* ---- colour bracket fused with the DWT/SVD pair -------------------------- */
/*
 * Embedder.__mark_frame (video/embedder.py:33-39) in one kernel: uint8 H x W x 3 frames as FileDecoder
 * yields them (video/frame_reader.py:59-63) -> float YUV (OpenCV BGR2YUV float formulas) -> DWT/SVD
 * embed on every channel c with scales[c] > 0 (host array of 3 floats; the reference default is
 * {0, 15, 0}) -> YUV2BGR -> clip -> round-half-even -> uint8.  dst may equal src.
 */
B200WM_API int b200wm_dwtsvd_embed_rgb8(const uint8_t* src, uint8_t* dst, int32_t n_frames, int32_t height, int32_t width,
                            int64_t pitch_bytes, int64_t frame_stride_bytes, const float* scales, const uint32_t* wm_packed,
                            int32_t n_wm_rows, int32_t wm_words, int64_t wm_len, const int32_t* frame_wm_row, void* stream);
/*
 * Extractor.__check_frame's conversion + DwtDctSvdDecoder.decode on YUV channel `channel`
 * (video/extractor.py:30-33, extract/dwt_dct_svd_decoder.py:12-37) in one kernel.  Outputs as
 * b200wm_dwtsvd_extract.
 */
B200WM_API int b200wm_dwtsvd_extract_rgb8(const uint8_t* src, int32_t n_frames, int32_t height, int32_t width, int64_t pitch_bytes,
                              int64_t frame_stride_bytes, int32_t channel, float scale, uint32_t* raw_bits,
                              int32_t words_per_frame, int32_t payload_len, int32_t* pos_counts, void* stream);

/* ---- host-buffer entry points ---------------------------------------------------------------------- */
/*
 * The whole path on frames that live in HOST memory, as the reference's drivers see them (frames
 * come from and go to ffmpeg pipes, video/frame_reader.py:53-64, video/frame_writer.py:41-44).
 * `plane` describes planar uint8 planes in host memory; every other pointer is a host pointer too.
 * The batch is streamed through the current device in chunks of `chunk_frames` (0 = automatic) on
 * three internal streams (upload, kernels and download overlap); the calls return when the results
 * are in host memory.  Pinned host memory gives full PCIe speed.
 *
 * mark:   Embedder's per-frame encode for the whole batch (b200wm_dwtsvd_embed semantics).
 * detect: DwtDctSvdDecoder.decode + DeShuffler.degenerate per frame: patterns_host [n_frames, payload_len]
 *         uint8; raw_bits_host [n_frames, words] and pos_counts_host [n_frames, payload_len] are optional.
 */
B200WM_API int b200wm_dwtsvd_mark_host(const uint8_t* src_host, uint8_t* dst_host, const b200wm_plane* plane,
                           const uint32_t* wm_packed_host, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                           const int32_t* frame_wm_row_host, float scale, int32_t chunk_frames);
B200WM_API int b200wm_dwtsvd_detect_host(const uint8_t* src_host, const b200wm_plane* plane, float scale, int32_t payload_len,
                             const int32_t* perm_host, uint8_t* patterns_host, uint32_t* raw_bits_host,
                             int32_t* pos_counts_host, int32_t chunk_frames);

/*
 * mark + verify with ONE trip over the link: upload, embed, extract + per-frame vote on the marked chunk while
 * it is still in device memory, download the marked planes: 2*W*H bytes per frame on PCIe instead of the 3*W*H
 * of the two calls above.  This is the reference's own "watermark, then always verify" step
 * (tests/mark_video_to_hls.py:356-399 re-decodes every marked segment through detect_patterns_in_segment,
 * :253-266; tests/segment_mark_detect_hls.py:195-243).  Outputs as mark (dst_host) and detect (patterns_host,
 * optional raw_bits_host / pos_counts_host); the patterns are those of the MARKED planes.
 */
B200WM_API int b200wm_dwtsvd_mark_verify_host(const uint8_t* src_host, uint8_t* dst_host, const b200wm_plane* plane,
                                  const uint32_t* wm_packed_host, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                                  const int32_t* frame_wm_row_host, float scale, int32_t payload_len, const int32_t* perm_host,
                                  uint8_t* patterns_host, uint32_t* raw_bits_host, int32_t* pos_counts_host,
                                  int32_t chunk_frames);

/* Streams and device scratch of the calls above persist between calls (per device, grow-only); this
 * frees them for the current device. */
B200WM_API int b200wm_host_scratch_release(void);

/* ---- distortion channel for robustness studies (no counterpart in the reference) ---- */
/*
 * JPEG-like requantisation of planar uint8 planes: per 8x8 block DCT(x-128), quantise and dequantise
 * with the libjpeg luminance table scaled to `quality` (1..100), IDCT, round, clip.  dst may equal src.
 */
B200WM_API int b200wm_attack_jpeg_requant(const void* src, void* dst, const b200wm_plane* plane, int32_t quality, void* stream);
/* x + noise (float32 [n_frames, height, width], contiguous), round, clip.  dst may equal src. */
B200WM_API int b200wm_attack_add_noise(const void* src, void* dst, const b200wm_plane* plane, const float* noise, void* stream);
/*
 * cv2.resize of planar uint8 planes, bit for bit as OpenCV 4.x computes it on uint8: INTER_AREA
 * (reductions only: float cell weights, or integer sums for whole-number ratios) and INTER_LINEAR
 * (11-bit fixed-point weights).  Same n_frames on both sides; src and dst must not overlap.
 * The resize attack of config 5 is AREA 1080p -> 720p followed by LINEAR 720p -> 1080p.
 */
#define B200WM_INTER_LINEAR 1   /* cv2.INTER_LINEAR */
#define B200WM_INTER_AREA 3     /* cv2.INTER_AREA */
B200WM_API int b200wm_attack_resize(const void* src, const b200wm_plane* src_plane, void* dst, const b200wm_plane* dst_plane,
                        int32_t interpolation, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200WM_H */
