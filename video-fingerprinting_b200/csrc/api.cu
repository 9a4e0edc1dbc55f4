// extern "C" surface of libb200wm.so (see include/b200wm.h for the contract).
#include "common.cuh"
#include <atomic>
#include <cstdio>
#include <cstring>

namespace b200wm {

static thread_local char g_cuda_error[512] = "";
static std::atomic<int> g_launches{0};

void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int retain_async_pool() {
    static std::atomic<unsigned long long> done{0};
    int dev = 0;
    B200WM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && ((done.load(std::memory_order_relaxed) >> dev) & 1ull)) return B200WM_OK;
    cudaMemPool_t pool;
    B200WM_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ull;
    B200WM_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_relaxed);
    return B200WM_OK;
}

int launch_dwtsvd_embed(const void*, void*, const b200wm_plane*, const uint32_t*, int, int, long long, const int32_t*, float,
                        cudaStream_t);
int launch_dwtsvd_extract(const void*, const b200wm_plane*, float, uint32_t*, int, int, int32_t*, float*, cudaStream_t);
int launch_vote_counts(const uint32_t*, int, int, long long, int, int32_t*, cudaStream_t);
int launch_vote_finish(const int32_t*, int, int, long long, const int32_t*, uint8_t*, uint64_t*, cudaStream_t);
int launch_pattern_hist(const uint64_t*, const int32_t*, const int32_t*, int, int, int, int, int32_t*, int32_t*, int32_t*,
                        int32_t*, cudaStream_t);
int launch_vote_state_reset(int32_t*, long long, long long, cudaStream_t);
int launch_vote_exchange_wait(void* const*, long long, int, int, unsigned, int*, cudaStream_t);
int launch_pattern_hist_publish(const uint64_t*, const int32_t*, const int32_t*, int, int, int, int, int32_t*, int32_t*, int32_t*,
                                int32_t*, void* const*, long long, int, int, unsigned, unsigned*, int*, cudaStream_t);
int launch_dct8_masks(const void*, const b200wm_plane*, float*, float*, double*, cudaStream_t);
int launch_dct8_embed(const void*, void*, const b200wm_plane*, const float*, const float*, const double*, const uint32_t*,
                      int, int, long long, const int32_t*, float, cudaStream_t);
int launch_dct8_extract(const void*, const b200wm_plane*, const float*, const float*, const double*, float, uint32_t*, int,
                        int, int32_t*, cudaStream_t);
int launch_dct8_encode(const void*, const b200wm_plane*, const void*, void*, const b200wm_plane*, double*, const uint32_t*, int, int,
                       long long, const int32_t*, float, cudaStream_t);
int launch_dct8_decode(const void*, const b200wm_plane*, const void*, const b200wm_plane*, double*, float, uint32_t*, int, int,
                       int32_t*, cudaStream_t);
int launch_bgr8_to_yuv32(const uint8_t*, float*, long long, cudaStream_t);
int launch_attack_jpeg(const void*, void*, const b200wm_plane*, int, cudaStream_t);
int launch_dwtsvd_embed_copies(const void*, const b200wm_plane*, void*, long long, int, const uint32_t*, int, int, long long,
                               const int32_t*, float, cudaStream_t);
int launch_attack_noise(const void*, void*, const b200wm_plane*, const float*, cudaStream_t);
int launch_attack_resize(const void*, const b200wm_plane*, void*, const b200wm_plane*, int, cudaStream_t);
int launch_embed_rgb8(const uint8_t*, uint8_t*, int, int, int, long long, long long, const float*, const uint32_t*, int, int, long long,
                      const int32_t*, cudaStream_t);
int launch_extract_rgb8(const uint8_t*, int, int, int, long long, long long, int, float, uint32_t*, int, int, int32_t*, cudaStream_t);
int mark_host(const uint8_t*, uint8_t*, const b200wm_plane*, const uint32_t*, int, int, long long, const int32_t*, float, int);
int detect_host(const uint8_t*, const b200wm_plane*, float, int, const int32_t*, uint8_t*, uint32_t*, int32_t*, int);
int mark_verify_host(const uint8_t*, uint8_t*, const b200wm_plane*, const uint32_t*, int, int, long long, const int32_t*, float, int,
                     const int32_t*, uint8_t*, uint32_t*, int32_t*, int);
int host_scratch_release();
int launch_dwtsvd_sigma_dct(const void*, const b200wm_plane*, float*, cudaStream_t);
void set_path(int);
int get_path();
int launch_yuv32_to_bgr8(const float*, uint8_t*, long long, cudaStream_t);

}  // namespace b200wm

using namespace b200wm;

extern "C" {

B200WM_API int b200wm_version(void) { return B200WM_VERSION_MAJOR * 1000 + B200WM_VERSION_MINOR; }

B200WM_API const char* b200wm_strerror(int status) {
    switch (status) {
        case B200WM_OK: return "ok";
        case B200WM_ERR_INVALID: return "invalid argument";
        case B200WM_ERR_SHORT_WM: return "watermark shorter than the number of blocks";
        case B200WM_ERR_CUDA: return "CUDA error";
        case B200WM_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required)";
        case B200WM_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

B200WM_API const char* b200wm_last_cuda_error(void) { return g_cuda_error; }

B200WM_API int b200wm_device_ok(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_cuda_error(e, "cudaGetDevice"); return B200WM_ERR_NO_DEVICE; }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) { set_cuda_error(e, "cudaDeviceGetAttribute"); return B200WM_ERR_NO_DEVICE; }
    return major == 10 ? B200WM_OK : B200WM_ERR_NO_DEVICE;
}

B200WM_API int b200wm_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

B200WM_API int b200wm_set_path(int path) {
    if (path != 0 && path != 1) return B200WM_ERR_INVALID;
    set_path(path);
    return B200WM_OK;
}
B200WM_API int b200wm_get_path(void) { return get_path(); }

B200WM_API int64_t b200wm_block_num(int height, int width) { return (int64_t)height * width / 64; }
B200WM_API int64_t b200wm_tile_count(int height, int width) {
    if (height <= 0 || width <= 0) return 0;
    return make_geom(height, width).n_tiles;
}
B200WM_API int32_t b200wm_words_per_frame(int height, int width) {
    if (height <= 0 || width <= 0) return 0;
    return make_geom(height, width).words;
}

B200WM_API int b200wm_dwtsvd_embed(const void* src, void* dst, const b200wm_plane* plane, const uint32_t* wm_packed,
                        int32_t n_wm_rows, int32_t wm_words, int64_t wm_len, const int32_t* frame_wm_row, float scale, void* stream) {
    return launch_dwtsvd_embed(src, dst, plane, wm_packed, n_wm_rows, wm_words, wm_len, frame_wm_row, scale, (cudaStream_t)stream);
}

B200WM_API int b200wm_dwtsvd_embed_copies(const void* src, const b200wm_plane* plane, void* dst, int64_t copy_stride_bytes,
                              int32_t n_copies, const uint32_t* wm_packed, int32_t n_wm_rows, int32_t wm_words,
                              int64_t wm_len, const int32_t* copy_wm_row, float scale, void* stream) {
    return launch_dwtsvd_embed_copies(src, plane, dst, copy_stride_bytes, n_copies, wm_packed, n_wm_rows, wm_words, wm_len,
                                      copy_wm_row, scale, (cudaStream_t)stream);
}

B200WM_API int b200wm_dwtsvd_extract(const void* src, const b200wm_plane* plane, float scale, uint32_t* raw_bits,
                          int32_t words_per_frame, int32_t payload_len, int32_t* pos_counts, void* stream) {
    return launch_dwtsvd_extract(src, plane, scale, raw_bits, words_per_frame, payload_len, pos_counts, nullptr,
                                 (cudaStream_t)stream);
}

B200WM_API int b200wm_dwtsvd_sigma_dct(const void* src, const b200wm_plane* plane, float* sigma, void* stream) {
    return launch_dwtsvd_sigma_dct(src, plane, sigma, (cudaStream_t)stream);
}

B200WM_API int b200wm_dwtsvd_sigma(const void* src, const b200wm_plane* plane, float* sigma, void* stream) {
    if (!sigma || !plane) return B200WM_ERR_INVALID;
    // raw bits are a by-product; park them in a scratch allocation private to this debug call
    const int words = b200wm_words_per_frame(plane->height, plane->width);
    uint32_t* scratch = nullptr;
    const size_t bytes = sizeof(uint32_t) * (size_t)(words > 0 ? words : 1) * (size_t)(plane->n_frames > 0 ? plane->n_frames : 1);
    B200WM_CUDA_TRY(cudaMallocAsync((void**)&scratch, bytes, (cudaStream_t)stream));
    const int rc = launch_dwtsvd_extract(src, plane, 15.0f, scratch, words, 0, nullptr, sigma, (cudaStream_t)stream);
    cudaFreeAsync(scratch, (cudaStream_t)stream);
    return rc;
}

B200WM_API int b200wm_dct8_masks(const void* lum, const b200wm_plane* lum_plane, float* block_mean, float* tex_mask,
                      double* frame_sum, void* stream) {
    return launch_dct8_masks(lum, lum_plane, block_mean, tex_mask, frame_sum, (cudaStream_t)stream);
}

B200WM_API int b200wm_dct8_embed(const void* src, void* dst, const b200wm_plane* plane, const float* block_mean,
                      const float* tex_mask, const double* frame_sum, const uint32_t* wm_packed, int32_t n_wm_rows,
                      int32_t wm_words, int64_t wm_len, const int32_t* frame_wm_row, float alpha, void* stream) {
    return launch_dct8_embed(src, dst, plane, block_mean, tex_mask, frame_sum, wm_packed, n_wm_rows, wm_words, wm_len, frame_wm_row,
                             alpha, (cudaStream_t)stream);
}

B200WM_API int b200wm_dct8_extract(const void* src, const b200wm_plane* plane, const float* block_mean, const float* tex_mask,
                        const double* frame_sum, float alpha, uint32_t* raw_bits, int32_t words_per_frame,
                        int32_t payload_len, int32_t* pos_counts, void* stream) {
    return launch_dct8_extract(src, plane, block_mean, tex_mask, frame_sum, alpha, raw_bits, words_per_frame, payload_len,
                               pos_counts, (cudaStream_t)stream);
}

B200WM_API int b200wm_dct8_encode(const void* lum, const b200wm_plane* lum_plane, const void* src, void* dst, const b200wm_plane* plane,
                      double* frame_sum, const uint32_t* wm_packed, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                      const int32_t* frame_wm_row, float alpha, void* stream) {
    return launch_dct8_encode(lum, lum_plane, src, dst, plane, frame_sum, wm_packed, n_wm_rows, wm_words, wm_len, frame_wm_row, alpha,
                              (cudaStream_t)stream);
}

B200WM_API int b200wm_dct8_decode(const void* lum, const b200wm_plane* lum_plane, const void* src, const b200wm_plane* plane,
                      double* frame_sum, float alpha, uint32_t* raw_bits, int32_t words_per_frame, int32_t payload_len,
                      int32_t* pos_counts, void* stream) {
    return launch_dct8_decode(lum, lum_plane, src, plane, frame_sum, alpha, raw_bits, words_per_frame, payload_len, pos_counts,
                              (cudaStream_t)stream);
}

B200WM_API int b200wm_vote_counts(const uint32_t* raw_bits, int32_t n_frames, int32_t words_per_frame, int64_t block_num,
                       int32_t payload_len, int32_t* pos_counts, void* stream) {
    return launch_vote_counts(raw_bits, n_frames, words_per_frame, block_num, payload_len, pos_counts, (cudaStream_t)stream);
}

B200WM_API int b200wm_vote_finish(const int32_t* pos_counts, int32_t n_frames, int32_t payload_len, int64_t block_num,
                       const int32_t* perm, uint8_t* patterns, uint64_t* packed, void* stream) {
    return launch_vote_finish(pos_counts, n_frames, payload_len, block_num, perm, patterns, packed, (cudaStream_t)stream);
}

B200WM_API int b200wm_pattern_hist(const uint64_t* packed, const int32_t* frame_segment, const int32_t* frame_order,
                        int32_t order_offset, int32_t n_frames, int32_t payload_len, int32_t n_segments, int32_t* hist,
                        int32_t* first_seen, int32_t* bit_votes, int32_t* seg_frames, void* stream) {
    return launch_pattern_hist(packed, frame_segment, frame_order, order_offset, n_frames, payload_len, n_segments, hist,
                               first_seen, bit_votes, seg_frames, (cudaStream_t)stream);
}

B200WM_API int b200wm_pattern_hist_publish(const uint64_t* packed, const int32_t* frame_segment, const int32_t* frame_order,
                               int32_t order_offset, int32_t n_frames, int32_t payload_len, int32_t n_segments, int32_t* hist,
                               int32_t* first_seen, int32_t* bit_votes, int32_t* seg_frames, void* const* peer_states,
                               int64_t block_len, int32_t world, int32_t rank, uint32_t epoch, uint32_t* ticket, int32_t* status,
                               void* stream) {
    return launch_pattern_hist_publish(packed, frame_segment, frame_order, order_offset, n_frames, payload_len, n_segments, hist,
                                       first_seen, bit_votes, seg_frames, peer_states, block_len, world, rank, epoch, ticket, status,
                                       (cudaStream_t)stream);
}

B200WM_API int b200wm_vote_exchange_wait(void* const* peer_states, int64_t block_len, int32_t world, int32_t rank, uint32_t epoch,
                             int32_t* status, void* stream) {
    return launch_vote_exchange_wait(peer_states, block_len, world, rank, epoch, status, (cudaStream_t)stream);
}

B200WM_API int b200wm_vote_state_reset(int32_t* state, int64_t n_zero, int64_t n_total, void* stream) {
    return launch_vote_state_reset(state, n_zero, n_total, (cudaStream_t)stream);
}

B200WM_API int b200wm_bgr8_to_yuv32(const uint8_t* bgr, float* yuv, int64_t n_pixels, void* stream) {
    return launch_bgr8_to_yuv32(bgr, yuv, n_pixels, (cudaStream_t)stream);
}

B200WM_API int b200wm_yuv32_to_bgr8(const float* yuv, uint8_t* bgr, int64_t n_pixels, void* stream) {
    return launch_yuv32_to_bgr8(yuv, bgr, n_pixels, (cudaStream_t)stream);
}

B200WM_API int b200wm_host_scratch_release(void) { return host_scratch_release(); }

B200WM_API int b200wm_attack_jpeg_requant(const void* src, void* dst, const b200wm_plane* plane, int32_t quality, void* stream) {
    return launch_attack_jpeg(src, dst, plane, quality, (cudaStream_t)stream);
}

B200WM_API int b200wm_attack_add_noise(const void* src, void* dst, const b200wm_plane* plane, const float* noise, void* stream) {
    return launch_attack_noise(src, dst, plane, noise, (cudaStream_t)stream);
}

B200WM_API int b200wm_attack_resize(const void* src, const b200wm_plane* src_plane, void* dst, const b200wm_plane* dst_plane,
                        int32_t interpolation, void* stream) {
    return launch_attack_resize(src, src_plane, dst, dst_plane, interpolation, (cudaStream_t)stream);
}

B200WM_API int b200wm_dwtsvd_embed_rgb8(const uint8_t* src, uint8_t* dst, int32_t n_frames, int32_t height, int32_t width,
                            int64_t pitch_bytes, int64_t frame_stride_bytes, const float* scales, const uint32_t* wm_packed,
                            int32_t n_wm_rows, int32_t wm_words, int64_t wm_len, const int32_t* frame_wm_row, void* stream) {
    return launch_embed_rgb8(src, dst, n_frames, height, width, pitch_bytes, frame_stride_bytes, scales, wm_packed, n_wm_rows, wm_words,
                             wm_len, frame_wm_row, (cudaStream_t)stream);
}

B200WM_API int b200wm_dwtsvd_extract_rgb8(const uint8_t* src, int32_t n_frames, int32_t height, int32_t width, int64_t pitch_bytes,
                              int64_t frame_stride_bytes, int32_t channel, float scale, uint32_t* raw_bits,
                              int32_t words_per_frame, int32_t payload_len, int32_t* pos_counts, void* stream) {
    return launch_extract_rgb8(src, n_frames, height, width, pitch_bytes, frame_stride_bytes, channel, scale, raw_bits,
                               words_per_frame, payload_len, pos_counts, (cudaStream_t)stream);
}

B200WM_API int b200wm_dwtsvd_mark_host(const uint8_t* src_host, uint8_t* dst_host, const b200wm_plane* plane,
                           const uint32_t* wm_packed_host, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                           const int32_t* frame_wm_row_host, float scale, int32_t chunk_frames) {
    return mark_host(src_host, dst_host, plane, wm_packed_host, n_wm_rows, wm_words, wm_len, frame_wm_row_host, scale, chunk_frames);
}

B200WM_API int b200wm_dwtsvd_detect_host(const uint8_t* src_host, const b200wm_plane* plane, float scale, int32_t payload_len,
                             const int32_t* perm_host, uint8_t* patterns_host, uint32_t* raw_bits_host,
                             int32_t* pos_counts_host, int32_t chunk_frames) {
    return detect_host(src_host, plane, scale, payload_len, perm_host, patterns_host, raw_bits_host, pos_counts_host, chunk_frames);
}

B200WM_API int b200wm_dwtsvd_mark_verify_host(const uint8_t* src_host, uint8_t* dst_host, const b200wm_plane* plane,
                                  const uint32_t* wm_packed_host, int32_t n_wm_rows, int32_t wm_words, int64_t wm_len,
                                  const int32_t* frame_wm_row_host, float scale, int32_t payload_len, const int32_t* perm_host,
                                  uint8_t* patterns_host, uint32_t* raw_bits_host, int32_t* pos_counts_host,
                                  int32_t chunk_frames) {
    return mark_verify_host(src_host, dst_host, plane, wm_packed_host, n_wm_rows, wm_words, wm_len, frame_wm_row_host, scale,
                            payload_len, perm_host, patterns_host, raw_bits_host, pos_counts_host, chunk_frames);
}

}  // extern "C"
