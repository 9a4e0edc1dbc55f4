// Host-buffer entry points: the caller's frames live in HOST memory (the situation of the
// reference's drivers, which get frames from an ffmpeg pipe: src/offmark/video/frame_reader.py:53-64
// and hand them back to one: frame_writer.py:41-44).  The batch is streamed through the GPU in
// chunks over three device buffers and three streams (upload, kernels, download) chained by events,
// so that the upload of chunk i+1, the kernels of chunk i and the download of chunk i-1 overlap
// (PCIe is full duplex).  Streams and device scratch persist between calls (per device, grow-only).
// Pinned host memory gives full PCIe speed; pageable memory works but is staged by the driver.
#include <mutex>
#include "common.cuh"

namespace b200wm {

int validate_plane(const b200wm_plane* pl);
int launch_dwtsvd_embed(const void*, void*, const b200wm_plane*, const uint32_t*, int, int, long long, const int32_t*, float,
                        cudaStream_t);
int launch_dwtsvd_extract(const void*, const b200wm_plane*, float, uint32_t*, int, int, int32_t*, float*, cudaStream_t);
int launch_vote_finish(const int32_t*, int, int, long long, const int32_t*, uint8_t*, uint64_t*, cudaStream_t);

namespace {

// copy `frames` planes between a strided host layout and a dense [frames, h, w] device chunk
int copy_planes(void* dst, const void* src, const b200wm_plane* pl, int frames, bool to_device, cudaStream_t s) {
    const size_t row = (size_t)pl->width, plane = row * pl->height;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (pl->pitch_bytes == pl->width && (frames == 1 || pl->frame_stride_bytes == (long long)plane)) {
        // the whole chunk is one contiguous run on both sides
        B200WM_CUDA_TRY(cudaMemcpyAsync(dst, src, plane * frames, kind, s));
        return B200WM_OK;
    }
    if (pl->pitch_bytes == pl->width) {
        // every plane is one contiguous run (e.g. the Y plane of an I420 frame): one plain copy per
        // frame - measured much faster than a single 2-D copy whose "rows" are whole planes
        for (int f = 0; f < frames; ++f) {
            if (to_device)
                B200WM_CUDA_TRY(cudaMemcpyAsync((uint8_t*)dst + f * plane, (const uint8_t*)src + (size_t)f * pl->frame_stride_bytes,
                                                plane, kind, s));
            else
                B200WM_CUDA_TRY(cudaMemcpyAsync((uint8_t*)dst + (size_t)f * pl->frame_stride_bytes, (const uint8_t*)src + f * plane,
                                                plane, kind, s));
        }
        return B200WM_OK;
    }
    for (int f = 0; f < frames; ++f) {
        if (to_device)
            B200WM_CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)dst + f * plane, row, (const uint8_t*)src + (size_t)f * pl->frame_stride_bytes,
                                              (size_t)pl->pitch_bytes, row, pl->height, kind, s));
        else
            B200WM_CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)dst + (size_t)f * pl->frame_stride_bytes, (size_t)pl->pitch_bytes,
                                              (const uint8_t*)src + f * plane, row, row, pl->height, kind, s));
    }
    return B200WM_OK;
}

int check_host_plane(const b200wm_plane* pl) {
    int rc = validate_plane(pl);
    if (rc) return rc;
    if (pl->dtype != B200WM_U8 || pl->elem_stride != 1) return B200WM_ERR_UNSUPPORTED;
    if (pl->n_frames > 1 && pl->frame_stride_bytes < pl->pitch_bytes * (long long)pl->height) return B200WM_ERR_INVALID;
    return B200WM_OK;
}

int pick_chunk(int requested, const b200wm_plane* pl) {
    if (requested > 0) return requested < pl->n_frames ? requested : (pl->n_frames > 0 ? pl->n_frames : 1);
    const long long plane = (long long)pl->width * pl->height;
    long long c = (64ll << 20) / (plane > 0 ? plane : 1);        // ~64 MB per chunk: 32 frames of 1080p luma
    if (c < 1) c = 1;
    if (c > pl->n_frames) c = pl->n_frames > 0 ? pl->n_frames : 1;
    return (int)c;
}

// Streaming context of one device, kept between calls: three chunk buffers, one stream per direction
// (upload, kernels, download) and the events that chain them.  Creating and freeing a few hundred
// megabytes of device memory on every call cost more than the copies themselves and made the timings
// erratic, so the scratch only ever grows; b200wm_host_scratch_release() gives it back.
constexpr int kSlots = 3, kMaxDevices = 64;

struct Grow {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B200WM_OK;
        if (p) B200WM_CUDA_TRY(cudaFree(p));
        p = nullptr;
        cap = 0;
        B200WM_CUDA_TRY(cudaMalloc(&p, bytes));
        cap = bytes;
        return B200WM_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostCtx {
    std::mutex mu;                       // one streaming call at a time per device
    bool ready = false;
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    cudaEvent_t uploaded[kSlots], computed[kSlots], drained[kSlots];
    Grow buf[kSlots], raw[kSlots], wm, rows, perm, pat, cnt;

    int init() {
        if (ready) return B200WM_OK;
        B200WM_CUDA_TRY(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        B200WM_CUDA_TRY(cudaStreamCreateWithFlags(&run, cudaStreamNonBlocking));
        B200WM_CUDA_TRY(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
        for (int i = 0; i < kSlots; ++i) {
            B200WM_CUDA_TRY(cudaEventCreateWithFlags(&uploaded[i], cudaEventDisableTiming));
            B200WM_CUDA_TRY(cudaEventCreateWithFlags(&computed[i], cudaEventDisableTiming));
            B200WM_CUDA_TRY(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming));
        }
        ready = true;
        return B200WM_OK;
    }
    int sync_all() {
        B200WM_CUDA_TRY(cudaStreamSynchronize(up));
        B200WM_CUDA_TRY(cudaStreamSynchronize(run));
        B200WM_CUDA_TRY(cudaStreamSynchronize(down));
        return B200WM_OK;
    }
    void release() {
        if (ready) {
            cudaStreamSynchronize(up); cudaStreamSynchronize(run); cudaStreamSynchronize(down);
            for (int i = 0; i < kSlots; ++i) { cudaEventDestroy(uploaded[i]); cudaEventDestroy(computed[i]); cudaEventDestroy(drained[i]); }
            cudaStreamDestroy(up); cudaStreamDestroy(run); cudaStreamDestroy(down);
            ready = false;
        }
        for (int i = 0; i < kSlots; ++i) { buf[i].release(); raw[i].release(); }
        wm.release(); rows.release(); perm.release(); pat.release(); cnt.release();
    }
};

HostCtx g_ctx[kMaxDevices];

int current_ctx(HostCtx** ctx) {
    int dev = 0;
    B200WM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return B200WM_ERR_UNSUPPORTED;
    *ctx = &g_ctx[dev];
    return B200WM_OK;
}

}  // namespace

int host_scratch_release() {
    HostCtx* c = nullptr;
    int rc = current_ctx(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    c->release();
    return B200WM_OK;
}

int mark_host(const uint8_t* src, uint8_t* dst, const b200wm_plane* pl, const uint32_t* wm_host, int n_rows, int wm_words,
              long long wm_len, const int32_t* frame_row_host, float scale, int chunk_frames) {
    int rc = check_host_plane(pl);
    if (rc) return rc;
    if (!src || !dst || !wm_host || n_rows <= 0 || wm_words <= 0) return B200WM_ERR_INVALID;
    if (frame_row_host)
        for (int f = 0; f < pl->n_frames; ++f)
            if (frame_row_host[f] < 0 || frame_row_host[f] >= n_rows) return B200WM_ERR_INVALID;
    if (pl->n_frames == 0) return B200WM_OK;
    const int chunk = pick_chunk(chunk_frames, pl);
    const size_t plane = (size_t)pl->width * pl->height;
    HostCtx* c = nullptr;
    if ((rc = current_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    if ((rc = c->init())) return rc;
    for (int i = 0; i < kSlots; ++i)
        if ((rc = c->buf[i].reserve(plane * chunk))) return rc;
    if ((rc = c->wm.reserve(sizeof(uint32_t) * (size_t)n_rows * wm_words))) return rc;
    B200WM_CUDA_TRY(cudaMemcpyAsync(c->wm.p, wm_host, sizeof(uint32_t) * (size_t)n_rows * wm_words, cudaMemcpyHostToDevice, c->run));
    if (frame_row_host) {
        if ((rc = c->rows.reserve(sizeof(int32_t) * (size_t)pl->n_frames))) return rc;
        B200WM_CUDA_TRY(cudaMemcpyAsync(c->rows.p, frame_row_host, sizeof(int32_t) * (size_t)pl->n_frames, cudaMemcpyHostToDevice, c->run));
    }
    // enqueue everything, then drain the three streams even if an enqueue failed (work already queued
    // still points into the caller's buffers)
    rc = [&]() -> int {
    int rc = B200WM_OK;
    int k = 0;
    for (int f0 = 0; f0 < pl->n_frames; f0 += chunk, ++k) {
        const int m = pl->n_frames - f0 < chunk ? pl->n_frames - f0 : chunk;
        const int s = k % kSlots;
        uint8_t* buf = (uint8_t*)c->buf[s].p;
        b200wm_plane host = *pl, devp = *pl;
        host.n_frames = devp.n_frames = m;
        devp.pitch_bytes = pl->width;
        devp.frame_stride_bytes = (long long)plane;
        if (k >= kSlots) B200WM_CUDA_TRY(cudaStreamWaitEvent(c->up, c->drained[s], 0));     // the slot's previous chunk is back on the host
        if ((rc = copy_planes(buf, src + (size_t)f0 * pl->frame_stride_bytes, &host, m, true, c->up))) return rc;
        B200WM_CUDA_TRY(cudaEventRecord(c->uploaded[s], c->up));
        B200WM_CUDA_TRY(cudaStreamWaitEvent(c->run, c->uploaded[s], 0));
        if ((rc = launch_dwtsvd_embed(buf, buf, &devp, (const uint32_t*)c->wm.p, n_rows, wm_words, wm_len,
                                      frame_row_host ? (const int32_t*)c->rows.p + f0 : nullptr, scale, c->run)))
            return rc;
        B200WM_CUDA_TRY(cudaEventRecord(c->computed[s], c->run));
        B200WM_CUDA_TRY(cudaStreamWaitEvent(c->down, c->computed[s], 0));
        if ((rc = copy_planes(dst + (size_t)f0 * pl->frame_stride_bytes, buf, &host, m, false, c->down))) return rc;
        B200WM_CUDA_TRY(cudaEventRecord(c->drained[s], c->down));
    }
    return B200WM_OK;
    }();
    const int rc_sync = c->sync_all();
    return rc ? rc : rc_sync;
}

// The de-shuffling table scatters vote j to payload position perm[j]: every entry must be a position.
static bool perm_in_range(const int32_t* perm_host, int payload_len) {
    for (int j = 0; j < payload_len; ++j)
        if (perm_host[j] < 0 || perm_host[j] >= payload_len) return false;
    return true;
}

int detect_host(const uint8_t* src, const b200wm_plane* pl, float scale, int payload_len, const int32_t* perm_host,
                uint8_t* patterns_host, uint32_t* raw_bits_host, int32_t* pos_counts_host, int chunk_frames) {
    int rc = check_host_plane(pl);
    if (rc) return rc;
    if (!src || !perm_host || !patterns_host || payload_len <= 0) return B200WM_ERR_INVALID;
    if (!perm_in_range(perm_host, payload_len)) return B200WM_ERR_INVALID;
    if (pl->n_frames == 0) return B200WM_OK;
    const int chunk = pick_chunk(chunk_frames, pl);
    const size_t plane = (size_t)pl->width * pl->height;
    const TileGeom g = make_geom(pl->height, pl->width);
    const size_t words = g.words ? g.words : 1;
    HostCtx* c = nullptr;
    if ((rc = current_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    if ((rc = c->init())) return rc;
    for (int i = 0; i < kSlots; ++i) {
        if ((rc = c->buf[i].reserve(plane * chunk))) return rc;
        if ((rc = c->raw[i].reserve(sizeof(uint32_t) * (size_t)chunk * words))) return rc;
    }
    // patterns and counts of the WHOLE batch stay on the device until the end (8 + 32 bytes per frame):
    // a download into pageable host memory after every chunk would stall the upload queue
    if ((rc = c->pat.reserve((size_t)pl->n_frames * payload_len))) return rc;
    if ((rc = c->cnt.reserve(sizeof(int32_t) * (size_t)pl->n_frames * payload_len))) return rc;
    if ((rc = c->perm.reserve(sizeof(int32_t) * (size_t)payload_len))) return rc;
    B200WM_CUDA_TRY(cudaMemcpyAsync(c->perm.p, perm_host, sizeof(int32_t) * (size_t)payload_len, cudaMemcpyHostToDevice, c->run));
    rc = [&]() -> int {
    int rc = B200WM_OK;
    int k = 0;
    for (int f0 = 0; f0 < pl->n_frames; f0 += chunk, ++k) {
        const int m = pl->n_frames - f0 < chunk ? pl->n_frames - f0 : chunk;
        const int s = k % kSlots;
        uint8_t* buf = (uint8_t*)c->buf[s].p;
        uint32_t* raw = (uint32_t*)c->raw[s].p;
        int32_t* cnt = (int32_t*)c->cnt.p + (size_t)f0 * payload_len;
        uint8_t* pat = (uint8_t*)c->pat.p + (size_t)f0 * payload_len;
        b200wm_plane host = *pl, devp = *pl;
        host.n_frames = devp.n_frames = m;
        devp.pitch_bytes = pl->width;
        devp.frame_stride_bytes = (long long)plane;
        if (k >= kSlots) B200WM_CUDA_TRY(cudaStreamWaitEvent(c->up, c->drained[s], 0));     // the slot's previous chunk has been read
        if ((rc = copy_planes(buf, src + (size_t)f0 * pl->frame_stride_bytes, &host, m, true, c->up))) return rc;
        B200WM_CUDA_TRY(cudaEventRecord(c->uploaded[s], c->up));
        B200WM_CUDA_TRY(cudaStreamWaitEvent(c->run, c->uploaded[s], 0));
        if ((rc = launch_dwtsvd_extract(buf, &devp, scale, raw, g.words, payload_len, cnt, nullptr, c->run))) return rc;
        if ((rc = launch_vote_finish(cnt, m, payload_len, g.block_num, (const int32_t*)c->perm.p, pat, nullptr, c->run))) return rc;
        B200WM_CUDA_TRY(cudaEventRecord(c->computed[s], c->run));
        if (raw_bits_host && g.words) {
            B200WM_CUDA_TRY(cudaStreamWaitEvent(c->down, c->computed[s], 0));
            B200WM_CUDA_TRY(cudaMemcpyAsync(raw_bits_host + (size_t)f0 * g.words, raw, sizeof(uint32_t) * (size_t)m * g.words,
                                            cudaMemcpyDeviceToHost, c->down));
            B200WM_CUDA_TRY(cudaEventRecord(c->drained[s], c->down));
        } else {
            B200WM_CUDA_TRY(cudaEventRecord(c->drained[s], c->run));
        }
    }
    B200WM_CUDA_TRY(cudaMemcpyAsync(patterns_host, c->pat.p, (size_t)pl->n_frames * payload_len, cudaMemcpyDeviceToHost, c->run));
    if (pos_counts_host)
        B200WM_CUDA_TRY(cudaMemcpyAsync(pos_counts_host, c->cnt.p, sizeof(int32_t) * (size_t)pl->n_frames * payload_len,
                                        cudaMemcpyDeviceToHost, c->run));
    return B200WM_OK;
    }();
    const int rc_sync = c->sync_all();
    return rc ? rc : rc_sync;
}

// Mark and verify in one pass over the link: upload a chunk, embed in place, extract + per-frame vote from the
// marked chunk while it is still resident, download the marked planes.  This is the reference's own "watermark,
// then always verify" step (tests/mark_video_to_hls.py:356-399 re-reads every marked segment through
// detect_patterns_in_segment, :253-266; tests/segment_mark_detect_hls.py:195-243 does the same per segment) with
// 2*W*H bytes per frame on the link instead of the 3*W*H of mark_host + detect_host.
int mark_verify_host(const uint8_t* src, uint8_t* dst, const b200wm_plane* pl, const uint32_t* wm_host, int n_rows, int wm_words,
                     long long wm_len, const int32_t* frame_row_host, float scale, int payload_len, const int32_t* perm_host,
                     uint8_t* patterns_host, uint32_t* raw_bits_host, int32_t* pos_counts_host, int chunk_frames) {
    int rc = check_host_plane(pl);
    if (rc) return rc;
    if (!src || !dst || !wm_host || n_rows <= 0 || wm_words <= 0 || !perm_host || !patterns_host || payload_len <= 0)
        return B200WM_ERR_INVALID;
    if (frame_row_host)
        for (int f = 0; f < pl->n_frames; ++f)
            if (frame_row_host[f] < 0 || frame_row_host[f] >= n_rows) return B200WM_ERR_INVALID;
    if (!perm_in_range(perm_host, payload_len)) return B200WM_ERR_INVALID;
    if (pl->n_frames == 0) return B200WM_OK;
    const int chunk = pick_chunk(chunk_frames, pl);
    const size_t plane = (size_t)pl->width * pl->height;
    const TileGeom g = make_geom(pl->height, pl->width);
    const size_t words = g.words ? g.words : 1;
    HostCtx* c = nullptr;
    if ((rc = current_ctx(&c))) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    if ((rc = c->init())) return rc;
    for (int i = 0; i < kSlots; ++i) {
        if ((rc = c->buf[i].reserve(plane * chunk))) return rc;
        if ((rc = c->raw[i].reserve(sizeof(uint32_t) * (size_t)chunk * words))) return rc;
    }
    if ((rc = c->wm.reserve(sizeof(uint32_t) * (size_t)n_rows * wm_words))) return rc;
    if ((rc = c->pat.reserve((size_t)pl->n_frames * payload_len))) return rc;
    if ((rc = c->cnt.reserve(sizeof(int32_t) * (size_t)pl->n_frames * payload_len))) return rc;
    if ((rc = c->perm.reserve(sizeof(int32_t) * (size_t)payload_len))) return rc;
    B200WM_CUDA_TRY(cudaMemcpyAsync(c->wm.p, wm_host, sizeof(uint32_t) * (size_t)n_rows * wm_words, cudaMemcpyHostToDevice, c->run));
    B200WM_CUDA_TRY(cudaMemcpyAsync(c->perm.p, perm_host, sizeof(int32_t) * (size_t)payload_len, cudaMemcpyHostToDevice, c->run));
    if (frame_row_host) {
        if ((rc = c->rows.reserve(sizeof(int32_t) * (size_t)pl->n_frames))) return rc;
        B200WM_CUDA_TRY(cudaMemcpyAsync(c->rows.p, frame_row_host, sizeof(int32_t) * (size_t)pl->n_frames, cudaMemcpyHostToDevice, c->run));
    }
    rc = [&]() -> int {
    int rc = B200WM_OK;
    int k = 0;
    for (int f0 = 0; f0 < pl->n_frames; f0 += chunk, ++k) {
        const int m = pl->n_frames - f0 < chunk ? pl->n_frames - f0 : chunk;
        const int s = k % kSlots;
        uint8_t* buf = (uint8_t*)c->buf[s].p;
        uint32_t* raw = (uint32_t*)c->raw[s].p;
        int32_t* cnt = (int32_t*)c->cnt.p + (size_t)f0 * payload_len;
        uint8_t* pat = (uint8_t*)c->pat.p + (size_t)f0 * payload_len;
        b200wm_plane host = *pl, devp = *pl;
        host.n_frames = devp.n_frames = m;
        devp.pitch_bytes = pl->width;
        devp.frame_stride_bytes = (long long)plane;
        if (k >= kSlots) B200WM_CUDA_TRY(cudaStreamWaitEvent(c->up, c->drained[s], 0));     // the slot's previous chunk is back on the host
        if ((rc = copy_planes(buf, src + (size_t)f0 * pl->frame_stride_bytes, &host, m, true, c->up))) return rc;
        B200WM_CUDA_TRY(cudaEventRecord(c->uploaded[s], c->up));
        B200WM_CUDA_TRY(cudaStreamWaitEvent(c->run, c->uploaded[s], 0));
        if ((rc = launch_dwtsvd_embed(buf, buf, &devp, (const uint32_t*)c->wm.p, n_rows, wm_words, wm_len,
                                      frame_row_host ? (const int32_t*)c->rows.p + f0 : nullptr, scale, c->run)))
            return rc;
        if ((rc = launch_dwtsvd_extract(buf, &devp, scale, raw, g.words, payload_len, cnt, nullptr, c->run))) return rc;
        if ((rc = launch_vote_finish(cnt, m, payload_len, g.block_num, (const int32_t*)c->perm.p, pat, nullptr, c->run))) return rc;
        B200WM_CUDA_TRY(cudaEventRecord(c->computed[s], c->run));
        B200WM_CUDA_TRY(cudaStreamWaitEvent(c->down, c->computed[s], 0));
        if ((rc = copy_planes(dst + (size_t)f0 * pl->frame_stride_bytes, buf, &host, m, false, c->down))) return rc;
        if (raw_bits_host && g.words)
            B200WM_CUDA_TRY(cudaMemcpyAsync(raw_bits_host + (size_t)f0 * g.words, raw, sizeof(uint32_t) * (size_t)m * g.words,
                                            cudaMemcpyDeviceToHost, c->down));
        B200WM_CUDA_TRY(cudaEventRecord(c->drained[s], c->down));
    }
    B200WM_CUDA_TRY(cudaMemcpyAsync(patterns_host, c->pat.p, (size_t)pl->n_frames * payload_len, cudaMemcpyDeviceToHost, c->run));
    if (pos_counts_host)
        B200WM_CUDA_TRY(cudaMemcpyAsync(pos_counts_host, c->cnt.p, sizeof(int32_t) * (size_t)pl->n_frames * payload_len,
                                        cudaMemcpyDeviceToHost, c->run));
    return B200WM_OK;
    }();
    const int rc_sync = c->sync_all();
    return rc ? rc : rc_sync;
}

}  // namespace b200wm
