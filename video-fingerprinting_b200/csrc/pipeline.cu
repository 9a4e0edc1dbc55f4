// Host-buffer entry points: the caller's frames live in HOST memory (the situation of the
// reference's drivers, which get frames from an ffmpeg pipe: src/offmark/video/frame_reader.py:53-64
// and hand them back to one: frame_writer.py:41-44).  The batch is streamed through the GPU in
// chunks on two streams so that the upload of chunk i+1, the kernels of chunk i and the download of
// chunk i-1 overlap (PCIe is full duplex); device scratch is allocated once per call.
// Pinned host memory gives full PCIe speed; pageable memory works but is staged by the driver.
#include "common.cuh"

namespace b200wm {

int validate_plane(const b200wm_plane* pl);
int launch_dwtsvd_embed(const void*, void*, const b200wm_plane*, const uint32_t*, int, long long, const int32_t*, float,
                        cudaStream_t);
int launch_dwtsvd_extract(const void*, const b200wm_plane*, float, uint32_t*, int, int, int32_t*, float*, cudaStream_t);
int launch_vote_finish(const int32_t*, int, int, long long, const int32_t*, uint8_t*, uint64_t*, cudaStream_t);

namespace {

struct Scratch {
    void* ptrs[16];
    int n = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    ~Scratch() {
        for (int i = 0; i < 2; ++i)
            if (streams[i]) { cudaStreamSynchronize(streams[i]); cudaStreamDestroy(streams[i]); }
        for (int i = 0; i < n; ++i) cudaFree(ptrs[i]);
    }
    template <typename T>
    int alloc(T** p, size_t bytes) {
        void* q = nullptr;
        B200WM_CUDA_TRY(cudaMalloc(&q, bytes ? bytes : 1));
        ptrs[n++] = q;
        *p = (T*)q;
        return B200WM_OK;
    }
};

// copy `frames` planes between a strided host layout and a dense [frames, h, w] device chunk
int copy_planes(void* dst, const void* src, const b200wm_plane* pl, int frames, bool to_device, cudaStream_t s) {
    const size_t row = (size_t)pl->width, plane = row * pl->height;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (pl->pitch_bytes == pl->width && (frames == 1 || pl->frame_stride_bytes == (long long)plane)) {
        // the whole chunk is one contiguous run on both sides
        if (to_device) B200WM_CUDA_TRY(cudaMemcpyAsync(dst, src, plane * frames, kind, s));
        else B200WM_CUDA_TRY(cudaMemcpyAsync(dst, src, plane * frames, kind, s));
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
    long long c = (256ll << 20) / (plane > 0 ? plane : 1);       // ~256 MB per chunk
    if (c < 1) c = 1;
    if (c > pl->n_frames) c = pl->n_frames > 0 ? pl->n_frames : 1;
    return (int)c;
}

}  // namespace

int mark_host(const uint8_t* src, uint8_t* dst, const b200wm_plane* pl, const uint32_t* wm_host, int n_rows, int wm_words,
              long long wm_len, const int32_t* frame_row_host, float scale, int chunk_frames) {
    int rc = check_host_plane(pl);
    if (rc) return rc;
    if (!src || !dst || !wm_host || n_rows <= 0 || wm_words <= 0) return B200WM_ERR_INVALID;
    if (pl->n_frames == 0) return B200WM_OK;
    const int chunk = pick_chunk(chunk_frames, pl);
    const size_t plane = (size_t)pl->width * pl->height;
    Scratch sc;
    uint8_t* buf[2];
    uint32_t* wm = nullptr;
    int32_t* rows = nullptr;
    for (int i = 0; i < 2; ++i) {
        if ((rc = sc.alloc(&buf[i], plane * chunk))) return rc;
        B200WM_CUDA_TRY(cudaStreamCreateWithFlags(&sc.streams[i], cudaStreamNonBlocking));
    }
    if ((rc = sc.alloc(&wm, sizeof(uint32_t) * (size_t)n_rows * wm_words))) return rc;
    B200WM_CUDA_TRY(cudaMemcpy(wm, wm_host, sizeof(uint32_t) * (size_t)n_rows * wm_words, cudaMemcpyHostToDevice));
    if (frame_row_host) {
        if ((rc = sc.alloc(&rows, sizeof(int32_t) * (size_t)pl->n_frames))) return rc;
        B200WM_CUDA_TRY(cudaMemcpy(rows, frame_row_host, sizeof(int32_t) * (size_t)pl->n_frames, cudaMemcpyHostToDevice));
    }
    int i = 0;
    for (int f0 = 0; f0 < pl->n_frames; f0 += chunk, i ^= 1) {
        const int m = pl->n_frames - f0 < chunk ? pl->n_frames - f0 : chunk;
        cudaStream_t s = sc.streams[i];
        b200wm_plane host = *pl, devp = *pl;
        host.n_frames = devp.n_frames = m;
        devp.pitch_bytes = pl->width;
        devp.frame_stride_bytes = (long long)plane;
        if ((rc = copy_planes(buf[i], src + (size_t)f0 * pl->frame_stride_bytes, &host, m, true, s))) return rc;
        if ((rc = launch_dwtsvd_embed(buf[i], buf[i], &devp, wm, wm_words, wm_len, rows ? rows + f0 : nullptr, scale, s))) return rc;
        if ((rc = copy_planes(dst + (size_t)f0 * pl->frame_stride_bytes, buf[i], &host, m, false, s))) return rc;
    }
    for (int k = 0; k < 2; ++k) B200WM_CUDA_TRY(cudaStreamSynchronize(sc.streams[k]));
    return B200WM_OK;
}

int detect_host(const uint8_t* src, const b200wm_plane* pl, float scale, int payload_len, const int32_t* perm_host,
                uint8_t* patterns_host, uint32_t* raw_bits_host, int32_t* pos_counts_host, int chunk_frames) {
    int rc = check_host_plane(pl);
    if (rc) return rc;
    if (!src || !perm_host || !patterns_host || payload_len <= 0) return B200WM_ERR_INVALID;
    if (pl->n_frames == 0) return B200WM_OK;
    const int chunk = pick_chunk(chunk_frames, pl);
    const size_t plane = (size_t)pl->width * pl->height;
    const TileGeom g = make_geom(pl->height, pl->width);
    Scratch sc;
    uint8_t* buf[2];
    uint32_t* raw[2];
    int32_t* cnt[2];
    uint8_t* pat[2];
    int32_t* perm = nullptr;
    for (int i = 0; i < 2; ++i) {
        if ((rc = sc.alloc(&buf[i], plane * chunk))) return rc;
        if ((rc = sc.alloc(&raw[i], sizeof(uint32_t) * (size_t)chunk * (g.words ? g.words : 1)))) return rc;
        if ((rc = sc.alloc(&cnt[i], sizeof(int32_t) * (size_t)chunk * payload_len))) return rc;
        if ((rc = sc.alloc(&pat[i], (size_t)chunk * payload_len))) return rc;
        B200WM_CUDA_TRY(cudaStreamCreateWithFlags(&sc.streams[i], cudaStreamNonBlocking));
    }
    if ((rc = sc.alloc(&perm, sizeof(int32_t) * (size_t)payload_len))) return rc;
    B200WM_CUDA_TRY(cudaMemcpy(perm, perm_host, sizeof(int32_t) * (size_t)payload_len, cudaMemcpyHostToDevice));
    int i = 0;
    for (int f0 = 0; f0 < pl->n_frames; f0 += chunk, i ^= 1) {
        const int m = pl->n_frames - f0 < chunk ? pl->n_frames - f0 : chunk;
        cudaStream_t s = sc.streams[i];
        b200wm_plane host = *pl, devp = *pl;
        host.n_frames = devp.n_frames = m;
        devp.pitch_bytes = pl->width;
        devp.frame_stride_bytes = (long long)plane;
        if ((rc = copy_planes(buf[i], src + (size_t)f0 * pl->frame_stride_bytes, &host, m, true, s))) return rc;
        if ((rc = launch_dwtsvd_extract(buf[i], &devp, scale, raw[i], g.words, payload_len, cnt[i], nullptr, s))) return rc;
        if ((rc = launch_vote_finish(cnt[i], m, payload_len, g.block_num, perm, pat[i], nullptr, s))) return rc;
        B200WM_CUDA_TRY(cudaMemcpyAsync(patterns_host + (size_t)f0 * payload_len, pat[i], (size_t)m * payload_len,
                                        cudaMemcpyDeviceToHost, s));
        if (raw_bits_host && g.words)
            B200WM_CUDA_TRY(cudaMemcpyAsync(raw_bits_host + (size_t)f0 * g.words, raw[i], sizeof(uint32_t) * (size_t)m * g.words,
                                            cudaMemcpyDeviceToHost, s));
        if (pos_counts_host)
            B200WM_CUDA_TRY(cudaMemcpyAsync(pos_counts_host + (size_t)f0 * payload_len, cnt[i],
                                            sizeof(int32_t) * (size_t)m * payload_len, cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < 2; ++k) B200WM_CUDA_TRY(cudaStreamSynchronize(sc.streams[k]));
    return B200WM_OK;
}

}  // namespace b200wm
