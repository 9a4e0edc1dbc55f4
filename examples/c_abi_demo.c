/*
 * The drop-in boundary from plain C: no Python, no torch.  Marks a synthetic 1080p luma plane with a
 * tiled 8-bit payload through b200wm_dwtsvd_embed, reads it back with b200wm_dwtsvd_extract +
 * b200wm_vote_finish and prints the recovered pattern.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/c_abi_demo.c -o /tmp/c_abi_demo \
 *       -Lvideo-fingerprinting_b200/lib -lb200wm -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/video-fingerprinting_b200/lib
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "b200wm.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_WM(x) do { int s_ = (x); if (s_ != B200WM_OK) { fprintf(stderr, "%s: %s (%s)\n", #x, b200wm_strerror(s_), b200wm_last_cuda_error()); return 3; } } while (0)

int main(void) {
    const int h = 1080, w = 1920, payload_len = 8;
    const int payload[8] = {0, 1, 1, 0, 0, 1, 0, 1};
    /* Shuffler(key=0) permutation of 8 positions (generator/shuffler.py:22) */
    const int32_t perm[8] = {6, 2, 1, 7, 3, 0, 5, 4};
    if (b200wm_device_ok() != B200WM_OK) { fprintf(stderr, "no sm_100 device\n"); return 1; }

    const int64_t block_num = b200wm_block_num(h, w);
    const int32_t words = b200wm_words_per_frame(h, w);
    /* wm[c] = shuffled_payload[c mod L]; np.random.shuffle applied to the payload: shuffled[i] = payload[perm[i]] */
    uint32_t* wm = (uint32_t*)calloc((size_t)words, sizeof(uint32_t));
    for (int64_t c = 0; c < block_num; ++c)
        if (payload[perm[c % payload_len]]) wm[c >> 5] |= 1u << (c & 31);

    uint8_t* plane = (uint8_t*)malloc((size_t)h * w);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) plane[(size_t)y * w + x] = (uint8_t)(96 + ((x * 7 + y * 13) % 64) + ((x / 16 + y / 16) % 2) * 20);

    uint8_t *d_plane, *d_patterns;
    uint32_t *d_wm, *d_raw;
    int32_t *d_counts, *d_perm;
    CHECK_CUDA(cudaMalloc((void**)&d_plane, (size_t)h * w));
    CHECK_CUDA(cudaMalloc((void**)&d_wm, sizeof(uint32_t) * words));
    CHECK_CUDA(cudaMalloc((void**)&d_raw, sizeof(uint32_t) * words));
    CHECK_CUDA(cudaMalloc((void**)&d_counts, sizeof(int32_t) * payload_len));
    CHECK_CUDA(cudaMalloc((void**)&d_perm, sizeof(int32_t) * payload_len));
    CHECK_CUDA(cudaMalloc((void**)&d_patterns, payload_len));
    CHECK_CUDA(cudaMemcpy(d_plane, plane, (size_t)h * w, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_wm, wm, sizeof(uint32_t) * words, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_perm, perm, sizeof(perm), cudaMemcpyHostToDevice));

    b200wm_plane pl;
    memset(&pl, 0, sizeof(pl));
    pl.dtype = B200WM_U8; pl.n_frames = 1; pl.height = h; pl.width = w;
    pl.pitch_bytes = w; pl.frame_stride_bytes = (int64_t)h * w; pl.elem_stride = 1;

    CHECK_WM(b200wm_dwtsvd_embed(d_plane, d_plane, &pl, d_wm, 1, words, block_num, NULL, 15.0f, NULL));
    CHECK_WM(b200wm_dwtsvd_extract(d_plane, &pl, 15.0f, d_raw, words, payload_len, d_counts, NULL));
    CHECK_WM(b200wm_vote_finish(d_counts, 1, payload_len, block_num, d_perm, d_patterns, NULL, NULL));
    CHECK_CUDA(cudaDeviceSynchronize());

    uint8_t pattern[8];
    CHECK_CUDA(cudaMemcpy(pattern, d_patterns, payload_len, cudaMemcpyDeviceToHost));
    int ok = 1;
    printf("recovered payload:");
    for (int i = 0; i < payload_len; ++i) { printf(" %d", pattern[i]); ok &= pattern[i] == payload[i]; }
    printf("  (%s, %d kernel launches, library version %d)\n", ok ? "matches" : "DIFFERS", b200wm_kernel_launches(), b200wm_version());
    return ok ? 0 : 4;
}
