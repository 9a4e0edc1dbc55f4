// Shared helpers for the b200wm kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200wm.h"

namespace b200wm {

// ---- status plumbing --------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* where);
void count_launch(int n = 1);

#define B200WM_CUDA_TRY(expr)                                              \
    do {                                                                   \
        cudaError_t e__ = (expr);                                          \
        if (e__ != cudaSuccess) {                                          \
            ::b200wm::set_cuda_error(e__, #expr);                          \
            return B200WM_ERR_CUDA;                                        \
        }                                                                  \
    } while (0)

#define B200WM_LAUNCH_CHECK(name)                                          \
    do {                                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) {                                          \
            ::b200wm::set_cuda_error(e__, name);                           \
            return B200WM_ERR_CUDA;                                        \
        }                                                                  \
        ::b200wm::count_launch();                                          \
    } while (0)

// Stream-ordered scratch (cudaMallocAsync) is used for per-call tables and mask arrays.  By default the runtime's pool
// hands freed memory back to the driver at the next synchronisation, which makes every call pay for a fresh allocation
// (measured: 1 ms per 130 MB); this keeps what the pool has, once per device.
int retain_async_pool();

// ---- geometry -----------------------------------------------------------------------
struct TileGeom {
    int tiles_x;         // 8x8-sample tiles per row that the reference walks
    int tiles_y;
    int n_tiles;         // tiles_x * tiles_y
    int words;           // uint32 words per frame of raw bits (covers block_num)
    long long block_num; // height*width/64
    unsigned long long div_magic;  // floor(2^40 / tiles_x) + 1: c / tiles_x == (c * magic) >> 40 for c < 2^26
};

inline TileGeom make_geom(int height, int width) {
    TileGeom g;
    g.tiles_x = ((width / 4 * 4) / 2) / 4;
    g.tiles_y = ((height / 4 * 4) / 2) / 4;
    g.n_tiles = g.tiles_x * g.tiles_y;
    g.block_num = (long long)height * width / 64;
    g.words = (int)((g.block_num + 31) / 32);
    g.div_magic = g.tiles_x > 0 ? ((1ull << 40) / (unsigned long long)g.tiles_x + 1ull) : 0ull;
    return g;
}

// ---- device helpers -----------------------------------------------------------------
__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_nc_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u2(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// a / b and sqrt(a) for positive normal operands: hardware approximation plus one Newton step.
// Within 1 ulp (almost always correctly rounded) and, unlike the IEEE routines, straight-line
// code with no slow-path branch.
__device__ __forceinline__ float div_pos(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    const float q = a * r;
    return fmaf(fmaf(-q, b, a), r, q);
}
__device__ __forceinline__ float sqrt_pos(float a) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    const float s = a * r;
    return fmaf(fmaf(-s, s, a) * 0.5f, r, s);
}

// exact float32 remainder of a non-negative x by a positive m, and the floor quotient
// (numpy's float `%` and `//` on positive operands: embed/dwt_dct_svd_encoder.py:44,
// extract/dwt_dct_svd_decoder.py:36).  Valid while x/m < 2^23.  Branch-free: the quotient
// estimate is off by at most one, fixed with selects, and the remainder of the TRUE floor is
// exactly representable, so the final FMA is exact.
__device__ __forceinline__ void floor_divmod(float x, float m, float inv_m, float& q, float& r) {
    q = floorf(x * inv_m);
    r = fmaf(-m, q, x);
    const float fix = r < 0.0f ? -1.0f : (r >= m ? 1.0f : 0.0f);
    q += fix;
    r = fmaf(-m, q, x);
}

}  // namespace b200wm
