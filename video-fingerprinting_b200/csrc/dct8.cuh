// Register-resident 8-point DCT-II / DCT-III butterflies (orthonormal, like cv2.dct / cv2.idct).
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace b200wm {

// cos(k*pi/16), and the orthonormal scale factors
#define C1 0.98078528040323043f
#define C2 0.92387953251128674f
#define C3 0.83146961230254524f
#define C4 0.70710678118654752f
#define C5 0.55557023301960218f
#define C6 0.38268343236508977f
#define C7 0.19509032201612825f

// Orthonormal 8-point DCT-II in place: X_k = a_k sum_n x_n cos((2n+1) k pi / 16), a_0 = 1/sqrt(8), a_k = 1/2.
// kBias8: the inputs are integers that all carry the same additive bias B (bytes dropped into a float's mantissa by
// PRMT, no conversion instruction) with kBias8 = 8 * B and every partial sum still exact in float32.  Every term
// but the DC one is built from differences, in which B cancels exactly; the DC term removes 8 * B from the exact
// integer sum before its only rounding - the outputs are bit-identical to the unbiased transform's.
template <int kBias8 = 0>
__device__ __forceinline__ void dct8_1d(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6,
                                        float& x7) {
    const float s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;
    const float d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;
    const float e0 = s0 + s3, e1 = s1 + s2, e2 = s1 - s2, e3 = s0 - s3;
    x0 = (kBias8 ? (e0 + e1) - (float)kBias8 : (e0 + e1)) * (0.5f * C4);
    x4 = (e0 - e1) * (0.5f * C4);
    x2 = fmaf(e3, 0.5f * C2, e2 * (0.5f * C6));
    x6 = fmaf(e3, 0.5f * C6, e2 * (-0.5f * C2));
    x1 = fmaf(d3, 0.5f * C7, fmaf(d2, 0.5f * C5, fmaf(d1, 0.5f * C3, d0 * (0.5f * C1))));
    x3 = fmaf(d3, -0.5f * C5, fmaf(d2, -0.5f * C1, fmaf(d1, -0.5f * C7, d0 * (0.5f * C3))));
    x5 = fmaf(d3, 0.5f * C3, fmaf(d2, 0.5f * C7, fmaf(d1, -0.5f * C1, d0 * (0.5f * C5))));
    x7 = fmaf(d3, -0.5f * C1, fmaf(d2, 0.5f * C3, fmaf(d1, -0.5f * C5, d0 * (0.5f * C7))));
}

// Two rows per instruction: dct8_1d() with the two lanes of packed FP32 registers holding the same sample position of
// two different rows (FFMA2 / FMUL2 / FADD2, sm_100): per lane exactly the scalar operations in the scalar order, so the
// outputs are bit-identical to two dct8_1d() calls, for half the issue slots.
__device__ __forceinline__ float2 p2_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 p2_sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }   // b * -1 + a == a - b exactly
__device__ __forceinline__ float2 p2_mul(float2 a, float k) { return __fmul2_rn(a, make_float2(k, k)); }
__device__ __forceinline__ float2 p2_fma(float2 a, float k, float2 c) { return __ffma2_rn(a, make_float2(k, k), c); }

template <int kBias8 = 0>
__device__ __forceinline__ void dct8_1d_x2(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4, float2& x5, float2& x6,
                                           float2& x7) {
    const float2 s0 = p2_add(x0, x7), s1 = p2_add(x1, x6), s2 = p2_add(x2, x5), s3 = p2_add(x3, x4);
    const float2 d0 = p2_sub(x0, x7), d1 = p2_sub(x1, x6), d2 = p2_sub(x2, x5), d3 = p2_sub(x3, x4);
    const float2 e0 = p2_add(s0, s3), e1 = p2_add(s1, s2), e2 = p2_sub(s1, s2), e3 = p2_sub(s0, s3);
    const float2 dc = p2_add(e0, e1);
    x0 = p2_mul(kBias8 ? p2_add(dc, make_float2(-(float)kBias8, -(float)kBias8)) : dc, 0.5f * C4);
    x4 = p2_mul(p2_sub(e0, e1), 0.5f * C4);
    x2 = p2_fma(e3, 0.5f * C2, p2_mul(e2, 0.5f * C6));
    x6 = p2_fma(e3, 0.5f * C6, p2_mul(e2, -0.5f * C2));
    x1 = p2_fma(d3, 0.5f * C7, p2_fma(d2, 0.5f * C5, p2_fma(d1, 0.5f * C3, p2_mul(d0, 0.5f * C1))));
    x3 = p2_fma(d3, -0.5f * C5, p2_fma(d2, -0.5f * C1, p2_fma(d1, -0.5f * C7, p2_mul(d0, 0.5f * C3))));
    x5 = p2_fma(d3, 0.5f * C3, p2_fma(d2, 0.5f * C7, p2_fma(d1, -0.5f * C1, p2_mul(d0, 0.5f * C5))));
    x7 = p2_fma(d3, -0.5f * C1, p2_fma(d2, 0.5f * C3, p2_fma(d1, -0.5f * C5, p2_mul(d0, 0.5f * C7))));
}

// 2-D transform with the row pass packed two rows at a time and the column pass scalar: bit-identical to dct8x8()
template <int kBias8 = 0>
__device__ __forceinline__ void dct8x8_rows_x2(float (&b)[64]) {
#pragma unroll
    for (int yp = 0; yp < 4; ++yp) {
        float2 r[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) r[x] = make_float2(b[16 * yp + x], b[16 * yp + 8 + x]);
        dct8_1d_x2<kBias8>(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
#pragma unroll
        for (int x = 0; x < 8; ++x) { b[16 * yp + x] = r[x].x; b[16 * yp + 8 + x] = r[x].y; }
    }
#pragma unroll
    for (int x = 0; x < 8; ++x) dct8_1d(b[x], b[8 + x], b[16 + x], b[24 + x], b[32 + x], b[40 + x], b[48 + x], b[56 + x]);
}

// Inverse of dct8_1d (orthonormal DCT-III): x_n = sum_k a_k X_k cos((2n+1) k pi / 16).
// Even coefficients give a part symmetric in n <-> 7-n, odd coefficients an antisymmetric part.
__device__ __forceinline__ void idct8_1d(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6,
                                         float& x7) {
    const float a = (x0 + x4) * (0.5f * C4), b = (x0 - x4) * (0.5f * C4);
    const float c = fmaf(x2, 0.5f * C2, x6 * (0.5f * C6)), d = fmaf(x2, 0.5f * C6, x6 * (-0.5f * C2));
    const float e0 = a + c, e1 = b + d, e2 = b - d, e3 = a - c;
    const float o0 = fmaf(x7, 0.5f * C7, fmaf(x5, 0.5f * C5, fmaf(x3, 0.5f * C3, x1 * (0.5f * C1))));
    const float o1 = fmaf(x7, -0.5f * C5, fmaf(x5, -0.5f * C1, fmaf(x3, -0.5f * C7, x1 * (0.5f * C3))));
    const float o2 = fmaf(x7, 0.5f * C3, fmaf(x5, 0.5f * C7, fmaf(x3, -0.5f * C1, x1 * (0.5f * C5))));
    const float o3 = fmaf(x7, -0.5f * C1, fmaf(x5, 0.5f * C3, fmaf(x3, -0.5f * C5, x1 * (0.5f * C7))));
    x0 = e0 + o0; x7 = e0 - o0;
    x1 = e1 + o1; x6 = e1 - o1;
    x2 = e2 + o2; x5 = e2 - o2;
    x3 = e3 + o3; x4 = e3 - o3;
}

// 2-D transforms of an 8x8 block held in registers (row-major b[8*y+x]).
template <int kBias8 = 0>
__device__ __forceinline__ void dct8x8(float (&b)[64]) {
#pragma unroll
    for (int y = 0; y < 8; ++y)
        dct8_1d<kBias8>(b[8 * y], b[8 * y + 1], b[8 * y + 2], b[8 * y + 3], b[8 * y + 4], b[8 * y + 5], b[8 * y + 6], b[8 * y + 7]);
#pragma unroll
    for (int x = 0; x < 8; ++x) dct8_1d(b[x], b[8 + x], b[16 + x], b[24 + x], b[32 + x], b[40 + x], b[48 + x], b[56 + x]);
}
__device__ __forceinline__ void idct8x8(float (&b)[64]) {
#pragma unroll
    for (int x = 0; x < 8; ++x) idct8_1d(b[x], b[8 + x], b[16 + x], b[24 + x], b[32 + x], b[40 + x], b[48 + x], b[56 + x]);
#pragma unroll
    for (int y = 0; y < 8; ++y)
        idct8_1d(b[8 * y], b[8 * y + 1], b[8 * y + 2], b[8 * y + 3], b[8 * y + 4], b[8 * y + 5], b[8 * y + 6], b[8 * y + 7]);
}

// Fast path for planar uint8 with 8-byte aligned rows: eight 64-bit loads per block, and every byte
// becomes a float by PRMT into the mantissa of 1.5 * 2^23 (value 12582912 + byte, exact) - no
// byte loads, no I2F.  Differences of two such values are the exact sample differences, and an FMA
// that adds a small increment to one rounds sample + increment to the nearest integer (ties to even)
// and leaves it in the low mantissa bits, ready for the 16-bit saturating pack below.
constexpr unsigned kBiasBits = 0x4B400000u;
constexpr float kBias = 12582912.0f;

__device__ __forceinline__ void load_rows_u8(const uint8_t* p, long long pitch, uint2 (&rows)[8]) {
#pragma unroll
    for (int y = 0; y < 8; ++y) rows[y] = ldg_stream_u2(p + y * pitch);
}
template <int kByte>
__device__ __forceinline__ float mid_biased_byte(unsigned w) {      // 32768 + byte: [0x47][0x00][byte][0x00]
    return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7404 | (kByte << 4)));
}
template <int kByte>
__device__ __forceinline__ float biased_byte(unsigned w) {
    return __uint_as_float(__byte_perm(w, kBiasBits, 0x7640 | kByte));
}
// b[8*y + x] = kBias + sample
__device__ __forceinline__ void biased_block(const uint2 (&rows)[8], float (&b)[64]) {
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        b[8 * y + 0] = biased_byte<0>(rows[y].x); b[8 * y + 1] = biased_byte<1>(rows[y].x);
        b[8 * y + 2] = biased_byte<2>(rows[y].x); b[8 * y + 3] = biased_byte<3>(rows[y].x);
        b[8 * y + 4] = biased_byte<0>(rows[y].y); b[8 * y + 5] = biased_byte<1>(rows[y].y);
        b[8 * y + 6] = biased_byte<2>(rows[y].y); b[8 * y + 7] = biased_byte<3>(rows[y].y);
    }
}
// eight floats holding kBias + integer -> eight bytes clamped to [0, 255]
__device__ __forceinline__ uint2 pack_biased_row(const float (&f)[8]) {
    unsigned lanes[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        lanes[k] = __viaddmin_s16x2_relu(__byte_perm(__float_as_uint(f[2 * k]), __float_as_uint(f[2 * k + 1]), 0x5410), 0u, 0x00FF00FFu);
    return make_uint2(__byte_perm(lanes[0], lanes[1], 0x6420), __byte_perm(lanes[2], lanes[3], 0x6420));
}


}  // namespace b200wm
