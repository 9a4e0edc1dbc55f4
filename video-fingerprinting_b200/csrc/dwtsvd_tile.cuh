// Per-tile pieces shared by the LDG kernels (dwtsvd.cu) and the TMA kernels (dwtsvd_tma.cu).
#pragma once
#include "common.cuh"
#include "svd4.cuh"
#include "svd4x2.cuh"

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
    int n_rows;               // rows of wm; row indices are clamped into [0, n_rows) (never an out-of-bounds read)
    int wm_words;
    float scale, inv_scale;
};

struct ExtractArgs {
    uint32_t* raw_bits;       // [n_frames, words]
    int32_t* pos_counts;      // nullable, [n_frames, payload_len]; only when 32 % payload_len == 0
    float* sigma;             // nullable debug output [n_frames, n_tiles]
    int payload_len;
    unsigned every;           // bit i*payload_len set for every i (payload_len divides 32), else 0
    float scale, inv_scale;
};

__device__ __forceinline__ int clamp_row(int row, int n_rows) { return min(max(row, 0), n_rows - 1); }

inline unsigned every_mask(int payload_len) {
    if (payload_len <= 0 || payload_len > 32 || 32 % payload_len) return 0u;
    return payload_len == 32 ? 1u : (0xFFFFFFFFu / ((1u << payload_len) - 1u));
}

// ------------------------------------------------------------------------------------------
// flat tiles on a quantisation boundary
// ------------------------------------------------------------------------------------------
// The kernels take sigma_0 from the exact 2x2 sums; the reference takes it from a float32 Haar band
// whose taps are float32(1/sqrt(2)) (PyWavelets; oracle/haar.py), i.e. from LL values that are a few
// 1e-8 (relative) below the exact ones.  That only matters where sigma_0 sits on a quantisation boundary
// (k*scale for the floor, (k+1/2)*scale for the bit) - and there the reference is still deterministic
// for FLAT tiles (all 64 samples equal v): LL = 4*fl(c*fl(c*v)) in every coefficient, cv2.dct and LAPACK
// are exact on a constant block, so sigma_0 = 16*fl(c*fl(c*|v|)) (checked against the oracle for every
// uint8 v and random float v in tests/test_oracle_golden.py).  E.g. a flat 255 tile has sigma_0 =
// 2039.99988 (bit 1, floor 135), not 2040 (bit 0, floor 136); black 0/16 and white 235 are no boundary cases.
// So: when the eigen-iteration reports a possibly flat block (one compare, svd4.cuh) whose sigma_0 lands
// within 2^-21 (relative) of a boundary, the kernel probes whether the tile is flat and, if so, takes the
// reference's value.  Non-flat tiles on a boundary stay what they are:
// decided by float32 rounding inside cv2.dct / sgesdd, which no independent implementation reproduces.
constexpr float kHaarTap = 0.70710677f;          // float32(1/sqrt(2))
__device__ __forceinline__ float flat_sigma_ref(float sample) {
    return 16.0f * __fmul_rn(kHaarTap, __fmul_rn(kHaarTap, fabsf(sample)));
}
template <bool kBitThreshold>                    // extract: floor and bit boundaries; embed: the floor only
__device__ __forceinline__ bool on_boundary(float sigma, float rem, float scale) {
    const float tol = sigma * 4.7683716e-7f;     // 2^-21 relative (0 for an all-zero block: never taken)
    bool b = (rem < tol) | (scale - rem < tol);
    if (kBitThreshold) b |= fabsf(rem - 0.5f * scale) < tol;
    return b;
}
// Probes: the common sample value of a flat tile, NaN otherwise.  Out of line: they run once in a few
// thousand tiles and must not cost the hot path registers.
template <typename T>
static __device__ __noinline__ float flat_probe_global(const uint8_t* p, unsigned pitch, int es) {
    const float v0 = (float)*reinterpret_cast<const T*>(p);
    bool flat = true;
#pragma unroll 1
    for (int y = 0; y < 8; ++y) {
        const T* r = reinterpret_cast<const T*>(p + (unsigned long long)y * pitch);
#pragma unroll 1
        for (int x = 0; x < 8; ++x) flat &= ((float)r[x * es] == v0);
    }
    return flat ? v0 : __int_as_float(0x7FC00000);
}
static __device__ __noinline__ float flat_probe_shared(unsigned addr, unsigned pitch) {
    unsigned w0;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(addr));
    const unsigned want = __byte_perm(w0, 0u, 0x0000);
    bool flat = true;
#pragma unroll 1
    for (int y = 0; y < 8; ++y) {
        unsigned a, b;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr + y * pitch));
        flat &= (a == want) & (b == want);
    }
    return flat ? (float)(w0 & 0xFFu) : __int_as_float(0x7FC00000);
}
struct NoProbe {     // callers whose input cannot be probed cheaply keep the exact-sum value
    __device__ __forceinline__ float operator()() const { return __int_as_float(0x7FC00000); }
};

// ------------------------------------------------------------------------------------------
// tile loaders: produce S[16] (2x2 sums, row-major over the 4x4 block)
// ------------------------------------------------------------------------------------------
// Fast path: planar uint8, 8-byte aligned rows.  rows[r] keeps the raw bytes for the embed.
// S[4*i+j] = sum of the 2x2 samples (rows 2i,2i+1; columns 2j,2j+1) of an 8x8 uint8 tile.
__device__ __forceinline__ void sums_from_rows(const uint2 (&rows)[8], float (&S)[16]) {
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

template <bool kReadOnly>
__device__ __forceinline__ void load_tile_u8(const uint8_t* p, unsigned pitch, uint2 (&rows)[8], float (&S)[16]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) rows[r] = kReadOnly ? ldg_nc_u2(row_ptr(p, r, pitch)) : ldg_stream_u2(row_ptr(p, r, pitch));
    sums_from_rows(rows, S);
}

// One row of a tile (8 uint8 samples in a uint2) plus the integer increments of its four 2x2
// columns, saturated to [0, 255].  d01 / d23 hold two increments each as int16 lanes; the bytes
// are widened to int16 lanes (even / odd samples), added and clamped with one DPX instruction
// per pair (VIADDMNMX.S16x2.RELU) and narrowed back.
__device__ __forceinline__ uint2 add_clamp_row(uint2 w, unsigned d01, unsigned d23) {
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
    return out;
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
// kDeep: straight-line squarings of the eigen-iteration (svd4.cuh): 0 for uint8 luma planes, 3 for float / chroma input
template <bool kStash, int kDeep = 0, typename Probe>
__device__ __forceinline__ void embed_deltas(float (&S)[16], int bit, float scale, float inv_scale,
                                             float bias, float (&D)[16], float4* stash, Probe probe) {
    float v[4];
    bool zero, maybe_flat;
    if (kStash) {
#pragma unroll
        for (int i = 0; i < 4; ++i)   // asm: the compiler must not forward these stores to the loads below
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(stash + i * kThreads)),
                         "f"(S[4 * i]), "f"(S[4 * i + 1]), "f"(S[4 * i + 2]), "f"(S[4 * i + 3]) : "memory");
    }
    float sigma = 0.5f * top_singular<true, kDeep>(S, v, zero, maybe_flat);   // sigma_0 of the LL block
    float q, rem;
    floor_divmod(sigma, scale, inv_scale, q, rem);
    if (maybe_flat && on_boundary<false>(sigma, rem, scale)) {
        const float flat = probe();
        if (flat == flat) {
            sigma = flat_sigma_ref(flat);
            floor_divmod(sigma, scale, inv_scale, q, rem);
        }
    }
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

// The bit-independent half of the quantisation, for callers that embed SEVERAL payloads into the same
// block (one read, N marked copies): sigma_0, the floor quotient, v0 and S v0.  embed_copy_deltas()
// then repeats, bit by bit, exactly the arithmetic of embed_deltas() above, so copy c of a multi-copy
// embed is bit-identical to a single embed of payload c.
struct BlockPair {
    float sigma, q;      // sigma_0 of the LL block and floor(sigma_0 / scale)
    float v[4], sv[4];   // right singular vector and S v (per LL row)
    bool zero;
};

template <typename Probe>
__device__ __forceinline__ void embed_prepare(const float (&S)[16], float scale, float inv_scale, BlockPair& bp, Probe probe) {
    bool maybe_flat;
    bp.sigma = 0.5f * top_singular<true>(S, bp.v, bp.zero, maybe_flat);
    float rem;
    floor_divmod(bp.sigma, scale, inv_scale, bp.q, rem);
    if (maybe_flat && on_boundary<false>(bp.sigma, rem, scale)) {
        const float flat = probe();
        if (flat == flat) {
            bp.sigma = flat_sigma_ref(flat);
            floor_divmod(bp.sigma, scale, inv_scale, bp.q, rem);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        bp.sv[i] = fmaf(S[4 * i + 3], bp.v[3], fmaf(S[4 * i + 2], bp.v[2], fmaf(S[4 * i + 1], bp.v[1], S[4 * i] * bp.v[0])));
}

__device__ __forceinline__ void embed_copy_deltas(const BlockPair& bp, int bit, float scale, float bias, float (&D)[16]) {
    const float target = (bp.q + 0.25f + 0.5f * (float)bit) * scale;
    if (bp.zero) {
#pragma unroll
        for (int k = 0; k < 16; ++k) D[k] = fmaf(target, 0.125f, bias);
        return;
    }
    const float t = 0.25f * ((target - bp.sigma) / bp.sigma);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = t * bp.sv[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) D[4 * i + j] = fmaf(a, bp.v[j], bias);
    }
}

template <int kDeep = 0, typename Probe>
__device__ __forceinline__ int extract_bit(const float (&S)[16], float scale, float inv_scale, float& sigma, Probe probe) {
    float v[4];
    bool zero, maybe_flat;
    sigma = 0.5f * top_singular<false, kDeep>(S, v, zero, maybe_flat);
    float q, rem;
    floor_divmod(sigma, scale, inv_scale, q, rem);
    if (maybe_flat && on_boundary<true>(sigma, rem, scale)) {
        const float flat = probe();
        if (flat == flat) {
            sigma = flat_sigma_ref(flat);
            floor_divmod(sigma, scale, inv_scale, q, rem);
        }
    }
    return rem > 0.5f * scale ? 1 : 0;
}


// ------------------------------------------------------------------------------------------
// two tiles per thread, packed FP32 (svd4x2.cuh); lane .x = first tile, lane .y = second tile
// ------------------------------------------------------------------------------------------
// The four bytes of a 2x2 are first gathered into one register (PRMT, integer pipe) and then summed
// into the magic float pattern by ONE dp4a: IDP issues at half rate on the FMA-heavy pipe, which is the
// co-critical resource of these kernels next to the issue slots (profiles/r01_summary.md).
__device__ __forceinline__ void sums_from_rows_x2(const uint2 (&ra)[8], const uint2 (&rb)[8], f2 (&S)[16]) {
    constexpr unsigned kMagic = 0x4B000000u, kOnes = 0x01010101u;
    const f2 minus_magic = bc2(-8388608.0f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned sa[4], sb[4];
        {
            const uint2 a = ra[2 * i], b = ra[2 * i + 1];
            sa[0] = __dp4a(__byte_perm(a.x, b.x, 0x5410), kOnes, kMagic);
            sa[1] = __dp4a(__byte_perm(a.x, b.x, 0x7632), kOnes, kMagic);
            sa[2] = __dp4a(__byte_perm(a.y, b.y, 0x5410), kOnes, kMagic);
            sa[3] = __dp4a(__byte_perm(a.y, b.y, 0x7632), kOnes, kMagic);
        }
        {
            const uint2 a = rb[2 * i], b = rb[2 * i + 1];
            sb[0] = __dp4a(__byte_perm(a.x, b.x, 0x5410), kOnes, kMagic);
            sb[1] = __dp4a(__byte_perm(a.x, b.x, 0x7632), kOnes, kMagic);
            sb[2] = __dp4a(__byte_perm(a.y, b.y, 0x5410), kOnes, kMagic);
            sb[3] = __dp4a(__byte_perm(a.y, b.y, 0x7632), kOnes, kMagic);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            S[4 * i + j] = add2(make_float2(__uint_as_float(sa[j]), __uint_as_float(sb[j])), minus_magic);
    }
}

// extract_bit() for both lanes: bit 0 = first tile, bit 1 = second tile
// probe(0) / probe(1): flat probe of the first / second tile
template <typename Probe>
__device__ __forceinline__ unsigned extract_bits_x2(const f2 (&S)[16], float scale, float inv_scale, Probe probe) {
    f2 v[4];
    const Pair2 p = top_singular_x2<false>(S, v);
    f2 sigma = mul2(bc2(0.5f), p.sigma0);
    f2 q, rem;
    floor_divmod2(sigma, scale, inv_scale, q, rem);
#ifndef B200WM_NO_FLAT_RULE      // (tuning aid: what the rule costs)
    if (p.flat_x | p.flat_y)
#else
    if (false)
#endif
    {
        const bool edge_x = p.flat_x && on_boundary<true>(sigma.x, rem.x, scale), edge_y = p.flat_y && on_boundary<true>(sigma.y, rem.y, scale);
        const float fx = edge_x ? probe(0) : __int_as_float(0x7FC00000), fy = edge_y ? probe(1) : __int_as_float(0x7FC00000);
        if (fx == fx) sigma.x = flat_sigma_ref(fx);
        if (fy == fy) sigma.y = flat_sigma_ref(fy);
        floor_divmod2(sigma, scale, inv_scale, q, rem);
    }
    const float half = 0.5f * scale;
    return (rem.x > half ? 1u : 0u) | (rem.y > half ? 2u : 0u);
}

// embed_deltas() for both lanes (bits: bit 0 = first tile, bit 1 = second tile)
template <typename Probe>
__device__ __forceinline__ void embed_deltas_x2(const f2 (&S)[16], unsigned bits, float scale, float inv_scale, float bias,
                                                f2 (&D)[16], Probe probe) {
    f2 v[4];
    const Pair2 p = top_singular_x2<true>(S, v);
    f2 sigma = mul2(bc2(0.5f), p.sigma0);
    f2 q, rem;
    floor_divmod2(sigma, scale, inv_scale, q, rem);
#ifndef B200WM_NO_FLAT_RULE
    if (p.flat_x | p.flat_y)
#else
    if (false)
#endif
    {
        const bool edge_x = p.flat_x && on_boundary<false>(sigma.x, rem.x, scale), edge_y = p.flat_y && on_boundary<false>(sigma.y, rem.y, scale);
        const float fx = edge_x ? probe(0) : __int_as_float(0x7FC00000), fy = edge_y ? probe(1) : __int_as_float(0x7FC00000);
        if (fx == fx) sigma.x = flat_sigma_ref(fx);
        if (fy == fy) sigma.y = flat_sigma_ref(fy);
        floor_divmod2(sigma, scale, inv_scale, q, rem);
    }
    const f2 bit = make_float2((float)(bits & 1u), (float)((bits >> 1) & 1u));
    const f2 target = mul2(add2(add2(q, bc2(0.25f)), mul2(bc2(0.5f), bit)), bc2(scale));
    const f2 diff = add2(target, neg2(sigma));
    const f2 t = mul2(bc2(0.25f), make_float2(diff.x / sigma.x, diff.y / sigma.y));
    const f2 bias2 = bc2(bias);
    // svd(0) = (I, 0, I): an all-zero block gets the flat increment target/8 on every sample
    // (fmaf(target, 0.125f, bias) in the scalar code).  target/8 is exact, so the same value comes out
    // of the common formula fma(a, v, bias) with a = target/8 and v = 1 on that lane.
    const f2 flat = mul2(target, bc2(0.125f));
    if (p.zero_x | p.zero_y) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = make_float2(p.zero_x ? 1.0f : v[j].x, p.zero_y ? 1.0f : v[j].y);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const f2 sv = fma2(S[4 * i + 3], v[3], fma2(S[4 * i + 2], v[2], fma2(S[4 * i + 1], v[1], mul2(S[4 * i], v[0]))));
        f2 a = mul2(t, sv);
        a = make_float2(p.zero_x ? flat.x : a.x, p.zero_y ? flat.y : a.y);
#pragma unroll
        for (int j = 0; j < 4; ++j) D[4 * i + j] = fma2(a, v[j], bias2);
    }
}

}  // namespace b200wm
