// Two blocks per thread in packed FP32 (FFMA2 / FMUL2 / FADD2, new with sm_100).
//
// The eigen-iteration of svd4.cuh is issue-bound, not pipe-bound (ncu: ~80 % issue slots busy, FMA pipe
// under 50 %), and a packed instruction does two IEEE operations for ONE issue slot (measured with
// scripts/micro/ffma2_probe.cu: FFMA2 issues every second cycle, so the flop rate is unchanged but
// half the slots are freed for the integer, shared-memory and conversion work of the tile).  Lane .x of
// every float2 below belongs to the thread's first block, lane .y to its second: the two lanes never
// mix, every operation is the same correctly rounded operation, in the same order, as the scalar code
// in svd4.cuh - so results are bit-identical to the scalar kernels (tests compare the two families).
//
// Only the straight-line fast path is packed.  A lane whose certificate fails is redone from scratch by
// the scalar routine (which continues into the squaring loop and the Jacobi solver); luma planes
// essentially never take that branch.
#pragma once
#include "svd4.cuh"

namespace b200wm {

typedef float2 f2;
__device__ __forceinline__ f2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }

__device__ __forceinline__ void gram4x2(const f2 (&S)[16], f2 (&G)[10]) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            f2 acc = mul2(S[i], S[j]);
            acc = fma2(S[4 + i], S[4 + j], acc);
            acc = fma2(S[8 + i], S[8 + j], acc);
            acc = fma2(S[12 + i], S[12 + j], acc);
            G[k++] = acc;
        }
}

__device__ __forceinline__ void symv4x2(const f2 (&G)[10], const f2 (&x)[4], f2 (&y)[4]) {
    y[0] = fma2(G[3], x[3], fma2(G[2], x[2], fma2(G[1], x[1], mul2(G[0], x[0]))));
    y[1] = fma2(G[6], x[3], fma2(G[5], x[2], fma2(G[4], x[1], mul2(G[1], x[0]))));
    y[2] = fma2(G[8], x[3], fma2(G[7], x[2], fma2(G[5], x[1], mul2(G[2], x[0]))));
    y[3] = fma2(G[9], x[3], fma2(G[8], x[2], fma2(G[6], x[1], mul2(G[3], x[0]))));
}

__device__ __forceinline__ f2 dot4x2(const f2 (&a)[4], const f2 (&b)[4]) {
    return fma2(a[3], b[3], fma2(a[2], b[2], fma2(a[1], b[1], mul2(a[0], b[0]))));
}

__device__ __forceinline__ float rcp_approx(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float a) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
// div_pos / sqrt_pos of common.cuh, two lanes
__device__ __forceinline__ f2 div_pos2(f2 a, f2 b) {
    const f2 r = make_float2(rcp_approx(b.x), rcp_approx(b.y));
    const f2 q = mul2(a, r);
    return fma2(fma2(neg2(q), b, a), r, q);
}
__device__ __forceinline__ f2 sqrt_pos2(f2 a) {
    const f2 r = make_float2(rsqrt_approx(a.x), rsqrt_approx(a.y));
    const f2 s = mul2(a, r);
    return fma2(mul2(fma2(neg2(s), s, a), bc2(0.5f)), r, s);
}
// floor_divmod of common.cuh, two lanes
__device__ __forceinline__ void floor_divmod2(f2 x, float m, float inv_m, f2& q, f2& r) {
    const f2 p = mul2(x, bc2(inv_m));
    q = make_float2(floorf(p.x), floorf(p.y));
    const f2 nm = bc2(-m);
    r = fma2(nm, q, x);
    const f2 fix = make_float2(r.x < 0.0f ? -1.0f : (r.x >= m ? 1.0f : 0.0f), r.y < 0.0f ? -1.0f : (r.y >= m ? 1.0f : 0.0f));
    q = add2(q, fix);
    r = fma2(nm, q, x);
}

struct Pair2 {          // per-lane outputs of the packed routine
    f2 sigma0;          // largest singular value of S (not halved)
    bool zero_x, zero_y;
    bool flat_x, flat_y;   // maybe_flat of svd4.cuh, per lane
};

static __device__ __noinline__ Top5 top_singular_scalar(float s0, float s1, float s2, float s3, float s4, float s5, float s6, float s7,
                                                 float s8, float s9, float s10, float s11, float s12, float s13, float s14,
                                                 float s15) {
    const float S[16] = {s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15};
    float v[4];
    bool zero, flat;
    Top5 t;
    t.sigma = top_singular<true>(S, v, zero, flat);
    t.v0 = v[0]; t.v1 = v[1]; t.v2 = v[2]; t.v3 = v[3];
    return t;
}

// Packed counterpart of top_singular<kWantVec>() in svd4.cuh (same operations, same order, per lane).
template <bool kWantVec>
__device__ __forceinline__ Pair2 top_singular_x2(const f2 (&S)[16], f2 (&v)[4]) {
    f2 G[10];
    gram4x2(S, G);
    const f2 tr_raw = add2(add2(G[0], G[4]), add2(G[7], G[9]));
    Pair2 out;
    out.zero_x = !(tr_raw.x > 0.0f);
    out.zero_y = !(tr_raw.y > 0.0f);
    // The scalar routine first scales G by an exact power of two so that tr(G) lies in [0.5, 1).  That
    // scaling commutes with every operation below (powers of two, no rounding) and only exists to keep
    // arbitrary float input in range; the packed routine is used on uint8 planes only, where S <= 1020,
    // |G| < 2^23 and the largest intermediate (x.w) stays below 2^123, so it is skipped - same bits out.
    const f2 tr = tr_raw;

    f2 x[4], w[4];
    w[0] = add2(add2(G[0], G[1]), add2(G[2], G[3]));
    w[1] = add2(add2(G[1], G[4]), add2(G[5], G[6]));
    w[2] = add2(add2(G[2], G[5]), add2(G[7], G[8]));
    w[3] = add2(add2(G[3], G[6]), add2(G[8], G[9]));
    symv4x2(G, w, x);
#ifdef B200WM_FLAT_FROM_GRAM     // (tuning aid) the same filter from the first two Gram entries
    out.flat_x = G[0].x == G[1].x;
    out.flat_y = G[0].y == G[1].y;
#else
    out.flat_x = x[0].x == x[1].x;
    out.flat_y = x[0].y == x[1].y;
#endif
    symv4x2(G, x, w);
    // rayleigh_check, two lanes
    const f2 xw = dot4x2(x, w), xx = dot4x2(x, x);
    const f2 lam = mul2(make_float2(rcp_approx(xx.x), rcp_approx(xx.y)), xw);
    f2 r[4];
    const f2 nlam = neg2(lam);
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fma2(nlam, x[k], w[k]);
    const f2 rr = dot4x2(r, r);
    const f2 gap = fma2(bc2(2.0f), lam, neg2(tr));
    const f2 thr = mul2(mul2(mul2(bc2(kTau), lam), gap), xx);
    const bool done_x = ((gap.x > 0.0f) && (rr.x <= thr.x)) || out.zero_x;
    const bool done_y = ((gap.y > 0.0f) && (rr.y <= thr.y)) || out.zero_y;

    const f2 lam_full = div_pos2(xw, xx);
    const f2 root = sqrt_pos2(lam_full);
    out.sigma0 = make_float2(out.zero_x ? 0.0f : root.x, out.zero_y ? 0.0f : root.y);
    if (kWantVec) {
        // an all-zero block has x == 0 and n selected to 0, so the products are the zeros the scalar code selects
        const f2 n = make_float2(out.zero_x ? 0.0f : rsqrtf(xx.x), out.zero_y ? 0.0f : rsqrtf(xx.y));
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = mul2(x[k], n);
    }
    if (!(done_x && done_y)) {
        // rare: redo the failing lane(s) with the scalar routine (squaring loop, then Jacobi)
        if (!done_x) {
            const Top5 t = top_singular_scalar(S[0].x, S[1].x, S[2].x, S[3].x, S[4].x, S[5].x, S[6].x, S[7].x, S[8].x, S[9].x,
                                               S[10].x, S[11].x, S[12].x, S[13].x, S[14].x, S[15].x);
            out.sigma0.x = t.sigma;
            if (kWantVec) { v[0].x = t.v0; v[1].x = t.v1; v[2].x = t.v2; v[3].x = t.v3; }
        }
        if (!done_y) {
            const Top5 t = top_singular_scalar(S[0].y, S[1].y, S[2].y, S[3].y, S[4].y, S[5].y, S[6].y, S[7].y, S[8].y, S[9].y,
                                               S[10].y, S[11].y, S[12].y, S[13].y, S[14].y, S[15].y);
            out.sigma0.y = t.sigma;
            if (kWantVec) { v[0].y = t.v0; v[1].y = t.v1; v[2].y = t.v2; v[3].y = t.v3; }
        }
    }
    return out;
}

}  // namespace b200wm
