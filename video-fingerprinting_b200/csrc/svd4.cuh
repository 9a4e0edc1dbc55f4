// Largest singular value / right singular vector of a 4x4 block, per thread, in registers.
//
// The reference computes u, s, v = np.linalg.svd(cv2.dct(block)) for every 4x4 LL block
// (embed/dwt_dct_svd_encoder.py:43, extract/dwt_dct_svd_decoder.py:35) and only ever uses
// s[0] (extract) or s[0] with its singular pair (embed).  The 2-D DCT is orthogonal, so
// sigma(dct(B)) == sigma(B) and u0 v0^T of dct(B) maps back through idct to u0 v0^T of B:
// the block DCT/IDCT cancels out of this pair and is not computed (checked
// against the oracle in tests/).
//
// Algorithm.  G = S^T S (S = the block; exact in fp32 when S holds 2x2 sums of uint8).
// Power iteration on G from G*1, Rayleigh quotient lambda, residual r = G v - lambda v.
// Because every other eigenvalue is at most tr(G) - lambda, Kato-Temple gives
//     0 <= lambda_0 - lambda <= |r|^2 / (|v|^2 (2 lambda - tr))      whenever 2 lambda > tr,
// which both proves that the iterate sits on the TOP eigenvalue and bounds the error, so the
// loop exits on a rigorous relative bound (kTau).  Blocks that are not dominated
// (2 lambda <= tr) or converge slowly fall back to a cyclic Jacobi eigen-solver in fp64.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace b200wm {

constexpr float kTau = 5.8207661e-11f;     // 2^-34 relative bound on lambda_0 - lambda
constexpr int kSquarings = 8;               // slow path: G^(2^j) steps before giving up on power iteration

// Symmetric 4x4 in packed upper-triangular order: 00 01 02 03 11 12 13 22 23 33.
__device__ __forceinline__ void gram4(const float (&S)[16], float (&G)[10]) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            float acc = S[i] * S[j];
            acc = fmaf(S[4 + i], S[4 + j], acc);
            acc = fmaf(S[8 + i], S[8 + j], acc);
            acc = fmaf(S[12 + i], S[12 + j], acc);
            G[k++] = acc;
        }
}

__device__ __forceinline__ void symv4(const float (&G)[10], const float (&x)[4], float (&y)[4]) {
    y[0] = fmaf(G[3], x[3], fmaf(G[2], x[2], fmaf(G[1], x[1], G[0] * x[0])));
    y[1] = fmaf(G[6], x[3], fmaf(G[5], x[2], fmaf(G[4], x[1], G[1] * x[0])));
    y[2] = fmaf(G[8], x[3], fmaf(G[7], x[2], fmaf(G[5], x[1], G[2] * x[0])));
    y[3] = fmaf(G[9], x[3], fmaf(G[8], x[2], fmaf(G[6], x[1], G[3] * x[0])));
}

__device__ __forceinline__ float dot4(const float (&a)[4], const float (&b)[4]) {
    return fmaf(a[3], b[3], fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])));
}

// Rare path: cyclic Jacobi on G = S^T S in fp64, eigenvectors accumulated.
// out[0..3] = unit eigenvector of the largest eigenvalue (lowest index wins exact ties),
// out[4] = sigma_0 = sqrt of that eigenvalue, rounded once from fp64.
// Deliberately rolled loops over local-memory arrays: this function must not raise the register
// footprint of the kernels that call it once in a few thousand blocks.
struct Top5 { float v0, v1, v2, v3, sigma; };

static __device__ __noinline__ Top5 top_pair_jacobi(float s0, float s1, float s2, float s3, float s4, float s5, float s6, float s7,
                                             float s8, float s9, float s10, float s11, float s12, float s13, float s14,
                                             float s15) {
    // arguments arrive by value so that the caller's block never has its address taken
    float S[16] = {s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15};
    float out[5];
    double A[16], V[16];
#pragma unroll 1
    for (int i = 0; i < 4; ++i)
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k < 4; ++k) acc = fma((double)S[4 * k + i], (double)S[4 * k + j], acc);
            A[4 * i + j] = acc;
            V[4 * i + j] = (i == j) ? 1.0 : 0.0;
        }
    const double tr = A[0] + A[5] + A[10] + A[15];
    const double stop = 1e-34 * tr * tr;
#pragma unroll 1
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0;
#pragma unroll 1
        for (int p = 0; p < 3; ++p)
#pragma unroll 1
            for (int q = p + 1; q < 4; ++q) off = fma(A[4 * p + q], A[4 * p + q], off);
        if (!(off > stop)) break;
#pragma unroll 1
        for (int p = 0; p < 3; ++p)
#pragma unroll 1
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[4 * p + q];
                if (apq == 0.0) continue;
                const double theta = (A[4 * q + q] - A[4 * p + p]) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                const double c = 1.0 / sqrt(fma(t, t, 1.0));
                const double s = t * c;
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[4 * k + p], akq = A[4 * k + q];
                    A[4 * k + p] = c * akp - s * akq;
                    A[4 * k + q] = s * akp + c * akq;
                }
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[4 * p + k], aqk = A[4 * q + k];
                    A[4 * p + k] = c * apk - s * aqk;
                    A[4 * q + k] = s * apk + c * aqk;
                    const double vkp = V[4 * k + p], vkq = V[4 * k + q];
                    V[4 * k + p] = c * vkp - s * vkq;
                    V[4 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
#pragma unroll 1
    for (int i = 1; i < 4; ++i)
        if (A[5 * i] > A[5 * best]) best = i;
    const double lam = A[5 * best];
    double nn = 0.0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) nn = fma(V[4 * k + best], V[4 * k + best], nn);
    const double n = 1.0 / sqrt(nn);
#pragma unroll 1
    for (int k = 0; k < 4; ++k) out[k] = (float)(V[4 * k + best] * n);
    out[4] = (float)sqrt(lam > 0.0 ? lam : 0.0);
    return Top5{out[0], out[1], out[2], out[3], out[4]};
}

// H <- H*H for a symmetric 4x4 in packed upper-triangular order, renormalised by an exact power
// of two so that its trace stays in [0.5, 1).
__device__ __forceinline__ void square_sym4(float (&H)[10]) {
    const float a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5], g = H[6], h = H[7], i = H[8], j = H[9];
    float R[10];
    R[0] = fmaf(d, d, fmaf(c, c, fmaf(b, b, a * a)));
    R[1] = fmaf(d, g, fmaf(c, f, fmaf(b, e, a * b)));
    R[2] = fmaf(d, i, fmaf(c, h, fmaf(b, f, a * c)));
    R[3] = fmaf(d, j, fmaf(c, i, fmaf(b, g, a * d)));
    R[4] = fmaf(g, g, fmaf(f, f, fmaf(e, e, b * b)));
    R[5] = fmaf(g, i, fmaf(f, h, fmaf(e, f, b * c)));
    R[6] = fmaf(g, j, fmaf(f, i, fmaf(e, g, b * d)));
    R[7] = fmaf(i, i, fmaf(h, h, fmaf(f, f, c * c)));
    R[8] = fmaf(i, j, fmaf(h, i, fmaf(f, g, c * d)));
    R[9] = fmaf(j, j, fmaf(i, i, fmaf(g, g, d * d)));
    const float tr = (R[0] + R[4]) + (R[7] + R[9]);
    const float down = __uint_as_float(0x7E800000u - (__float_as_uint(tr) & 0x7F800000u));
#pragma unroll
    for (int k = 0; k < 10; ++k) H[k] = R[k] * down;
}

// One power step with the convergence test.  x is the current iterate, w = G x.
// Returns true when the Kato-Temple bound certifies lambda (see the header comment).
__device__ __forceinline__ bool rayleigh_check(const float (&x)[4], const float (&w)[4], float tr, float& xw, float& xx) {
    xw = dot4(x, w);
    xx = dot4(x, x);
    float lam;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(lam) : "f"(xx));
    lam *= xw;
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = fmaf(-lam, x[k], w[k]);
    const float rr = dot4(r, r);
    const float gap = fmaf(2.0f, lam, -tr);
    return (gap > 0.0f) && (rr <= kTau * lam * gap * xx);
}

// sigma0 = largest singular value of S (not squared); v = unit right singular vector.
// maybe_flat: the first two components of the second power iterate are bitwise equal - necessary for a block
// with sixteen equal entries (every entry of G, hence of G*1 and G*G*1, is then the same number) and next to
// never true otherwise; callers use it as the one-instruction filter in front of the flat-tile rule of
// dwtsvd_tile.cuh.
template <bool kWantVec>
__device__ __forceinline__ float top_singular(const float (&S)[16], float (&v)[4], bool& zero_block, bool& maybe_flat) {
    float G[10];
    gram4(S, G);
    const float tr_raw = (G[0] + G[4]) + (G[7] + G[9]);
    zero_block = !(tr_raw > 0.0f);           // handled with selects at the end: no early exit
    // Exact power-of-two normalisation: tr = tr_raw * 2^-(e+1) lies in [0.5, 1), so the iterates
    // need no rescaling and nothing is rounded (G stays exact for uint8 input).
    const unsigned ebits = __float_as_uint(tr_raw) & 0x7F800000u;
    const float down = __uint_as_float(0x7E800000u - ebits);
    const float up = __uint_as_float(ebits + 0x00800000u);
#pragma unroll
    for (int k = 0; k < 10; ++k) G[k] *= down;
    const float tr = tr_raw * down;

    float x[4], w[4];
    w[0] = (G[0] + G[1]) + (G[2] + G[3]);       // G * ones
    w[1] = (G[1] + G[4]) + (G[5] + G[6]);
    w[2] = (G[2] + G[5]) + (G[7] + G[8]);
    w[3] = (G[3] + G[6]) + (G[8] + G[9]);
    symv4(G, w, x);                              // second power step, unchecked: one step alone almost never certifies
    maybe_flat = x[0] == x[1];
    symv4(G, x, w);
    float xw, xx;
    bool done = rayleigh_check(x, w, tr, xw, xx) || zero_block;
    if (!done) {
        // Slow convergence (second singular value close to the first, typical of chroma planes
        // whose mean is near zero): power steps with G^2, G^4, G^8, ... - the convergence ratio is
        // squared every round - while the certificate is always evaluated with G itself.
        float H[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) H[k] = G[k];
#pragma unroll 1
        for (int round = 0; round < kSquarings && !done; ++round) {
            square_sym4(H);
            symv4(H, w, x);
            // keep the iterate's magnitude in range: exact power-of-two rescale by its largest entry
            const float m = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
            const float up2 = __uint_as_float(0x7E800000u - (__float_as_uint(m) & 0x7F800000u));
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] *= up2;
            symv4(G, x, w);
            done = rayleigh_check(x, w, tr, xw, xx);
            // A block that is not dominated (2 lambda_0 <= tr G: noise-like content, flat chroma) can never be
            // certified.  After three squaring rounds the iterate has seen G^16: if the Rayleigh quotient is still
            // not above tr/2 the block leaves for the direct solver now instead of after all kSquarings rounds
            // (a barely dominated block that would have been certified later takes the direct solver too: same
            // sigma_0 to float32 accuracy).
            if (!done && round >= 2 && !(2.0f * xw > tr * xx)) break;
        }
    }
    if (done) {
        const float lam = div_pos(xw, xx) * up;
        if (kWantVec) {
            const float n = zero_block ? 0.0f : rsqrtf(xx);
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = zero_block ? 0.0f : x[k] * n;
        }
        return zero_block ? 0.0f : sqrt_pos(lam);
    }
    const Top5 t = top_pair_jacobi(S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7], S[8], S[9], S[10], S[11], S[12],
                                   S[13], S[14], S[15]);
    v[0] = t.v0; v[1] = t.v1; v[2] = t.v2; v[3] = t.v3;
    return t.sigma;
}

}  // namespace b200wm
