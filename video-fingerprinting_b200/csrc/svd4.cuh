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
// (2 lambda <= tr) or converge slowly fall back to a cyclic Jacobi eigen-solver (float32, registers, Rayleigh-polished).
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

// Rare path: cyclic Jacobi on G = S^T S, eigenvectors accumulated, for blocks the power iteration cannot certify
// (not dominated: 2 lambda_0 <= tr G; or a start vector next to orthogonal to v_0).  float32 in registers, fully
// unrolled: float32 is all the reference's LAPACK path has, and the result is polished with one Rayleigh quotient
// on G itself, which squares the eigenvector error (Jacobi leaves ~1e-6, the quotient is then good to rounding,
// like the fast path's).  One call costs ~1,300 instructions with ONE lane of the warp active; its predecessor - fp64,
// rolled loops over local-memory arrays - cost 8,700 and was a third of the fused rgb24 extract's instruction
// count on natural chroma, where 0.3 % of the blocks come here (profiles/r02_fused_rgb.md).
// out: unit eigenvector of the largest eigenvalue (lowest index wins exact ties) and that eigenvalue.
struct Top5 { float v0, v1, v2, v3, sigma; };      // sigma: eigenvalue of G on return from top_pair_jacobi(), sigma_0 elsewhere

template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(float (&A)[4][4], float (&V)[4][4]) {
    const float apq = A[P][Q];
    const float d = A[Q][Q] - A[P][P];
    // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = d / (2 apq), written without the division by apq
    const float two = apq + apq;
    const float den = fabsf(d) + sqrtf(fmaf(d, d, two * two));
    const float t = den > 0.0f ? __fdividef(d >= 0.0f ? two : -two, den) : 0.0f;
    const float c = rsqrtf(fmaf(t, t, 1.0f));
    const float sn = t * c;
    A[P][P] = fmaf(-t, apq, A[P][P]);
    A[Q][Q] = fmaf(t, apq, A[Q][Q]);
    A[P][Q] = A[Q][P] = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k != P && k != Q) {
            const float akp = A[k][P], akq = A[k][Q];
            A[k][P] = A[P][k] = fmaf(c, akp, -sn * akq);
            A[k][Q] = A[Q][k] = fmaf(sn, akp, c * akq);
        }
        const float vkp = V[k][P], vkq = V[k][Q];
        V[k][P] = fmaf(c, vkp, -sn * vkq);
        V[k][Q] = fmaf(sn, vkp, c * vkq);
    }
}

static __device__ __noinline__ Top5 top_pair_jacobi(float g0, float g1, float g2, float g3, float g4, float g5, float g6, float g7,
                                             float g8, float g9) {
    // G = S^T S in packed upper-triangular order (00 01 02 03 11 12 13 22 23 33), any scaling (the caller's is
    // normalised to a trace in [0.5, 1)); arguments arrive by value so that the caller's registers never have their
    // address taken.  Returns the eigenvalue in the scaling of the input.
    float G[4][4], A[4][4], V[4][4];
    G[0][0] = g0; G[0][1] = G[1][0] = g1; G[0][2] = G[2][0] = g2; G[0][3] = G[3][0] = g3;
    G[1][1] = g4; G[1][2] = G[2][1] = g5; G[1][3] = G[3][1] = g6;
    G[2][2] = g7; G[2][3] = G[3][2] = g8; G[3][3] = g9;
    const float tr_raw = (G[0][0] + G[1][1]) + (G[2][2] + G[3][3]);
    if (!(tr_raw > 0.0f)) return Top5{0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    // exact power-of-two normalisation, as in top_singular(): tr in [0.5, 1)
    const unsigned ebits = __float_as_uint(tr_raw) & 0x7F800000u;
    const float down = __uint_as_float(0x7E800000u - ebits), up = __uint_as_float(ebits + 0x00800000u);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            G[i][j] *= down;
            A[i][j] = G[i][j];
            V[i][j] = i == j ? 1.0f : 0.0f;
        }
#pragma unroll 1
    for (int sweep = 0; sweep < 8; ++sweep) {
        const float off = fmaf(A[0][1], A[0][1], fmaf(A[0][2], A[0][2], fmaf(A[0][3], A[0][3],
                          fmaf(A[1][2], A[1][2], fmaf(A[1][3], A[1][3], A[2][3] * A[2][3])))));
        if (!(off > 1e-16f)) break;              // (1e-8 tr)^2: the quotient below squares what is left
        jacobi_rotate<0, 1>(A, V); jacobi_rotate<0, 2>(A, V); jacobi_rotate<0, 3>(A, V);
        jacobi_rotate<1, 2>(A, V); jacobi_rotate<1, 3>(A, V); jacobi_rotate<2, 3>(A, V);
    }
    int best = 0;
    float top = A[0][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (A[i][i] > top) { top = A[i][i]; best = i; }
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = best == 0 ? V[k][0] : (best == 1 ? V[k][1] : (best == 2 ? V[k][2] : V[k][3]));
    // Rayleigh quotient on G itself
    float w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = fmaf(G[k][3], x[3], fmaf(G[k][2], x[2], fmaf(G[k][1], x[1], G[k][0] * x[0])));
    const float xw = fmaf(x[3], w[3], fmaf(x[2], w[2], fmaf(x[1], w[1], x[0] * w[0])));
    const float xx = fmaf(x[3], x[3], fmaf(x[2], x[2], fmaf(x[1], x[1], x[0] * x[0])));
    const float n = rsqrtf(xx);
    return Top5{x[0] * n, x[1] * n, x[2] * n, x[3] * n, div_pos(xw, xx) * up};
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
// kDeep: number of squarings of G applied in STRAIGHT-LINE code before the first certificate.  0 suits luma
// (99.8 % of natural luma blocks certify after G^3 * 1).  Chroma blocks (signed float samples around zero, e.g.
// the U channel the reference marks) are far less dominated - sigma_1/sigma_0 has a median of 0.13 and a 99th
// percentile of 0.56 on the reference's fixture against 0.008 / 0.07 for its luma - so that half of them fail after
// G^3 * 1 and practically every warp would enter the data-dependent loop below with a few lanes.  With kDeep = 3 the
// iterate is G^16 e_k for 165 more instructions on every lane, which certifies blocks up to sigma_1/sigma_0 = 0.7.
template <bool kWantVec, int kDeep = 0>
__device__ __forceinline__ float top_singular_gram(float (&G)[10], float (&v)[4], bool& zero_block, bool& maybe_flat) {
    // G = S^T S as gram4() leaves it (callers that never hold the whole block accumulate it row by row in the same order)
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
    float H[10];
    if (kDeep > 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) H[k] = G[k];
#pragma unroll
        for (int s = 0; s < kDeep; ++s) square_sym4(H);           // H = G^(2^kDeep), trace in [0.5, 1)
        // Start vector: the column of H with the largest diagonal entry, i.e. G^(2^kDeep) e_k with the k whose
        // H_kk = sum_i mu_i v_i[k]^2 is largest.  Unlike the all-ones vector (ideal for positive luma blocks, the Perron
        // vector) it cannot be orthogonal to v_0, which zero-mean chroma blocks with an oscillating pattern are to
        // all-ones: those blocks never certified and took the direct solver.
        const bool k1 = H[4] > H[0], k23 = H[9] > H[7];
        const float d01 = k1 ? H[4] : H[0], d23 = k23 ? H[9] : H[7];
        const bool hi = d23 > d01;
        // columns of the packed symmetric H: 0 = {0,1,2,3}, 1 = {1,4,5,6}, 2 = {2,5,7,8}, 3 = {3,6,8,9}
        x[0] = hi ? (k23 ? H[3] : H[2]) : (k1 ? H[1] : H[0]);
        x[1] = hi ? (k23 ? H[6] : H[5]) : (k1 ? H[4] : H[1]);
        x[2] = hi ? (k23 ? H[8] : H[7]) : (k1 ? H[5] : H[2]);
        x[3] = hi ? (k23 ? H[9] : H[8]) : (k1 ? H[6] : H[3]);
        symv4(H, x, w);                                            // G^(2^(kDeep+1)) e_k
        const float m = fmaxf(fmaxf(fabsf(w[0]), fabsf(w[1])), fmaxf(fabsf(w[2]), fabsf(w[3])));
        const float up2 = __uint_as_float(0x7E800000u - (__float_as_uint(m) & 0x7F800000u));   // exact power of two
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = w[k] * up2;
    } else {
        symv4(G, w, x);                          // second power step, unchecked: one step alone almost never certifies
    }
    maybe_flat = x[0] == x[1];
    symv4(G, x, w);
    float xw, xx;
    bool done = rayleigh_check(x, w, tr, xw, xx) || zero_block;
    if (!done) {
        // Slow convergence (second singular value close to the first, typical of chroma planes
        // whose mean is near zero): power steps with G^2, G^4, G^8, ... - the convergence ratio is
        // squared every round - while the certificate is always evaluated with G itself.
        if (kDeep == 0) {
#pragma unroll
            for (int k = 0; k < 10; ++k) H[k] = G[k];
        }
#pragma unroll 1
        for (int round = 0; round < kSquarings - kDeep && !done; ++round) {
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
            if (!done && round + kDeep >= 2 && !(2.0f * xw > tr * xx)) break;
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
    const Top5 t = top_pair_jacobi(G[0], G[1], G[2], G[3], G[4], G[5], G[6], G[7], G[8], G[9]);      // on the normalised G
    v[0] = t.v0; v[1] = t.v1; v[2] = t.v2; v[3] = t.v3;
    const float lam = t.sigma * up;
    return lam > 0.0f ? sqrt_pos(lam) : 0.0f;
}

template <bool kWantVec, int kDeep = 0>
__device__ __forceinline__ float top_singular(const float (&S)[16], float (&v)[4], bool& zero_block, bool& maybe_flat) {
    float G[10];
    gram4(S, G);
    return top_singular_gram<kWantVec, kDeep>(G, v, zero_block, maybe_flat);
}

}  // namespace b200wm
