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

namespace b200wm {

constexpr float kTau = 5.8207661e-11f;     // 2^-34 relative bound on lambda_0 - lambda
constexpr int kPowerIters = 8;

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
__device__ __noinline__ void top_pair_jacobi(const float* __restrict__ S, float* __restrict__ out) {
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
}

// sigma0 = largest singular value of S (not squared); v = unit right singular vector.
template <bool kWantVec>
__device__ __forceinline__ float top_singular(const float (&S)[16], float (&v)[4], bool& zero_block) {
    float G[10];
    gram4(S, G);
    const float tr = (G[0] + G[4]) + (G[7] + G[9]);
    zero_block = !(tr > 0.0f);
    if (zero_block) {
        v[0] = v[1] = v[2] = v[3] = 0.0f;
        return 0.0f;
    }
    const float rtr = __frcp_rn(tr);
    float x[4];
    x[0] = ((G[0] + G[1]) + (G[2] + G[3])) * rtr;
    x[1] = ((G[1] + G[4]) + (G[5] + G[6])) * rtr;
    x[2] = ((G[2] + G[5]) + (G[7] + G[8])) * rtr;
    x[3] = ((G[3] + G[6]) + (G[8] + G[9])) * rtr;

    bool done = false;
    float xw = 0.0f, xx = 1.0f;
#pragma unroll 1
    for (int it = 0; it < kPowerIters; ++it) {
        float w[4];
        symv4(G, x, w);
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
        done = (gap > 0.0f) && (rr <= kTau * lam * gap * xx);
        if (done) break;
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = w[k] * rtr;
    }
    if (done) {
        const float lam = xw / xx;                 // IEEE division: this is the result
        if (kWantVec) {
            const float n = rsqrtf(xx);
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = x[k] * n;
        }
        return sqrtf(lam);
    }
    // copy so that S itself never has its address taken (keeps it in registers on the hot path)
    float Sc[16], vj[5];
#pragma unroll
    for (int k = 0; k < 16; ++k) Sc[k] = S[k];
    top_pair_jacobi(Sc, vj);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = vj[k];
    return vj[4];
}

}  // namespace b200wm
