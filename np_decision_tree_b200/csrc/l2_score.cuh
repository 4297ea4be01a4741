// Candidate record and scoring helpers shared by the L2 search kernels (l2.cu, fused.cu).
#pragma once
#include <math_constants.h>

#include "tree_kernels.cuh"

// 1/d to ~1e-12 relative: hardware seed + one Newton step (approximate search only; the
// near-tie threshold is 1e-9)
__device__ __forceinline__ double rcp_fast1(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    return fma(x, fma(-d, x, 1.0), x);
}

constexpr int SY_PITCH = WT_TILE + WT_TILE / 16;  // 1 pad double per 16: conflict-free blocked reads

struct Cand {
    double g;    // score (maximised)
    double g2;   // best score among the other candidates merged so far
    double s;    // prefix sum at k
    double dev;  // |S - (k+1) mean|
    int k;
};

__device__ __forceinline__ Cand cand_none() {
    Cand c;
    c.g = -CUDART_INF;
    c.g2 = -CUDART_INF;
    c.s = 0.0;
    c.dev = 0.0;
    c.k = 0x7fffffff;
    return c;
}

// same column: higher score wins, then lower position (ref one.py:62)
__device__ __forceinline__ void cand_merge(Cand &a, const Cand &b) {
    const bool b_wins = (b.g > a.g) || (b.g == a.g && b.k < a.k);
    const double loser = b_wins ? a.g : b.g;
    const double g2 = fmax(fmax(a.g2, b.g2), loser);
    if (b_wins) {
        a.g = b.g;
        a.s = b.s;
        a.dev = b.dev;
        a.k = b.k;
    }
    a.g2 = g2;
}

__device__ __forceinline__ Cand cand_shfl_xor(const Cand &c, int m) {
    Cand o;
    o.g = shfl_xor_d(c.g, m);
    o.g2 = shfl_xor_d(c.g2, m);
    o.s = shfl_xor_d(c.s, m);
    o.dev = shfl_xor_d(c.dev, m);
    o.k = __shfl_xor_sync(FULL_MASK, c.k, m);
    return o;
}

__device__ __forceinline__ void cand_warp_reduce(Cand &c) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        Cand o = cand_shfl_xor(c, m);
        cand_merge(c, o);
    }
}

// ---- approximate (parallel-scan) score at position kk of a node with n rows ------------
__device__ __forceinline__ double score_fast(double S, int kk, int n, double mean, int variant, double &dev) {
    if (variant == 1) {
        // ref one.py:57-61: w[k] * (S - (k+1) mean)^2, w = 1 / (float32(k+1) * (n-k-1))
        dev = fabs(S - (double)(kk + 1) * mean);
        const double ab = (double)__int2float_rn(kk + 1) * (double)(n - 1 - kk);
        return dev * dev * rcp_fast1(ab);
    }
    // ref one.py:85-90: |sqrt(1/(l r)) S - mean sqrt(l/r)| with float32 ratios
    const float lc = __int2float_rn(kk + 1), rc = __int2float_rn(n - 1 - kk);
    const float ra = __fsqrt_rn(__frcp_rn(__fmul_rn(lc, rc)));
    const float rb = __fsqrt_rn(__fdiv_rn(lc, rc));
    dev = 0.0;
    return fabs((double)ra * S - mean * (double)rb);
}

// ---- reference-order score (no contraction, IEEE division) ------------------------------
__device__ __forceinline__ double score_exact(double S, int kk, int n, double mean, int variant, double &dev) {
    if (variant == 1) {
        dev = fabs(__dsub_rn(S, __dmul_rn((double)(kk + 1), mean)));                       // one.py:57-58
        const double w = __ddiv_rn(1.0, __dmul_rn((double)__int2float_rn(kk + 1), (double)(n - 1 - kk)));  // one.py:52-53
        return __dmul_rn(w, __dmul_rn(dev, dev));                                            // one.py:61
    }
    const float lc = __int2float_rn(kk + 1), rc = __int2float_rn(n - 1 - kk);
    const float ra = __fsqrt_rn(__frcp_rn(__fmul_rn(lc, rc)));                               // one.py:85-86
    const float rb = __fsqrt_rn(__fdiv_rn(lc, rc));                                          // one.py:87-88
    dev = 0.0;
    return fabs(__dsub_rn(__dmul_rn((double)ra, S), __dmul_rn(mean, (double)rb)));          // one.py:89-90
}

// One warp owns one tile of WT_TILE sorted positions of one (node, column); warps of a persistent
// grid take tiles in a static interleaved order (tile-major: consecutive tiles are the same
// position range of different columns), so no block barrier and no atomic ticket is needed:
// a tile's predecessors in its chain have smaller indices and belong to resident warps that
// process their own tiles in increasing order.
