// yk_score.h -- the float64 node sort key, shared by host and device code so both produce the
// same bits.  Restates yunikorn-core objects/node.go GetResourceUsageShares + objects/nodesorting.go
// absResourceUsage / ScoreNode [EXT, SURVEY.md Appendix A.3]:
//   share_k = 1 - float64(available_k)/float64(total_k)      for every k in total
//   usage   = sum_k share_k * w_k,  W = sum_k w_k             over k with w_k != 0 and share not NaN
//   fair key = usage/W (0 when W == 0);  binpacking key = 1 - usage/W
// IEEE-754 double, round-to-nearest-even, NO fused multiply-add: this translation unit must be built
// with nvcc --fmad=false and the host compiler with -ffp-contract=off.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define YK_HD __host__ __device__ __forceinline__
#else
#define YK_HD inline
#endif

#define YK_KEY_NAN 0xFFFFFFFFFFFFFFFFull

// total/avail are read with stride `ld` (column-major SoA: element k of node n at p[k*ld + n])
YK_HD double yk_node_score(int D, uint32_t policy, const double* w, const int64_t* total, const int64_t* avail,
                           size_t ld) {
    double usage = 0.0, tw = 0.0;
    for (int k = 0; k < D; ++k) {
        if (w[k] == 0.0) continue;
        int64_t t = total[(size_t)k * ld];
        // Go float division [EXT GetResourceUsageShares]: x/0 = +-Inf (the infinite share counts), 0/0 = NaN (skipped)
        double share = 1.0 - (double)avail[(size_t)k * ld] / (double)t;
        if (share != share) continue;
        usage += share * w[k];
        tw += w[k];
    }
    // usage / tw; for tw = 1 or 2 (the default weights) the quotient is formed without a divide: x / 2 and x * 0.5 are the
    // same real number, so their correctly rounded doubles are the same bits
    double a = (tw == 0.0) ? 0.0 : (tw == 2.0 ? usage * 0.5 : (tw == 1.0 ? usage : usage / tw));
    double s = (policy == 1u) ? 1.0 - a : a;
    return s + 0.0;  // -0.0 -> +0.0
}

// monotone map double -> uint64 so that unsigned compare == float compare (NaN -> YK_KEY_NAN)
YK_HD uint64_t yk_key_bits(double s) {
    if (s != s) return YK_KEY_NAN;
    uint64_t b;
#if defined(__CUDA_ARCH__)
    b = (uint64_t)__double_as_longlong(s);
#else
    memcpy(&b, &s, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

YK_HD double yk_key_to_score(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double s;
#if defined(__CUDA_ARCH__)
    s = __longlong_as_double((long long)b);
#else
    memcpy(&s, &b, 8);
#endif
    return s;
}
