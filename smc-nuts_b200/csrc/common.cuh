// Common macros for the smcnuts B200 kernels.
//
// Every header under csrc/ that holds algorithm logic (philox.cuh, models.cuh, nuts_lane.cuh) is written
// against SMCB_HD so that the SAME source also compiles with plain g++ for tests/hostsim (a CPU
// simulation of the device lanes used only by the `not gpu` test-suite to validate kernel logic against
// the oracle; it is not part of, nor reachable from, the product library).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define SMCB_HD __host__ __device__ __forceinline__
#define SMCB_D __device__ __forceinline__
#else
#define SMCB_HD inline
#define SMCB_D inline
#endif

namespace smcb {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kLogPi = 1.1447298858494001741434273513531;
constexpr double kTwoPi = 6.283185307179586476925286766559;

SMCB_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// index of lowest set bit, v != 0
SMCB_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
SMCB_HD double neg_inf() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0xfff0000000000000ULL);
#else
    return -INFINITY;
#endif
}
SMCB_HD bool is_finite(double v) {
#if defined(__CUDA_ARCH__)
    return isfinite(v);
#else
    return std::isfinite(v);
#endif
}

// exp() for the hot loops (PRMwCD: one per observation per leapfrog; log-sum-exp / normalise: one per particle).
// Table-driven: x = (32 k + j) ln2/32 + r with |r| <= ln2/64, exp(x) = 2^k * T[j] * (1 + q(r)), T[j] = 2^(j/32) from a
// 256-byte table (two L1 lines, fetched while the polynomial runs) and q = exp(r) - 1 a degree-6 Taylor polynomial in
// Estrin form: 12 FP64-pipe instructions with a 7-deep dependency chain, against 19 / 12-deep for the table-free
// degree-13 version this replaces and ~25 for libdevice's exp -- the FP64 pipe is the bound of the PRMwCD kernels,
// 35 % of whose pipe time was exp.  Max error 0.97 ulp (tests/test_gpu_parity.py checks <= 1.5 ulp against mpmath).
// For |x| < 708 (one integer compare on the high word, which also catches inf and nan) the result is a normal number
// and 2^k is applied by adding k to the exponent field: no FP64 compares, selects or scaling multiplications on the
// hot path.  Everything else takes the rare out-of-line tail: two exact scaling steps so that overflow and gradual
// underflow come out of the multiplications themselves.  Host builds (tests/hostsim) use std::exp so that they stay
// bit-identical to the oracle.
#if defined(__CUDACC__)
static __device__ const double kExpT[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0,
    0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0,
    0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0,
    0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, 0x1.ae89f995ad3adp+0,
    0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
    0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
constexpr int kExpFastHi = 0x40862000;            // high word of 708.0
constexpr double kExpInvL = 0x1.71547652b82fep+5;  // 32 / ln 2
constexpr double kExpLHi = 0x1.62e42fef80000p-6;   // ln 2 / 32, 34 significant bits (k' * hi is exact)
constexpr double kExpLLo = 0x1.1cf79abc9e3b4p-41;
constexpr double kExpMagic = 6755399441055744.0;   // 1.5 * 2^52: the low word of x/L + magic is round(x/L)

// p * 2^k for out-of-range arguments (|x| >= 708, inf, nan); p = T[j] exp(r) of the reduced argument
static __device__ __noinline__ double fast_exp_tail(double x, double p, int k) {
    const int k1 = k >> 1, k2 = k - k1;
    double res = p * __hiloint2double((k1 + 1023) << 20, 0) * __hiloint2double((k2 + 1023) << 20, 0);
    res = (x > 709.782712893384) ? __longlong_as_double(0x7ff0000000000000LL) : res;
    res = (x < -745.2) ? 0.0 : res;
    return (x != x) ? x : res;
}
__device__ __forceinline__ double fast_exp_scale(double p, int k) {
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
#endif
SMCB_HD double fast_exp(double x) {
#if defined(__CUDA_ARCH__)
    double t = fma(x, kExpInvL, kExpMagic);
    const int kp = __double2loint(t);
    const double tj = __ldg(&kExpT[kp & 31]);
    t -= kExpMagic;
    double r = fma(t, -kExpLHi, x);
    r = fma(t, -kExpLLo, r);
    const double s = r * r;
    const double b0 = fma(1.0 / 6, r, 0.5);
    double b1 = fma(1.0 / 120, r, 1.0 / 24);
    b1 = fma(1.0 / 720, s, b1);
    const double q = fma(s, fma(s, b1, b0), r);
    const double p = fma(tj, q, tj);
    if ((__double2hiint(x) & 0x7fffffff) >= kExpFastHi) return fast_exp_tail(x, p, kp >> 5);
    return fast_exp_scale(p, kp >> 5);
#else
    return std::exp(x);
#endif
}

// Two independent exps with their instruction streams interleaved statement by statement (the compiler keeps the
// source order inside a basic block, so this doubles the ILP of the latency-bound chains).
SMCB_HD void fast_exp_pair(double xa, double xb, double& ea, double& eb) {
#if defined(__CUDA_ARCH__)
    double ta = fma(xa, kExpInvL, kExpMagic), tb = fma(xb, kExpInvL, kExpMagic);
    const int ka = __double2loint(ta), kb = __double2loint(tb);
    const double ja = __ldg(&kExpT[ka & 31]), jb = __ldg(&kExpT[kb & 31]);
    ta -= kExpMagic; tb -= kExpMagic;
    double ra = fma(ta, -kExpLHi, xa), rb = fma(tb, -kExpLHi, xb);
    ra = fma(ta, -kExpLLo, ra); rb = fma(tb, -kExpLLo, rb);
    const double sa = ra * ra, sb = rb * rb;
    const double a0 = fma(1.0 / 6, ra, 0.5), c0 = fma(1.0 / 6, rb, 0.5);
    double a1 = fma(1.0 / 120, ra, 1.0 / 24), c1 = fma(1.0 / 120, rb, 1.0 / 24);
    a1 = fma(1.0 / 720, sa, a1); c1 = fma(1.0 / 720, sb, c1);
    const double qa = fma(sa, fma(sa, a1, a0), ra), qb = fma(sb, fma(sb, c1, c0), rb);
    const double pa = fma(ja, qa, ja), pb = fma(jb, qb, jb);
    const int ha = __double2hiint(xa) & 0x7fffffff, hb = __double2hiint(xb) & 0x7fffffff;
    ea = fast_exp_scale(pa, ka >> 5); eb = fast_exp_scale(pb, kb >> 5);
    if ((ha > hb ? ha : hb) >= kExpFastHi) {   // rare: out-of-range, inf or nan argument
        if (ha >= kExpFastHi) ea = fast_exp_tail(xa, pa, ka >> 5);
        if (hb >= kExpFastHi) eb = fast_exp_tail(xb, pb, kb >> 5);
    }
#else
    ea = std::exp(xa); eb = std::exp(xb);
#endif
}

// model kinds of the C-ABI (include/smcnuts_b200.h)
enum ModelKind : int { kArma = 0, kPRMwCD = 1, kGauss = 2 };

}  // namespace smcb
