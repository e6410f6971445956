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
// Same algorithm as libdevice's exp (magic-number rounding of x/ln2, two-term Cody-Waite reduction, scaling through
// the exponent field) but with a degree-13 Taylor polynomial split into even and odd halves: two 5-FMA chains that
// run in parallel instead of one 13-deep Horner chain, and the coefficients come from the constant bank as FMA
// operands instead of being re-materialised with moves every call.  |error| <= 1.5 ulp on |x| <= 708 (checked against
// mpmath in tests/test_gpu_parity.py); overflow / underflow / inf / nan are handled with selects.  Host builds (tests/hostsim) use
// std::exp so that they stay bit-identical to the oracle.
#if defined(__CUDACC__)
__constant__ double kExpC[14] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
#endif
SMCB_HD double fast_exp(double x) {
#if defined(__CUDA_ARCH__)
    double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
    const int k = __double2loint(t);
    t -= 6755399441055744.0;
    double r = fma(t, -6.93147180559945286227e-01, x);
    r = fma(t, -2.31904681384629955842e-17, r);
    // exp(r) = 1 + (r + r^2 * R(r)),  R = c2 + c3 r + ... + c13 r^11 evaluated as Re(r^2) + r * Ro(r^2)
    const double r2 = r * r;
    double e = fma(kExpC[12], r2, kExpC[10]), o = fma(kExpC[13], r2, kExpC[11]);
    e = fma(e, r2, kExpC[8]); o = fma(o, r2, kExpC[9]);
    e = fma(e, r2, kExpC[6]); o = fma(o, r2, kExpC[7]);
    e = fma(e, r2, kExpC[4]); o = fma(o, r2, kExpC[5]);
    e = fma(e, r2, kExpC[2]); o = fma(o, r2, kExpC[3]);
    const double p = 1.0 + fma(r2, fma(r, o, e), r);
    // scale by 2^k in two exact steps (k = k1 + k2) so that overflow and gradual underflow come out of the
    // multiplications themselves; the range ends are selects, not branches, which keeps the surrounding loop body one
    // basic block (the compiler can then interleave the chains of neighbouring observations / elements)
    const int k1 = k >> 1, k2 = k - k1;
    double res = p * __hiloint2double((k1 + 1023) << 20, 0) * __hiloint2double((k2 + 1023) << 20, 0);
    res = (x > 709.782712893384) ? __longlong_as_double(0x7ff0000000000000LL) : res;
    res = (x < -745.2) ? 0.0 : res;
    res = (x != x) ? x : res;
    return res;
#else
    return std::exp(x);
#endif
}

// Two independent exps with their instruction streams interleaved statement by statement (the compiler keeps the
// source order inside a basic block, so this doubles the ILP of the latency-bound polynomial chains).
SMCB_HD void fast_exp_pair(double xa, double xb, double& ea, double& eb) {
#if defined(__CUDA_ARCH__)
    double ta = fma(xa, 1.4426950408889634074, 6755399441055744.0), tb = fma(xb, 1.4426950408889634074, 6755399441055744.0);
    const int ka = __double2loint(ta), kb = __double2loint(tb);
    ta -= 6755399441055744.0; tb -= 6755399441055744.0;
    double ra = fma(ta, -6.93147180559945286227e-01, xa), rb = fma(tb, -6.93147180559945286227e-01, xb);
    ra = fma(ta, -2.31904681384629955842e-17, ra); rb = fma(tb, -2.31904681384629955842e-17, rb);
    const double qa = ra * ra, qb = rb * rb;
    double pa = fma(kExpC[12], qa, kExpC[10]), pb = fma(kExpC[12], qb, kExpC[10]);
    double oa = fma(kExpC[13], qa, kExpC[11]), ob = fma(kExpC[13], qb, kExpC[11]);
    pa = fma(pa, qa, kExpC[8]); pb = fma(pb, qb, kExpC[8]); oa = fma(oa, qa, kExpC[9]); ob = fma(ob, qb, kExpC[9]);
    pa = fma(pa, qa, kExpC[6]); pb = fma(pb, qb, kExpC[6]); oa = fma(oa, qa, kExpC[7]); ob = fma(ob, qb, kExpC[7]);
    pa = fma(pa, qa, kExpC[4]); pb = fma(pb, qb, kExpC[4]); oa = fma(oa, qa, kExpC[5]); ob = fma(ob, qb, kExpC[5]);
    pa = fma(pa, qa, kExpC[2]); pb = fma(pb, qb, kExpC[2]); oa = fma(oa, qa, kExpC[3]); ob = fma(ob, qb, kExpC[3]);
    pa = fma(ra, oa, pa); pb = fma(rb, ob, pb);
    pa = 1.0 + fma(qa, pa, ra); pb = 1.0 + fma(qb, pb, rb);
    const int ka1 = ka >> 1, kb1 = kb >> 1;
    double sa = pa * __hiloint2double((ka1 + 1023) << 20, 0) * __hiloint2double((ka - ka1 + 1023) << 20, 0);
    double sb = pb * __hiloint2double((kb1 + 1023) << 20, 0) * __hiloint2double((kb - kb1 + 1023) << 20, 0);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    sa = (xa > 709.782712893384) ? inf : sa; sb = (xb > 709.782712893384) ? inf : sb;
    sa = (xa < -745.2) ? 0.0 : sa; sb = (xb < -745.2) ? 0.0 : sb;
    ea = (xa != xa) ? xa : sa; eb = (xb != xb) ? xb : sb;
#else
    ea = std::exp(xa); eb = std::exp(xb);
#endif
}

// model kinds of the C-ABI (include/smcnuts_b200.h)
enum ModelKind : int { kArma = 0, kPRMwCD = 1, kGauss = 2 };

}  // namespace smcb
