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

// model kinds of the C-ABI (include/smcnuts_b200.h)
enum ModelKind : int { kArma = 0, kPRMwCD = 1, kGauss = 2 };

}  // namespace smcb
