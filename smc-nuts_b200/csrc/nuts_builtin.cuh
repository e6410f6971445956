// Launch configuration and dispatch helpers of the BUILT-IN models, shared by nuts_kernel.cu and nuts_kernel_scaled.cu.
#pragma once
#include "nuts_launch.cuh"

namespace smcb {

#ifndef SMCB_ARMA_MIN_BLOCKS
#define SMCB_ARMA_MIN_BLOCKS 4
#endif
template <> struct LaunchCfg<ArmaModel> { static constexpr int NT = 128, MIN_BLOCKS = SMCB_ARMA_MIN_BLOCKS; };
template <> struct LaunchCfg<PrmModel> { static constexpr int NT = 128, MIN_BLOCKS = 2; };
template <> struct LaunchCfg<GaussModel> { static constexpr int NT = 128, MIN_BLOCKS = 1; };
// PrmModelG, MEASURED (N = 2^20): 4 CTAs/SM at 118 registers 130.5 ms; 3 CTAs/SM 140+ ms; 5 CTAs/SM (96 registers, spills) 132-152 ms
template <int T8> struct LaunchCfg<PrmModelG<T8>> { static constexpr int NT = 128, MIN_BLOCKS = 4; };
constexpr int kPrmTiles = 13;   // PrmModelG instantiation: 81..104 observations (the shipped PRMwCD has 100)

#if SMCB_ALIGN_PARITY == 2   // A/B experiments: also for the one-lane-per-particle kernels
template <> struct AlignCfg<ArmaModel> { static constexpr bool ON = true; };
template <> struct AlignCfg<PrmModel> { static constexpr bool ON = true; };
#endif

template <int NT8> struct StageOffset<GaussModelG<NT8>> { static int of(const ModelDesc& d) { return d.dim * d.dim; } };
template <int T8> struct StageOffset<PrmModelG<T8>> { static int of(const ModelDesc& d) { return PrmModel::HDR + d.T * PrmModel::ROW; } };

// PRMwCD runs on the tensor-core group kernel when the observation count fits the instantiated tile count
// (SMCB_PRM_SCALAR=1 forces the one-lane-per-particle kernel: A/B experiments and the parity test of the two)
// MEASURED (B200, tools/ab_time.py PRMwCD 16..20, fixed inputs): the group kernel wins at every size -- 13.2 vs 30.9 ms at
// N = 2^16, 23.2 vs 36.8 ms at 2^17 (its trip latency is ~4x shorter, and the 2047-leapfrog trees set the makespan of a
// small shard), 140.2 vs 150.7 ms at 2^20 -- once its tile loop is rolled so that the kernel fits the instruction cache.
static bool prm_use_group(const ModelDesc& d, long long N) {
    const char* e = getenv("SMCB_PRM_SCALAR");   // 1: one lane per particle (A/B experiments, parity test of the two)
    (void)N;
    return PrmModelG<kPrmTiles>::fits(d) && !(e && atoi(e) != 0);
}

// Gaussian: tensor-core group kernel for D <= 104, one-lane-per-particle fallback above (SMCB_GAUSS_SCALAR=1 forces the
// fallback: parity test of the two)
static bool gauss_force_scalar() {
    const char* e = getenv("SMCB_GAUSS_SCALAR");
    return e && atoi(e) != 0;
}
#define SMCB_GAUSS_DISPATCH(D, CALL_G, CALL_PLAIN) \
    (gauss_force_scalar() ? CALL_PLAIN : (D) <= 8 ? CALL_G(1) : (D) <= 16 ? CALL_G(2) : (D) <= 32 ? CALL_G(4) : (D) <= 64 ? CALL_G(8) : (D) <= 104 ? CALL_G(13) : CALL_PLAIN)


// diagonal-metric variants of the same dispatch (csrc/nuts_kernel_scaled.cu; separate translation unit: builds in parallel)
long long nuts_ws_bytes_scaled(const Model* m, long long N, int max_depth);
int launch_nuts_scaled(const Model* m, NutsArgs a, long long workspace_bytes, cudaStream_t st);

}  // namespace smcb
