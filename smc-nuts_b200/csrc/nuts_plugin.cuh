// Model plug-in: one generated (or hand-written) model struct compiled into its own shared object together with the
// NUTS / logp kernel templates of nuts_launch.cuh.  The product library loads it with smcb_model_create_plugin and routes
// smcb_logp_grad, smcb_nuts_workspace_bytes and smcb_nuts_transition of that model handle through the entry points below;
// every model-agnostic kernel (weights, resampling, L-kernels, estimators) is the library's own.
//
// Replaces the generic half of /root/reference/smcnuts/model/bridgestan.py:13-26 (any Stan program through BridgeStan):
// smcnuts/model/stan_codegen.py writes the struct, smcnuts/model/generated.py writes and compiles a translation unit
//
//     #define SMCB_PLUGIN_TU 1
//     #include "nuts_plugin.cuh"
//     #include "model_gen.cuh"          // struct GenModel { ... eval(x, phi, A, B, g) ... }
//     SMCB_DEFINE_PLUGIN(GenModel)
//
// with nvcc -gencode arch=compute_100a,code=sm_100a.
#pragma once
#ifndef SMCB_PLUGIN_TU
#error "define SMCB_PLUGIN_TU before including nuts_plugin.cuh"
#endif
#include "nuts_launch.cuh"

#define SMCB_PLUGIN_ABI 1

// launch shape of a plug-in model: 128 threads, two CTAs per SM (its register need is unknown; wide models spill)
#define SMCB_DEFINE_PLUGIN(MODEL)                                                                                        \
    namespace smcb {                                                                                                     \
    template <> struct LaunchCfg<MODEL> { static constexpr int NT = 128, MIN_BLOCKS = 2; };                              \
    }                                                                                                                    \
    extern "C" {                                                                                                         \
    int smcb_plugin_abi(void) { return SMCB_PLUGIN_ABI; }                                                                \
    int smcb_plugin_dim(void) { return MODEL::STATIC_D; }                                                                \
    int smcb_plugin_ndata(void) { return MODEL::NDATA; }                                                                 \
    const char* smcb_plugin_last_error(void) { return smcb::last_error_ref().c_str(); }                                  \
    long long smcb_plugin_nuts_workspace_bytes(const smcb::ModelDesc* d, long long N, int max_depth) {                   \
        smcb::Model m{*d, nullptr};                                                                                      \
        if (d->scale) return smcb::nuts_ws_bytes<smcb::ScaledModel<MODEL>>(&m, N, max_depth);                            \
        return smcb::nuts_ws_bytes<MODEL>(&m, N, max_depth);                                                             \
    }                                                                                                                    \
    int smcb_plugin_nuts_transition(const smcb::ModelDesc* d, const smcb::NutsArgs* a, long long ws_bytes, void* st) {   \
        smcb::Model m{*d, nullptr};                                                                                      \
        if (d->scale) return smcb::launch_nuts<smcb::ScaledModel<MODEL>>(&m, *a, ws_bytes, (cudaStream_t)st);            \
        return smcb::launch_nuts<MODEL>(&m, *a, ws_bytes, (cudaStream_t)st);                                             \
    }                                                                                                                    \
    int smcb_plugin_logp_grad(const smcb::ModelDesc* d, const double* x, long long N, double phi, double* A, double* B,  \
                              double* g, void* st) {                                                                     \
        smcb::Model m{*d, nullptr};                                                                                      \
        return smcb::launch_logp<MODEL>(&m, x, N, phi, A, B, g, (cudaStream_t)st);                                       \
    }                                                                                                                    \
    }
