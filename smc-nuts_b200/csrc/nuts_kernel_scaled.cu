// The NUTS transition kernels of the built-in models with a diagonal metric (ScaledModel<M>, models.cuh): the opt-in
// mass-matrix path (README.md:66-67 "future updates" of the reference).  Same kernel template, same launch shapes.
#include "nuts_builtin.cuh"

namespace smcb {

long long nuts_ws_bytes_scaled(const Model* m, long long N, int max_depth) {
    switch (m->desc.kind) {
        case kArma: return nuts_ws_bytes<ScaledModel<ArmaModel>>(m, N, max_depth);
        case kPRMwCD:
            return prm_use_group(m->desc, N) ? nuts_ws_bytes<ScaledModel<PrmModelG<kPrmTiles>>>(m, N, max_depth)
                                          : nuts_ws_bytes<ScaledModel<PrmModel>>(m, N, max_depth);
        default: {
#define WS_G(K) nuts_ws_bytes<ScaledModel<GaussModelG<K>>>(m, N, max_depth)
            return SMCB_GAUSS_DISPATCH(m->desc.dim, WS_G, nuts_ws_bytes<ScaledModel<GaussModel>>(m, N, max_depth));
#undef WS_G
        }
    }
}

int launch_nuts_scaled(const Model* m, NutsArgs a, long long workspace_bytes, cudaStream_t st) {
    switch (m->desc.kind) {
        case kArma: return launch_nuts<ScaledModel<ArmaModel>>(m, a, workspace_bytes, st);
        case kPRMwCD:
            return prm_use_group(m->desc, a.N) ? launch_nuts<ScaledModel<PrmModelG<kPrmTiles>>>(m, a, workspace_bytes, st)
                                            : launch_nuts<ScaledModel<PrmModel>>(m, a, workspace_bytes, st);
        default: {
#define LAUNCH_G(K) launch_nuts<ScaledModel<GaussModelG<K>>>(m, a, workspace_bytes, st)
            return SMCB_GAUSS_DISPATCH(m->desc.dim, LAUNCH_G, launch_nuts<ScaledModel<GaussModel>>(m, a, workspace_bytes, st));
#undef LAUNCH_G
        }
    }
}

}  // namespace smcb
