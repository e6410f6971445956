// K1 + K2 + K3 kernels and their C-ABI: model handles, batched logp/grad, the persistent work-queue NUTS
// transition.  See nuts_lane.cuh for the per-lane algorithm and models.cuh for the densities.
//
// Launch shape: persistent grid = (#SMs x resident CTAs) so every SM keeps its FP64 pipe fed; each lane
// owns one particle at a time and pulls the next index from a global atomic queue (warp-aggregated)
// when its tree ends, which absorbs the 1..2047-leapfrog raggedness of NUTS trees
// (/root/reference/smcnuts/proposal/nuts.py:50-53 loops particles serially instead).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "capi.cuh"
#include "nuts_lane.cuh"

namespace smcb {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
std::atomic<long long> g_launches{0};

int device_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Tuning knob for experiments (tools/quick_time.py): cap on resident CTAs per SM of the NUTS kernel.
static int blocks_per_sm_cap() {
    const char* e = getenv("SMCB_NUTS_BLOCKS_PER_SM");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 1 << 20;
}

// NT threads per CTA, MIN_BLOCKS resident CTAs/SM (register budget).
// MEASURED (round 1, profiles/README.md): more resident warps do not help arma (94 registers / 5 CTAs: same time,
// 64 registers / 8 CTAs: 12 % slower), and keeping the low slots of the per-lane tree workspace in shared memory made
// the kernel 37 % SLOWER -- the carve-out leaves almost no L1, and the L1 already serves the workspace and
// particle-row traffic.  The per-lane records therefore stay in global memory (L1/L2 resident).
template <class M> struct LaunchCfg { static constexpr int NT = 256, MIN_BLOCKS = 1; };   // GaussModelG: one CTA/SM, one copy of the B fragments
template <> struct LaunchCfg<ArmaModel> { static constexpr int NT = 128, MIN_BLOCKS = 4; };
template <> struct LaunchCfg<PrmModel> { static constexpr int NT = 128, MIN_BLOCKS = 2; };
template <> struct LaunchCfg<GaussModel> { static constexpr int NT = 128, MIN_BLOCKS = 1; };
// PrmModelG, MEASURED (N = 2^20): 4 CTAs/SM at 118 registers 130.5 ms; 3 CTAs/SM 140+ ms; 5 CTAs/SM (96 registers, spills) 132-152 ms
template <int T8> struct LaunchCfg<PrmModelG<T8>> { static constexpr int NT = 128, MIN_BLOCKS = 4; };
constexpr int kPrmTiles = 13;   // PrmModelG instantiation: 81..104 observations (the shipped PRMwCD has 100)

// staged model data, then (M::STAGE) one staging row per thread for the stored edge of the U-turn tests
// Parity alignment of the refill.  A particle that starts at trip t0 stores a leaf (first leaf of a two-leaf sub-tree)
// at trips t0+2, t0+4, ... and merges / ends doublings at the trips in between: every doubling after the first has an
// even number of leaves.  When the particles of a warp start at trips of mixed parity, every trip executes BOTH
// divergent paths of the lane bookkeeping; when new particles are only admitted at even trips, all particles of a warp
// store on even trips and merge on odd ones, and a trip executes one path.  The price is one idle trip for half of
// the refills (1 % of the work at 50 leapfrogs per particle); results do not depend on the lane assignment.
#ifndef SMCB_ALIGN_PARITY
#define SMCB_ALIGN_PARITY 0
#endif
template <class M> struct AlignCfg { static constexpr bool ON = (SMCB_ALIGN_PARITY != 0) && M::GROUP > 1; };
#if SMCB_ALIGN_PARITY == 2   // A/B experiments: also for the one-lane-per-particle kernels
template <> struct AlignCfg<ArmaModel> { static constexpr bool ON = true; };
template <> struct AlignCfg<PrmModel> { static constexpr bool ON = true; };
#endif

template <class M>
static size_t nuts_smem_bytes(const ModelDesc& d) {
    size_t n = (size_t)M::staged_doubles(d);
    if (M::STAGE) n += (size_t)LaunchCfg<M>::NT * nuts_stage_stride(M::STATIC_NL);
    return sizeof(double) * n;
}

template <class M> struct StageOffset { static int of(const ModelDesc&) { return 0; } };
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

// Model data (y[200]; the PRMwCD table or its tensor-core fragments; the Gaussian B-fragments) is staged once per CTA into shared memory,
// where every lane reads the same address each step (broadcast / conflict-free).  The plain Gaussian precision
// matrix of the one-lane-per-particle fallback stays in L1/L2.
template <class M>
__device__ __forceinline__ const double* stage_model(const ModelDesc& d, double* smem, int staged, int offset) {
    if constexpr (M::STATIC_NL != 0) {
        for (int i = threadIdx.x; i < staged; i += blockDim.x) smem[i] = d.data[offset + i];
        __syncthreads();
        return smem;
    } else {
        return d.data;
    }
}

template <class M>
__global__ void __launch_bounds__(LaunchCfg<M>::NT, LaunchCfg<M>::MIN_BLOCKS)
nuts_transition_kernel(NutsArgs a, int staged, int stage_offset, int rec_doubles) {
    extern __shared__ double smem[];
    constexpr int G = M::GROUP;
    M model(a.model, stage_model<M>(a.model, smem, staged, stage_offset));
    const unsigned lane_id = threadIdx.x & 31u;
    Lane<M> lane;
    lane.idle_init(model, (int)(lane_id % G));
    lane.stg = M::STAGE ? smem + staged + (size_t)threadIdx.x * nuts_stage_stride(M::STATIC_NL) : nullptr;
    double* ws = a.ws + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * rec_doubles;
    constexpr unsigned kLeaders = G == 1 ? 0xffffffffu : 0x11111111u;   // first lane of every particle group
    const unsigned group_first = lane_id & ~(unsigned)(G - 1);
    bool drained = false;
    // with gradient carry-over the tree starts in the trip a particle is admitted, one trip earlier than otherwise
    const unsigned admit_parity = a.g_in ? 1u : 0u;
    for (unsigned trip = 0;; ++trip) {
        // ---- refill finished particle groups from the work queue (warp-aggregated atomic)
        const bool admit = !AlignCfg<M>::ON || ((trip & 1u) == admit_parity);
        const bool want = (lane.phase == kIdle) && !drained && admit;
        const unsigned m = __ballot_sync(0xffffffffu, want) & kLeaders;
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if ((int)lane_id == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                const long long p = (long long)base + __popc(m & ((1u << group_first) - 1u));
                if (p < a.N) lane.begin(a, model, p, ws);
                else drained = true;
            }
        }
        if (__all_sync(0xffffffffu, lane.phase == kIdle)) {
            if (!AlignCfg<M>::ON || __all_sync(0xffffffffu, drained)) break;
            continue;   // nobody active at a non-admitting trip: the queue is asked again at the next one
        }
        // ---- one model evaluation per particle per trip: the initial point or one leapfrog.  The evaluation is
        //      executed by every lane (idle ones carry zeros) so that warp-wide tensor-core instructions stay legal.
        if (lane.phase != kIdle) lane.pre_eval(a);
        double A, B, g[M::NLOC];
        if constexpr (G > 1) __syncwarp();   // the group models issue warp-wide mma.sync.aligned: reconverge explicitly
        model.eval(lane.xa, a.phi, A, B, g);
        lane.take_grad(g);
        if (lane.phase != kIdle) lane.post_eval(a, A, B);
    }
}

// Batched value + gradient (one thread per particle).
template <class M>
__global__ void __launch_bounds__(128) logp_grad_kernel(ModelDesc md, const double* __restrict__ x, long long N,
                                                        double phi, double* __restrict__ Aout,
                                                        double* __restrict__ Bout, double* __restrict__ grad,
                                                        int staged) {
    extern __shared__ double smem[];
    M model(md, stage_model<M>(md, smem, staged, 0));
    const int D = M::STATIC_D ? M::STATIC_D : model.dim();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double xv[M::DMAX], g[M::DMAX], A, B;
#pragma unroll
        for (int d = 0; d < (M::STATIC_D ? M::STATIC_D : D); ++d) xv[d] = x[i * D + d];
        model.eval(xv, phi, A, B, g);
        if (Aout) Aout[i] = A;
        if (Bout) Bout[i] = B;
        if (grad) {
            const bool bad = !is_finite(A + phi * B);
#pragma unroll
            for (int d = 0; d < (M::STATIC_D ? M::STATIC_D : D); ++d) grad[i * D + d] = bad ? neg_inf() : g[d];
        }
    }
}

__global__ void combine_logp_kernel(const double* __restrict__ A, const double* __restrict__ B, double phi,
                                    long long N, double* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double lp = A[i] + phi * B[i];
        out[i] = is_finite(lp) ? lp : neg_inf();
    }
}

template <class M>
static long long nuts_blocks(const Model* mdl, long long N, size_t smem, int* occ_out) {
    const int NT = LaunchCfg<M>::NT;
    auto kern = nuts_transition_kernel<M>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem) != cudaSuccess || occ < 1) return -1;
    if (occ > blocks_per_sm_cap()) occ = blocks_per_sm_cap();
    long long blocks = (long long)device_sm_count() * occ;
    const long long need = (N * M::GROUP + NT - 1) / NT;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    if (occ_out) *occ_out = occ;
    (void)mdl;
    return blocks;
}

template <class M>
static long long nuts_ws_bytes(const Model* mdl, long long N, int max_depth) {
    const size_t smem = nuts_smem_bytes<M>(mdl->desc);
    const long long blocks = nuts_blocks<M>(mdl, N, smem, nullptr);
    if (blocks < 0) return -1;
    M probe(mdl->desc, nullptr);
    return (long long)sizeof(double) * nuts_ws_doubles(probe.nloc(), max_depth) * blocks * LaunchCfg<M>::NT + 256;
}

template <class M>
static int launch_nuts(const Model* mdl, NutsArgs a, long long ws_bytes, cudaStream_t st) {
    const int NT = LaunchCfg<M>::NT;
    const int staged = M::staged_doubles(mdl->desc);
    const size_t smem = nuts_smem_bytes<M>(mdl->desc);
    const long long blocks = nuts_blocks<M>(mdl, a.N, smem, nullptr);
    if (blocks < 0) return fail("smcb_nuts_transition", "kernel does not fit on an SM");
    M probe(mdl->desc, nullptr);
    const int rec = nuts_ws_doubles(probe.nloc(), a.max_depth, a.g_new != nullptr);
    const long long ws_need = (long long)sizeof(double) * rec * blocks * NT + 256;
    if (ws_bytes < ws_need) return fail("smcb_nuts_transition", "workspace too small (see smcb_nuts_workspace_bytes)");
    // queue head lives in the last 256 bytes of the workspace
    a.queue = (unsigned long long*)((char*)a.ws + (ws_need - 256));
    SMCB_CUDA(cudaMemsetAsync(a.queue, 0, sizeof(unsigned long long), st));
    nuts_transition_kernel<M><<<(int)blocks, NT, smem, st>>>(a, staged, StageOffset<M>::of(mdl->desc), rec);
    return check_launch("nuts_transition_kernel");
}

// Gaussian: tensor-core group kernel for D <= 104, one-lane-per-particle fallback above (SMCB_GAUSS_SCALAR=1 forces the
// fallback: parity test of the two)
static bool gauss_force_scalar() {
    const char* e = getenv("SMCB_GAUSS_SCALAR");
    return e && atoi(e) != 0;
}
#define SMCB_GAUSS_DISPATCH(D, CALL_G, CALL_PLAIN) \
    (gauss_force_scalar() ? CALL_PLAIN : (D) <= 8 ? CALL_G(1) : (D) <= 16 ? CALL_G(2) : (D) <= 32 ? CALL_G(4) : (D) <= 64 ? CALL_G(8) : (D) <= 104 ? CALL_G(13) : CALL_PLAIN)

template <class M>
static int launch_logp(const Model* mdl, const double* x, long long N, double phi, double* A, double* B, double* g,
                       cudaStream_t st) {
    const int staged = M::staged_doubles(mdl->desc);
    const size_t smem = sizeof(double) * (size_t)staged;
    const int grid = stride_grid(N, 128, 8);
    logp_grad_kernel<M><<<grid, 128, smem, st>>>(mdl->desc, x, N, phi, A, B, g, staged);
    return check_launch("logp_grad_kernel");
}

}  // namespace smcb

using namespace smcb;

extern "C" {

int smcb_version(void) { return 100; }
int smcb_build_flavour(void) { return SMCB_PARITY; }
const char* smcb_last_error(void) { return last_error_ref().c_str(); }
long long smcb_launch_count(void) { return g_launches.load(); }

int smcb_model_create(int kind, const double* host_data, long long n, int dim, void** handle) {
    SMCB_REQUIRE(handle && host_data && n > 0, "null argument");
    std::vector<double> packed;
    ModelDesc d{};
    d.kind = kind;
    if (kind == SMCB_MODEL_ARMA) {
        d.dim = 4; d.T = (int)n;
        packed.assign(host_data, host_data + n);
    } else if (kind == SMCB_MODEL_PRMWCD) {
        SMCB_REQUIRE((n - 1) % 13 == 0, "PRMwCD blob must be [q, y(NO), lgamma(y+1)(NO), X(NO*11)]");
        const int NO = (int)((n - 1) / 13);
        d.dim = 13; d.T = NO; d.q = host_data[0];
        const double *y = host_data + 1, *lg = y + NO, *X = lg + NO;
        packed.assign((size_t)PrmModel::HDR + (size_t)NO * PrmModel::ROW, 0.0);
        for (int i = 0; i < NO; ++i) {
            double* row = packed.data() + PrmModel::HDR + (size_t)i * PrmModel::ROW;
            for (int j = 0; j < 11; ++j) row[j] = X[i * 11 + j];
            row[11] = y[i];
            packed[0] += y[i];
            for (int j = 0; j < 11; ++j) packed[1 + j] += y[i] * X[i * 11 + j];
            packed[12] += lg[i];
        }
        if (PrmModelG<kPrmTiles>::fits(d)) {   // staged block of the tensor-core NUTS kernel, appended to the scalar blob
            const size_t n0 = packed.size();
            packed.resize(n0 + PrmModelG<kPrmTiles>::TOTAL);
            pack_prm_fragments(packed.data(), NO, kPrmTiles, packed.data() + n0);
        }
    } else if (kind == SMCB_MODEL_GAUSS) {
        SMCB_REQUIRE(dim >= 1 && dim <= GaussModel::DMAX && (long long)dim * dim == n, "gauss: need P[D*D], D <= 128");
        d.dim = dim;
        packed.assign(host_data, host_data + n);
        if (dim <= 104) {   // B-fragment packing for the tensor-core NUTS kernel, appended to the plain matrix
            const int nt8 = dim <= 8 ? 1 : dim <= 16 ? 2 : dim <= 32 ? 4 : dim <= 64 ? 8 : 13;
            packed.resize((size_t)n + (size_t)nt8 * 2 * nt8 * 32);
            pack_gauss_fragments(host_data, dim, nt8, packed.data() + n);
        }
    } else {
        return fail("smcb_model_create", "unknown model kind");
    }
    d.n_data = (int)packed.size();
    Model* m = new Model{d, nullptr};
    if (cudaMalloc(&m->d_data, sizeof(double) * packed.size()) != cudaSuccess) {
        delete m;
        return fail("smcb_model_create", "cudaMalloc failed (is a CUDA device present?)");
    }
    SMCB_CUDA(cudaMemcpy(m->d_data, packed.data(), sizeof(double) * packed.size(), cudaMemcpyHostToDevice));
    m->desc.data = m->d_data;
    *handle = m;
    return 0;
}

int smcb_model_destroy(void* handle) {
    if (!handle) return 0;
    Model* m = (Model*)handle;
    cudaFree(m->d_data);
    delete m;
    return 0;
}

int smcb_model_dim(void* handle) { return handle ? ((Model*)handle)->desc.dim : -1; }

int smcb_debug_pack_prm(const double* host_scalar_blob, int n_obs, int tiles, double* host_out, long long n_out) {
    SMCB_REQUIRE(host_scalar_blob && host_out && n_obs > 0 && tiles > 0 && n_obs <= 8 * tiles, "bad argument");
    SMCB_REQUIRE(n_out >= 32 + (long long)tiles * 9 * 32, "output too small: 32 + tiles * 288 doubles");
    pack_prm_fragments(host_scalar_blob, n_obs, tiles, host_out);
    return 0;
}

int smcb_logp_grad(void* handle, const double* x, long long N, double phi, double* A, double* B, double* grad,
                   void* stream) {
    SMCB_REQUIRE(handle && x && N >= 0, "bad argument");
    if (N == 0) return 0;
    const Model* m = (const Model*)handle;
    cudaStream_t st = (cudaStream_t)stream;
    switch (m->desc.kind) {
        case kArma: return launch_logp<ArmaModel>(m, x, N, phi, A, B, grad, st);
        case kPRMwCD: return launch_logp<PrmModel>(m, x, N, phi, A, B, grad, st);
        default: return launch_logp<GaussModel>(m, x, N, phi, A, B, grad, st);
    }
}

int smcb_combine_logp(const double* A, const double* B, double phi, long long N, double* out, void* stream) {
    SMCB_REQUIRE(A && B && out && N >= 0, "bad argument");
    if (N == 0) return 0;
    combine_logp_kernel<<<stride_grid(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(A, B, phi, N, out);
    return check_launch("combine_logp_kernel");
}

int smcb_nuts_workspace_bytes(void* handle, long long N, int max_depth, long long* bytes) {
    SMCB_REQUIRE(handle && bytes && N >= 0, "bad argument");
    SMCB_REQUIRE(max_depth >= 1 && max_depth <= 10, "max_depth must be in [1, 10] (reference: MAX_TREE_DEPTH = 10)");
    const Model* m = (const Model*)handle;
    long long b;
    switch (m->desc.kind) {
        case kArma: b = nuts_ws_bytes<ArmaModel>(m, N, max_depth); break;
        case kPRMwCD:
            b = prm_use_group(m->desc, N) ? nuts_ws_bytes<PrmModelG<kPrmTiles>>(m, N, max_depth) : nuts_ws_bytes<PrmModel>(m, N, max_depth);
            break;
        default: {
#define WS_G(K) nuts_ws_bytes<GaussModelG<K>>(m, N, max_depth)
            b = SMCB_GAUSS_DISPATCH(m->desc.dim, WS_G, nuts_ws_bytes<GaussModel>(m, N, max_depth));
#undef WS_G
            break;
        }
    }
    if (b < 0) return fail("smcb_nuts_workspace_bytes", "occupancy query failed");
    *bytes = b;
    return 0;
}

int smcb_nuts_transition(void* handle, const double* x, const double* r, long long N, double eps, double phi,
                         int max_depth, int accrej, uint64_t seed, uint32_t iteration, uint64_t particle0,
                         double* x_new, double* r_new, double* A_old, double* B_old, double* A_new, double* B_new,
                         double* ke_old, double* ke_new, int* n_leapfrog, int* accepted, int* depth, double* accept_stat,
                         const double* A_in, const double* B_in, const double* g_in, double* g_new, void* workspace,
                         long long workspace_bytes, void* stream) {
    SMCB_REQUIRE(handle && x && r && x_new && r_new && workspace, "null argument");
    SMCB_REQUIRE((A_in && B_in && g_in) || (!A_in && !B_in && !g_in), "A_in, B_in, g_in: all three or none");
    SMCB_REQUIRE(!(accrej && (g_in || g_new)), "gradient carry-over is not available with the accept-reject epilogue");
    SMCB_REQUIRE(g_new != x_new && (!g_in || (g_in != g_new)), "g_new must not alias its inputs");
    SMCB_REQUIRE(x != x_new && r != r_new, "x_new/r_new must not alias x/r");
    SMCB_REQUIRE(max_depth >= 1 && max_depth <= 10, "max_depth must be in [1, 10] (reference: MAX_TREE_DEPTH = 10)");
    SMCB_REQUIRE(iteration < (1u << 24), "iteration must be < 2^24");
    if (N <= 0) return N == 0 ? 0 : fail("smcb_nuts_transition", "negative N");
    const Model* m = (const Model*)handle;
    NutsArgs a{};
    a.model = m->desc;
    a.x = x; a.r = r; a.N = N; a.eps = eps; a.phi = phi; a.max_depth = max_depth; a.accrej = accrej;
    a.seed = seed; a.iteration = iteration; a.particle0 = particle0;
    a.x_new = x_new; a.r_new = r_new; a.A_old = A_old; a.B_old = B_old; a.A_new = A_new; a.B_new = B_new;
    a.ke_old = ke_old; a.ke_new = ke_new; a.n_leapfrog = n_leapfrog; a.accepted = accepted; a.depth = depth;
    a.accept_stat = accept_stat;
    a.A_in = A_in; a.B_in = B_in; a.g_in = g_in; a.g_new = g_new;
    a.ws = (double*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    switch (m->desc.kind) {
        case kArma: return launch_nuts<ArmaModel>(m, a, workspace_bytes, st);
        case kPRMwCD:
            return prm_use_group(m->desc, N) ? launch_nuts<PrmModelG<kPrmTiles>>(m, a, workspace_bytes, st)
                                          : launch_nuts<PrmModel>(m, a, workspace_bytes, st);
        default: {
#define LAUNCH_G(K) launch_nuts<GaussModelG<K>>(m, a, workspace_bytes, st)
            return SMCB_GAUSS_DISPATCH(m->desc.dim, LAUNCH_G, launch_nuts<GaussModel>(m, a, workspace_bytes, st));
#undef LAUNCH_G
        }
    }
}

}  // extern "C"
