// K2 / K1 kernel templates and their launch helpers, shared by the built-in models (nuts_kernel.cu) and by generated
// model plug-ins (nuts_plugin.cuh): the persistent work-queue NUTS transition kernel and the batched value + gradient
// kernel, instantiated per model struct (interface: models.cuh).
#pragma once
#include <cstdlib>

#include "capi.cuh"
#include "nuts_lane.cuh"

namespace smcb {

// Cap on resident CTAs per SM of the NUTS kernel: the caller's request (NutsArgs::blocks_per_sm, set through
// smcb_nuts_set_blocks_per_sm -- the chunked host path runs several launches side by side, each on a share of every
// SM), else the tuning knob for experiments (tools/quick_time.py), else none.
inline int blocks_per_sm_cap(int requested) {
    if (requested > 0) return requested;
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

template <class M>
static size_t nuts_smem_bytes(const ModelDesc& d) {
    size_t n = (size_t)M::staged_doubles(d);
    if (M::STAGE) n += (size_t)LaunchCfg<M>::NT * nuts_stage_stride(M::STATIC_NL);
    return sizeof(double) * n;
}

template <class M> struct StageOffset { static int of(const ModelDesc&) { return 0; } };
// the diagonal-metric wrapper launches like the model it wraps
template <class M> struct LaunchCfg<ScaledModel<M>> : LaunchCfg<M> {};
template <class M> struct StageOffset<ScaledModel<M>> : StageOffset<M> {};


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

// REJECTED (measured on B200, round 2, gpurun_out/r2j, r2k): a per-warp reserve of claimed queue indices topped up by an
// atomic whose result is read one trip later, plus an L2 prefetch of the claimed rows (meant to take the atomic round trip
// and the DRAM latency of a new particle's rows out of every trip).  arma N = 2^20: 2.86 ms on demand, 4.08 / 3.60 / 3.31 ms
// with chunks of 2 / 4 / 8; the prefetch was neutral.  A lane that finds the reserve empty idles a whole trip, and the
// kernel is not latency- but issue-bound at 4 warps per scheduler (a DFMA occupies the dispatch port for two cycles:
// 0.52 instructions per cycle issued + 0.69 x 0.5 for the FP64 second cycles = 86 % of the port), so hiding the
// refill latency buys nothing there.  Occupancy sweep of the same kernel (1 / 2 / 3 / 4 CTAs per SM): 5.19 / 3.49 /
// 3.02 / 2.90 ms; at 80 registers, 4 / 5 / 6 CTAs: 4.05 / 3.81 / 3.65 ms (spills cost more than the extra warps give).
//
// Tail compaction.  Once the work queue is empty the lanes of a warp finish one after the other, but the warp keeps
// paying full price for every trip until its longest tree ends (an evaluation costs the FP64 pipe the same with 1 or
// 32 live lanes): at N = 2^20 arma particles ~9 % of all evaluated lanes were idle, and the last ~0.5 ms of a 2.9 ms
// launch ran that way (ncu, round 2).  In tail mode the warps of a CTA meet at a barrier every trip; whenever the live
// particles of the CTA fit into fewer warps than currently hold one, the live lane states are packed into the lowest
// warps through a per-CTA exchange area in global memory (a lane's whole state is ~230 bytes; its workspace record is
// referenced by pointer and does not move) and the emptied warps stop evaluating.  Pure scheduling: results are
// independent of it.  MEASURED (B200, tools/ab_time.py, same call): arma N = 2^17 0.784 -> 0.670 ms, N = 2^20 2.915 ->
// 2.858 ms; the 4-lanes-per-particle PRMwCD kernel 126.5 -> 130.7 ms at 2^20 (its tail is the latency of single
// 2047-leapfrog trees, which packing cannot shorten, and the per-trip barrier costs) -- so one-lane models only.
// Not for the models with a per-thread shared staging row (M::STAGE) either.
#ifndef SMCB_TAIL_COMPACT
#define SMCB_TAIL_COMPACT 1
#endif
template <class M> struct TailCfg {
    static constexpr bool ON = (SMCB_TAIL_COMPACT != 0) && M::GROUP == 1 && !M::STAGE && !AlignCfg<M>::ON;
};

template <class M>
static size_t nuts_exchange_bytes(long long blocks) {
    return TailCfg<M>::ON ? (size_t)blocks * LaunchCfg<M>::NT * sizeof(Lane<M>) : 0;
}

template <class M>
__global__ void __launch_bounds__(LaunchCfg<M>::NT, LaunchCfg<M>::MIN_BLOCKS)
nuts_transition_kernel(NutsArgs a, int staged, int stage_offset, int rec_doubles) {
    extern __shared__ double smem[];
    constexpr int G = M::GROUP;
    constexpr int NW = LaunchCfg<M>::NT / 32;
    __shared__ int s_tail;            // number of warps of this CTA that have found the queue empty
    __shared__ int s_cnt[2][NW];      // tail mode: live particles per warp (double-buffered by trip parity)
    if (threadIdx.x == 0) s_tail = 0;
    M model(a.model, stage_model<M>(a.model, smem, staged, stage_offset));
    if constexpr (TailCfg<M>::ON) __syncthreads();
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
    bool reported = false;                       // this warp is counted in s_tail
    for (unsigned trip = 0;; ++trip) {
        if constexpr (TailCfg<M>::ON) {
            // tail mode starts once EVERY warp of the CTA has found the queue empty (warp-uniform test: one shared load)
            if (*(volatile int*)&s_tail == NW) break;
        }
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
            if constexpr (TailCfg<M>::ON) {
                if (!reported && __any_sync(0xffffffffu, drained)) {
                    if (lane_id == 0) atomicAdd(&s_tail, 1);
                    reported = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, lane.phase == kIdle)) {
            if (!AlignCfg<M>::ON || __all_sync(0xffffffffu, drained)) break;
            continue;   // nobody active at a non-admitting trip: the queue is asked again at the next one
        }
        // ---- one model evaluation per particle per trip: the initial point or one leapfrog.  The evaluation is
        //      executed by every lane (idle ones carry zeros) so that warp-wide tensor-core instructions stay legal.
        if (lane.phase != kIdle) { lane.pre_eval(a); lane.prefetch_ck(); }
        double A, B, g[M::NLOC];
        if constexpr (G > 1) __syncwarp();   // the group models issue warp-wide mma.sync.aligned: reconverge explicitly
        model.eval(lane.xa, a.phi, A, B, g);
        lane.take_grad(g);
        if (lane.phase != kIdle) lane.post_eval(a, A, B);
    }
    if constexpr (TailCfg<M>::ON) {
        // ---- tail mode: the queue is empty; every warp of the CTA arrives here within one trip of the first one
        Lane<M>* exch = reinterpret_cast<Lane<M>*>(a.exchange) + (size_t)blockIdx.x * blockDim.x;
        const int warp = threadIdx.x >> 5;
        constexpr int kPerWarp = 32 / G;
        for (unsigned trip = 0;; ++trip) {
            const bool act = lane.phase != kIdle;
            const unsigned b = __ballot_sync(0xffffffffu, act) & kLeaders;
            if (lane_id == 0) s_cnt[trip & 1u][warp] = __popc(b);
            __syncthreads();
            int total = 0, before = 0, have = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int c = s_cnt[trip & 1u][w];
                total += c;
                before += (w < warp) ? c : 0;
                have += (c > 0) ? 1 : 0;
            }
            if (total == 0) break;
            if (have > (total + kPerWarp - 1) / kPerWarp) {   // CTA-uniform: packing frees at least one warp
                if (act) exch[(before + __popc(b & ((1u << group_first) - 1u))) * G + lane.sub] = lane;
                // the new owner reads the lane's workspace record next: everything the old owner stored there (and the
                // state above) is made visible device-wide before the barrier, whatever cache operators the accesses use
                __threadfence();
                __syncthreads();
                if ((int)threadIdx.x < total * G) {
                    double* const stg = lane.stg;
                    lane = exch[threadIdx.x];
                    lane.stg = stg;
                } else {
                    lane.phase = kIdle;
                }
            }
            if (__all_sync(0xffffffffu, lane.phase == kIdle)) continue;   // an emptied warp only keeps the barriers
            if (lane.phase != kIdle) { lane.pre_eval(a); lane.prefetch_ck(); }
            double A, B, g[M::NLOC];
            if constexpr (G > 1) __syncwarp();
            model.eval(lane.xa, a.phi, A, B, g);
            lane.take_grad(g);
            if (lane.phase != kIdle) lane.post_eval(a, A, B);
        }
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


template <class M>
static long long nuts_blocks(const Model* mdl, long long N, size_t smem, int* occ_out, int requested = 0) {
    const int NT = LaunchCfg<M>::NT;
    auto kern = nuts_transition_kernel<M>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem) != cudaSuccess || occ < 1) return -1;
    if (occ > blocks_per_sm_cap(requested)) occ = blocks_per_sm_cap(requested);
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
    return (long long)sizeof(double) * nuts_ws_doubles(probe.nloc(), max_depth) * blocks * LaunchCfg<M>::NT + 256 +
           (long long)nuts_exchange_bytes<M>(blocks);
}

template <class M>
static int launch_nuts(const Model* mdl, NutsArgs a, long long ws_bytes, cudaStream_t st) {
    const int NT = LaunchCfg<M>::NT;
    const int staged = M::staged_doubles(mdl->desc);
    const size_t smem = nuts_smem_bytes<M>(mdl->desc);
    const long long blocks = nuts_blocks<M>(mdl, a.N, smem, nullptr, a.blocks_per_sm);
    if (blocks < 0) return fail("smcb_nuts_transition", "kernel does not fit on an SM");
    M probe(mdl->desc, nullptr);
    const int rec = nuts_ws_doubles(probe.nloc(), a.max_depth, a.g_new != nullptr);
    const long long rec_bytes = (long long)sizeof(double) * rec * blocks * NT;
    const long long ws_need = rec_bytes + 256 + (long long)nuts_exchange_bytes<M>(blocks);
    if (ws_bytes < ws_need) return fail("smcb_nuts_transition", "workspace too small (see smcb_nuts_workspace_bytes)");
    // the queue head lives in the 256 bytes after the lane records, the tail-compaction exchange area after that
    a.queue = (unsigned long long*)((char*)a.ws + rec_bytes);
    a.exchange = (char*)a.ws + rec_bytes + 256;
    SMCB_CUDA(cudaMemsetAsync(a.queue, 0, sizeof(unsigned long long), st));
    nuts_transition_kernel<M><<<(int)blocks, NT, smem, st>>>(a, staged, StageOffset<M>::of(mdl->desc), rec);
    return check_launch("nuts_transition_kernel");
}


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
