// K9, K10, K11 -- resampling: block-scan prefix sum of the normalised weights, ancestor search,
// coalesced particle gather.
//
// Replaces `rng.choice(arange(N), N, p=wn)` + `x[i_new]` (/root/reference/smcnuts/samples/samples.py:138-140,
// estimate_from_tempered.py:42-44).  numpy's choice is cdf = cumsum(p); cdf /= cdf[-1];
// searchsorted(cdf, uniforms, side='right') -- verified bit-identical in tests/golden/make_golden.py --
// so given the same cdf and the same uniforms the ancestor indices here are bit-exact.
// The scan is reduce-then-scan over 2048-element tiles with a fixed summation tree (deterministic).
#include <cstring>

#include "capi.cuh"
#include "philox.cuh"

namespace smcb {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kTile = kScanThreads * kScanItems;

__device__ __forceinline__ double warp_incl_scan(double v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// inclusive block scan of one value per thread; returns the thread's inclusive prefix and the block total
__device__ __forceinline__ double block_incl_scan(double v, double* total) {
    __shared__ double wsum[kScanThreads / 32];
    __shared__ double btotal;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double incl = warp_incl_scan(v);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = (lane < kScanThreads / 32) ? wsum[lane] : 0.0;
        const double wi = warp_incl_scan(w);
        if (lane < kScanThreads / 32) wsum[lane] = wi - w;  // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1) btotal = wi;
    }
    __syncthreads();
    const double r = incl + wsum[warp];
    *total = btotal;
    __syncthreads();
    return r;
}

// phase 1: per-tile sums.  Threads own kScanItems CONSECUTIVE elements (blocked arrangement via smem transpose).
__global__ void __launch_bounds__(kScanThreads) tile_sum_kernel(const double* __restrict__ w, long long N,
                                                                 double* __restrict__ tile_sums, long long ntiles) {
    __shared__ double sm[kTile + kTile / 32];
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * kTile;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = k * kScanThreads + threadIdx.x;
            sm[e + e / 32] = (base + e < N) ? w[base + e] : 0.0;
        }
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = threadIdx.x * kScanItems + k;
            s += sm[e + e / 32];
        }
        double total;
        block_incl_scan(s, &total);
        if (threadIdx.x == 0) tile_sums[tile] = total;
    }
}

// phase 2: exclusive scan of the tile sums by one block (sequential carry across chunks), grand total
__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(double* tile_sums, long long ntiles,
                                                                  const double* __restrict__ offset_in,
                                                                  const double* __restrict__ total_in,
                                                                  double* total_out, double* norm_total) {
    double carry = offset_in ? *offset_in : 0.0;
    const double carry0 = carry;
    for (long long c0 = 0; c0 < ntiles; c0 += kScanThreads) {
        const long long i = c0 + threadIdx.x;
        const double v = (i < ntiles) ? tile_sums[i] : 0.0;
        double total;
        const double incl = block_incl_scan(v, &total);
        if (i < ntiles) tile_sums[i] = carry + (incl - v);
        carry += total;
    }
    if (threadIdx.x == 0) {
        total_out[0] = carry - carry0;                    // this rank's local weight total
        norm_total[0] = total_in ? *total_in : carry;     // normaliser: global total when sharded
    }
}

// phase 2, wide: ONE pass of a 1024-thread block over all tile sums (thread t owns a contiguous run of them: sequential
// sum, block scan of the 1024 run totals, sequential write-back of the exclusive offsets).  Replaces the 256-wide chunked
// loop above, whose ntiles/256 dependent round trips (64 at 2^25 particles) were a third of the whole scan.
constexpr int kWideThreads = 1024;
__global__ void __launch_bounds__(kWideThreads) tile_scan_wide_kernel(double* tile_sums, long long ntiles,
                                                                       const double* __restrict__ offset_in,
                                                                       const double* __restrict__ total_in,
                                                                       double* total_out, double* norm_total) {
    __shared__ double wsum[kWideThreads / 32];
    __shared__ double btotal;
    const long long ipt = (ntiles + kWideThreads - 1) / kWideThreads;
    const long long lo = min(ntiles, (long long)threadIdx.x * ipt), hi = min(ntiles, lo + ipt);
    double s = 0.0;
    for (long long i = lo; i < hi; ++i) s += tile_sums[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double incl = warp_incl_scan(s);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const double w = wsum[lane];
        const double wi = warp_incl_scan(w);
        wsum[lane] = wi - w;
        if (lane == 31) btotal = wi;
    }
    __syncthreads();
    const double carry0 = offset_in ? *offset_in : 0.0;
    double off = carry0 + (wsum[warp] + (incl - s));
    for (long long i = lo; i < hi; ++i) {
        const double v = tile_sums[i];
        tile_sums[i] = off;
        off += v;
    }
    if (threadIdx.x == 0) {
        total_out[0] = btotal;                               // this rank's local weight total
        norm_total[0] = total_in ? *total_in : carry0 + btotal;   // normaliser: global total when sharded
    }
}

// samples.py:101-102 fused with phase 1 of the scan: wn = exp(logw - logZ) (0 where logw = -inf) is written AND summed
// per tile in the same pass, so the weights are read from HBM once for normalisation and tile sums together
// (lse + normalise + scan: 8 + 16 + 16 bytes per particle in total, the algorithmic figure).
__global__ void __launch_bounds__(kScanThreads) normalise_tile_sum_kernel(const double* __restrict__ logw, long long N,
                                                                           const double* __restrict__ logZ,
                                                                           double* __restrict__ wn,
                                                                           double* __restrict__ tile_sums,
                                                                           long long ntiles) {
    __shared__ double sm[kTile + kTile / 32];
    const double z = *logZ;
    // software-pipelined over tiles: the next tile's weights are loaded before this tile's exps, stores and block scan
    double xn[kScanItems];
    long long tile = blockIdx.x;
    if (tile < ntiles) {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const long long i = tile * kTile + k * kScanThreads + threadIdx.x;
            xn[k] = (i < N) ? logw[i] : neg_inf();
        }
    }
    for (; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * kTile;
        double x[kScanItems], w[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) x[k] = xn[k];
        const long long next = tile + gridDim.x;
        if (next < ntiles) {
#pragma unroll
            for (int k = 0; k < kScanItems; ++k) {
                const long long i = next * kTile + k * kScanThreads + threadIdx.x;
                xn[k] = (i < N) ? logw[i] : neg_inf();
            }
        }
#pragma unroll
        for (int k = 0; k < kScanItems; k += 2) {   // exps in interleaved pairs (two short dependency chains each)
            double e0, e1;
            fast_exp_pair(x[k] - z, x[k + 1] - z, e0, e1);
            w[k] = (x[k] == neg_inf()) ? 0.0 : e0;
            w[k + 1] = (x[k + 1] == neg_inf()) ? 0.0 : e1;
        }
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = k * kScanThreads + threadIdx.x;
            if (base + e < N) wn[base + e] = w[k];
            sm[e + e / 32] = w[k];
        }
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = threadIdx.x * kScanItems + k;
            s += sm[e + e / 32];
        }
        double total;
        block_incl_scan(s, &total);
        if (threadIdx.x == 0) tile_sums[tile] = total;
    }
}

// phase 3: scan each tile, add its offset, normalise: cdf = prefix / total  (numpy: cdf /= cdf[-1])
__global__ void __launch_bounds__(kScanThreads) tile_cdf_kernel(const double* __restrict__ w, long long N,
                                                                 const double* __restrict__ tile_offsets,
                                                                 const double* __restrict__ norm_total,
                                                                 const double* __restrict__ extra_offset,
                                                                 double* __restrict__ cdf, long long ntiles) {
    __shared__ double sm[kTile + kTile / 32];
    const double tot = extra_offset ? extra_offset[1] : *norm_total;   // (rank offset, global total) when sharded late
    const double xoff = extra_offset ? extra_offset[0] : 0.0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * kTile;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = k * kScanThreads + threadIdx.x;
            sm[e + e / 32] = (base + e < N) ? w[base + e] : 0.0;
        }
        __syncthreads();
        double v[kScanItems], s = 0.0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = threadIdx.x * kScanItems + k;
            s += sm[e + e / 32];
            v[k] = s;
        }
        double total;
        const double incl = block_incl_scan(s, &total);
        const double off = (xoff + tile_offsets[tile]) + (incl - s);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = threadIdx.x * kScanItems + k;
            sm[e + e / 32] = (off + v[k]) / tot;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = k * kScanThreads + threadIdx.x;
            if (base + e < N) cdf[base + e] = sm[e + e / 32];
        }
        __syncthreads();
    }
}

// first index with cdf[idx] > u  (numpy searchsorted side='right'), clamped to N-1
__device__ __forceinline__ long long upper_bound(const double* __restrict__ cdf, long long N, double u) {
    long long lo = 0, hi = N;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (u < cdf[mid]) hi = mid;
        else lo = mid + 1;
    }
    return lo < N ? lo : N - 1;
}

__global__ void ancestors_multinomial_kernel(const double* __restrict__ cdf, long long N, const double* __restrict__ u,
                                             long long M, int64_t* __restrict__ idx) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x)
        idx[j] = upper_bound(cdf, N, u[j]);
}

__global__ void ancestors_systematic_kernel(const double* __restrict__ cdf, long long N, double u0, long long j0,
                                            long long M_total, long long M, int64_t* __restrict__ idx) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
        const double pos = ((double)(j0 + j) + u0) / (double)M_total;
        idx[j] = upper_bound(cdf, N, pos);
    }
}

// out[j, :] = x[idx[j], :].  VEC doubles per thread (16-byte accesses when D is even).
template <int VEC>
__global__ void gather_rows_kernel(const double* __restrict__ x, const int64_t* __restrict__ idx, long long M, int D,
                                   double* __restrict__ out) {
    const int cpr = D / VEC;  // chunks per row
    const long long total = M * cpr;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long j = t / cpr;
        const int c = (int)(t - j * cpr);
        const long long src = idx[j];
        if (VEC == 2) {
            const double2 v = *reinterpret_cast<const double2*>(x + src * D + 2 * c);
            *reinterpret_cast<double2*>(out + j * D + 2 * c) = v;
        } else {
            out[j * D + c] = x[src * D + c];
        }
    }
}

// Systematic ancestors + gather fused (no index round trip through HBM).  Positions are sorted, so the ancestors of
// one tile of output rows lie in one contiguous cdf range [bounds[t], bounds[t+1]]:
//   pass 1 (tile_bounds_kernel): one full binary search per TILE, all tiles in parallel (latency hidden by parallelism);
//   pass 2 (resample_systematic_fused_kernel): warps take tiles in grid-stride order (balanced, address-local), every
//   row searches only inside its tile's (L1-resident) range and copies its ancestor's row with 16-byte accesses.
__global__ void tile_bounds_kernel(const double* __restrict__ cdf, long long N, double u0,
                                   const double* __restrict__ u0_dev, long long j0, long long M_total, long long M,
                                   int rows_per_tile, long long ntiles, long long* __restrict__ bounds) {
    if (u0_dev) u0 = *u0_dev;
    const double den = (double)M_total;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t <= ntiles; t += (long long)gridDim.x * blockDim.x) {
        const long long j = min(M - 1, t * rows_per_tile);   // first row of tile t (last row overall for t == ntiles)
        bounds[t] = upper_bound(cdf, N, ((double)(j0 + j) + u0) / den);
    }
}

template <int LPR>
__global__ void __launch_bounds__(256) resample_systematic_fused_kernel(const double* __restrict__ cdf, long long N,
                                                                        double u0, const double* __restrict__ u0_dev,
                                                                        long long j0, long long M_total,
                                                                        long long M, const double* __restrict__ x,
                                                                        double* __restrict__ out,
                                                                        int64_t* __restrict__ idx,
                                                                        const long long* __restrict__ bounds,
                                                                        double* const* __restrict__ peer_out,
                                                                        long long rows_per_rank) {
    constexpr int D = 2 * LPR, GROUPS = 32 / LPR, ROWS = 32;   // 32 rows per warp tile: lane l owns the search of row l
    const int lane = threadIdx.x & 31, sub = lane % LPR, g = lane / LPR;
    if (u0_dev) u0 = *u0_dev;
    const double den = (double)M_total;
    const long long ntiles = (M + ROWS - 1) / ROWS;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // Push variant: tiles are visited in a scattered order ((i * prime) mod ntiles, a bijection).  The slots a rank serves
    // are contiguous, so in plain grid-stride order every warp is in the same destination's sub-range at the same time
    // and the launch alternates between an NVLink-bound phase (HBM idle) and an HBM-bound one (NVLink idle): measured at
    // 2 GPUs, time = local time + remote bytes / 840 GB/s.  Scattered, the resident warps always hold a mix of local and
    // remote tiles and the two overlap.
    // Runs of 16 tiles (512 rows) stay together: neighbouring tiles share cdf lines.
    constexpr long long kPrime = 1000003;
    const long long nruns = ntiles >> 4;
    const bool scatter = peer_out != nullptr && nruns > 1 && (nruns % kPrime) != 0;
    for (long long it = wid; it < ntiles; it += nwarps) {
        const long long run = it >> 4;
        const long long tile = (scatter && run < nruns) ? ((((run * kPrime) % nruns) << 4) | (it & 15)) : it;
        const long long jlo = tile * ROWS;
        const long long blo = bounds[tile], bhi = bounds[tile + 1];   // every ancestor of the tile lies in [blo, bhi]
        const long long jmine = jlo + lane;
        const double pos = ((double)(j0 + jmine) + u0) / den;
        long long anc = blo;
        const long long range = bhi - blo;   // entries blo .. bhi-1 may be <= pos
        if (range <= 2048) {
            // warp-cooperative rank: ancestor(row) = blo + #{i in [blo, bhi) : cdf[i] <= pos(row)}.  The cdf chunk is
            // loaded once, coalesced; ranks come from ballots -- no dependent chain of loads per row.
            int cnt = 0;
            for (long long c0 = 0; c0 < range; c0 += 32) {
                const long long i = blo + c0 + lane;
                const double cv = (i < bhi) ? cdf[i] : 1e300;
#pragma unroll 8
                for (int rrow = 0; rrow < 32; ++rrow) {
                    const double pr = __shfl_sync(0xffffffffu, pos, rrow);
                    const unsigned b = __ballot_sync(0xffffffffu, !(pr < cv));
                    if (lane == rrow) cnt += __popc(b);
                }
            }
            anc = blo + cnt;
        } else if (jmine < M) {
            long long lo = blo, hi = bhi;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (pos < cdf[mid]) hi = mid;
                else lo = mid + 1;
            }
            anc = lo;
        }
        if (anc > N - 1) anc = N - 1;
        if (idx && jmine < M) idx[jmine] = anc;
#pragma unroll
        for (int k = 0; k < 32 / GROUPS; ++k) {
            const int rrow = g + k * GROUPS;
            const long long a = __shfl_sync(0xffffffffu, anc, rrow);
            const long long j = jlo + rrow;
            if (j < M) {
                const double2 v = *reinterpret_cast<const double2*>(x + a * D + 2 * sub);
                if (peer_out) {
                    // fused migration: the row goes straight into the destination rank's buffer over NVLink (peer
                    // store); slot j0 + j of the global output belongs to rank (j0 + j) / rows_per_rank
                    const long long gj = j0 + j;
                    const long long dst = gj / rows_per_rank;
                    *reinterpret_cast<double2*>(peer_out[dst] + (gj - dst * rows_per_rank) * D + 2 * sub) = v;
                } else {
                    *reinterpret_cast<double2*>(out + j * D + 2 * sub) = v;
                }
            }
        }
    }
}

// Sharded multinomial resampling, owner-push formulation (samples.py:138-140 over P GPUs).
// Output slot j of the GLOBAL particle set draws u_j from the Philox resampling stream keyed by j, so every rank can
// regenerate every uniform.  The rank whose cdf segment contains u_j (ends[q-1] <= u_j < ends[q], ends = last cdf value of
// each rank) is the only one that can resolve the ancestor: it searches its local segment and stores the ancestor's row
// straight into the buffer of the rank that owns slot j (peer store over NVLink), together with the global ancestor
// index.  No cdf all-gather, no sort / bincount, no request-response all-to-all: N Philox draws + compares per rank
// (cheap, ~60 integer instructions each), then only the hits do memory work.
__global__ void __launch_bounds__(256) resample_multinomial_push_kernel(
    const double* __restrict__ cdf, long long n, const double* __restrict__ ends, int rank, int P, uint64_t seed,
    uint32_t iteration, uint32_t stream_id, long long N_total, long long particle0, const double* __restrict__ x, int D,
    double* const* __restrict__ peer_out, int64_t* const* __restrict__ peer_idx, long long rows_per_rank) {
    const double lo = rank == 0 ? neg_inf() : ends[rank - 1];
    const double hi = rank == P - 1 ? -neg_inf() : ends[rank];
    const bool vec2 = (D % 2 == 0) && (((uintptr_t)x) % 16 == 0);
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < N_total; j += (long long)gridDim.x * blockDim.x) {
        const double u = stream_uniform(seed, iteration, stream_id, (uint64_t)j, 0);
        if (!(u >= lo && u < hi)) continue;
        const long long a = upper_bound(cdf, n, u);
        const long long dst = j / rows_per_rank, slot = j - dst * rows_per_rank;
        const double* src = x + a * D;
        double* out = peer_out[dst] + slot * D;
        if (vec2) {
            for (int d = 0; d < D; d += 2) *reinterpret_cast<double2*>(out + d) = *reinterpret_cast<const double2*>(src + d);
        } else {
            for (int d = 0; d < D; ++d) out[d] = src[d];
        }
        if (peer_idx) peer_idx[dst][slot] = particle0 + a;
    }
}

// exclusive offset of this rank and the global total from the all-gathered rank totals (sequential fp64 sum in rank
// order: identical on every rank) -- keeps the global scan free of host round trips
__global__ void rank_offsets_kernel(const double* __restrict__ totals, int P, int rank, double* __restrict__ out2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double acc = 0.0, off = 0.0;
        for (int q = 0; q < P; ++q) {
            if (q == rank) off = acc;
            acc += totals[q];
        }
        out2[0] = off;
        out2[1] = acc;
    }
}

}  // namespace smcb

using namespace smcb;

extern "C" {

int smcb_rank_offsets(const double* totals, int P, int rank, double* out2, void* stream) {
    SMCB_REQUIRE(totals && out2 && P >= 1 && rank >= 0 && rank < P, "bad argument");
    rank_offsets_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(totals, P, rank, out2);
    return check_launch("rank_offsets_kernel");
}

int smcb_resample_multinomial_push(const double* cdf, long long n, const double* ends, int rank, int P, uint64_t seed,
                                   uint32_t iteration, uint32_t stream_id, long long N_total, long long particle0,
                                   const double* x, int D, double* const* peer_out, int64_t* const* peer_idx,
                                   long long rows_per_rank, void* stream) {
    SMCB_REQUIRE(cdf && ends && x && peer_out && n >= 1 && P >= 1 && rank >= 0 && rank < P && N_total >= 1 && D >= 1 &&
                     rows_per_rank >= 1, "bad argument");
    resample_multinomial_push_kernel<<<stride_grid(N_total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        cdf, n, ends, rank, P, seed, iteration, stream_id, N_total, particle0, x, D, peer_out, peer_idx, rows_per_rank);
    return check_launch("resample_multinomial_push_kernel");
}

long long smcb_scan_workspace_bytes(long long N) {
    const long long ntiles = (N + kTile - 1) / kTile;
    return (ntiles + 8) * 8;
}

int smcb_cdf(const double* wn, long long N, const double* offset_in, const double* total_in, double* cdf,
             double* total_out, void* workspace, void* stream) {
    SMCB_REQUIRE(wn && cdf && total_out && workspace && N >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long ntiles = (N + kTile - 1) / kTile;
    double* tile_sums = (double*)workspace;
    double* norm_total = tile_sums + ntiles;
    const long long cap = (long long)device_sm_count() * 8;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    tile_sum_kernel<<<grid, kScanThreads, 0, st>>>(wn, N, tile_sums, ntiles);
    if (check_launch("tile_sum_kernel")) return -1;
    tile_scan_wide_kernel<<<1, kWideThreads, 0, st>>>(tile_sums, ntiles, offset_in, total_in, total_out, norm_total);
    if (check_launch("tile_scan_wide_kernel")) return -1;
    tile_cdf_kernel<<<grid, kScanThreads, 0, st>>>(wn, N, tile_sums, norm_total, nullptr, cdf, ntiles);
    return check_launch("tile_cdf_kernel");
}

int smcb_normalise_tilesums(const double* logw, long long N, const double* logZ, double* wn, double* total_out,
                            void* workspace, void* stream) {
    SMCB_REQUIRE(logw && logZ && wn && total_out && workspace && N >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long ntiles = (N + kTile - 1) / kTile;
    double* tile_sums = (double*)workspace;
    double* norm_total = tile_sums + ntiles;
    const long long cap = (long long)device_sm_count() * 8;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    normalise_tile_sum_kernel<<<grid, kScanThreads, 0, st>>>(logw, N, logZ, wn, tile_sums, ntiles);
    if (check_launch("normalise_tile_sum_kernel")) return -1;
    tile_scan_wide_kernel<<<1, kWideThreads, 0, st>>>(tile_sums, ntiles, nullptr, nullptr, total_out, norm_total);
    return check_launch("tile_scan_wide_kernel");
}

int smcb_cdf_from_tilesums(const double* wn, long long N, const double* offset_total, double* cdf, void* workspace,
                           void* stream) {
    SMCB_REQUIRE(wn && cdf && workspace && N >= 1, "bad argument");
    const long long ntiles = (N + kTile - 1) / kTile;
    double* tile_sums = (double*)workspace;
    double* norm_total = tile_sums + ntiles;
    const long long cap = (long long)device_sm_count() * 8;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    tile_cdf_kernel<<<grid, kScanThreads, 0, (cudaStream_t)stream>>>(wn, N, tile_sums, norm_total, offset_total, cdf, ntiles);
    return check_launch("tile_cdf_kernel");
}

int smcb_ancestors_multinomial(const double* cdf, long long N, const double* u, long long M, int64_t* idx,
                               void* stream) {
    SMCB_REQUIRE(cdf && u && idx && N >= 1 && M >= 0, "bad argument");
    if (M == 0) return 0;
    ancestors_multinomial_kernel<<<stride_grid(M, 256, 8), 256, 0, (cudaStream_t)stream>>>(cdf, N, u, M, idx);
    return check_launch("ancestors_multinomial_kernel");
}

int smcb_ancestors_systematic(const double* cdf, long long N, double u0, long long j0, long long M_total,
                              long long M, int64_t* idx, void* stream) {
    SMCB_REQUIRE(cdf && idx && N >= 1 && M >= 0 && M_total >= 1, "bad argument");
    if (M == 0) return 0;
    ancestors_systematic_kernel<<<stride_grid(M, 256, 8), 256, 0, (cudaStream_t)stream>>>(cdf, N, u0, j0, M_total, M, idx);
    return check_launch("ancestors_systematic_kernel");
}

long long smcb_resample_workspace_bytes(long long M, int D) {
    (void)D;
    return ((M + 31) / 32 + 2) * 8;
}

int smcb_resample_systematic(const double* cdf, long long N, double u0, const double* u0_dev, long long j0,
                             long long M_total, long long M, const double* x, int D, double* out, int64_t* idx,
                             void* workspace, void* stream) {
    SMCB_REQUIRE(cdf && x && out && workspace && N >= 1 && M >= 0 && M_total >= 1 && D >= 1, "bad argument");
    SMCB_REQUIRE(x != out, "resampling cannot run in place");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = (((uintptr_t)x | (uintptr_t)out) % 16) == 0;
    if (aligned && (D == 2 || D == 4 || D == 8 || D == 16 || D == 32 || D == 64)) {
        const int lpr = D / 2;
        const int rows = 32;
        const long long ntiles = (M + rows - 1) / rows;
        long long* bounds = (long long*)workspace;
        tile_bounds_kernel<<<stride_grid(ntiles + 1, 256, 8), 256, 0, st>>>(cdf, N, u0, u0_dev, j0, M_total, M, rows, ntiles,
                                                                            bounds);
        if (check_launch("tile_bounds_kernel")) return -1;
        const long long blocks = (ntiles + 7) / 8;
        const long long cap = (long long)device_sm_count() * 8;
        const int grid = (int)(blocks < cap ? blocks : cap);
#define SMCB_RS(L) \
    resample_systematic_fused_kernel<L><<<grid, 256, 0, st>>>(cdf, N, u0, u0_dev, j0, M_total, M, x, out, idx, bounds, nullptr, 1)
        switch (lpr) {
            case 1: SMCB_RS(1); break;
            case 2: SMCB_RS(2); break;
            case 4: SMCB_RS(4); break;
            case 8: SMCB_RS(8); break;
            case 16: SMCB_RS(16); break;
            default: SMCB_RS(32); break;
        }
#undef SMCB_RS
        return check_launch("resample_systematic_fused_kernel");
    }
    // general D: separate ancestor search (needs idx scratch from the caller) and gather
    SMCB_REQUIRE(idx, "this D needs an idx buffer (unfused path)");
    SMCB_REQUIRE(!u0_dev, "the unfused path takes u0 by value");
    ancestors_systematic_kernel<<<stride_grid(M, 256, 8), 256, 0, st>>>(cdf, N, u0, j0, M_total, M, idx);
    if (check_launch("ancestors_systematic_kernel")) return -1;
    gather_rows_kernel<1><<<stride_grid(M * D, 256, 8), 256, 0, st>>>(x, idx, M, D, out);
    return check_launch("gather_rows_kernel");
}

int smcb_resample_systematic_push(const double* cdf, long long N, double u0, long long j0, long long M_total,
                                  long long M, const double* x, int D, double* const* peer_out, long long rows_per_rank,
                                  int64_t* idx, void* workspace, void* stream) {
    SMCB_REQUIRE(cdf && x && peer_out && workspace && N >= 1 && M >= 0 && M_total >= 1 && rows_per_rank >= 1, "bad argument");
    SMCB_REQUIRE(D == 2 || D == 4 || D == 8 || D == 16 || D == 32 || D == 64, "push path needs D in {2,4,8,16,32,64}");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int lpr = D / 2;
    const long long ntiles = (M + 31) / 32;
    long long* bounds = (long long*)workspace;
    tile_bounds_kernel<<<stride_grid(ntiles + 1, 256, 8), 256, 0, st>>>(cdf, N, u0, nullptr, j0, M_total, M, 32, ntiles, bounds);
    if (check_launch("tile_bounds_kernel")) return -1;
    const long long blocks = (ntiles + 7) / 8;
    const long long cap = (long long)device_sm_count() * 8;
    const int grid = (int)(blocks < cap ? blocks : cap);
#define SMCB_RSP(L)                                                                                                      \
    resample_systematic_fused_kernel<L><<<grid, 256, 0, st>>>(cdf, N, u0, nullptr, j0, M_total, M, x, nullptr, idx, bounds, \
                                                              peer_out, rows_per_rank)
    switch (lpr) {
        case 1: SMCB_RSP(1); break;
        case 2: SMCB_RSP(2); break;
        case 4: SMCB_RSP(4); break;
        case 8: SMCB_RSP(8); break;
        case 16: SMCB_RSP(16); break;
        default: SMCB_RSP(32); break;
    }
#undef SMCB_RSP
    return check_launch("resample_systematic_fused_kernel(push)");
}

// ---- peer-visible buffers (CUDA IPC): the destinations of the fused migration
int smcb_peer_alloc(long long bytes, void** ptr, void* handle64) {
    SMCB_REQUIRE(ptr && handle64 && bytes > 0, "bad argument");
    SMCB_CUDA(cudaMalloc(ptr, (size_t)bytes));
    cudaIpcMemHandle_t h;
    SMCB_CUDA(cudaIpcGetMemHandle(&h, *ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return 0;
}
int smcb_peer_open(const void* handle64, void** ptr) {
    SMCB_REQUIRE(ptr && handle64, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SMCB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int smcb_peer_close(void* ptr) {
    if (ptr) SMCB_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}
int smcb_peer_free(void* ptr) {
    if (ptr) SMCB_CUDA(cudaFree(ptr));
    return 0;
}

int smcb_gather_rows(const double* x, const int64_t* idx, long long M, int D, double* out, void* stream) {
    SMCB_REQUIRE(x && idx && out && M >= 0 && D >= 1, "bad argument");
    SMCB_REQUIRE(x != out, "gather cannot run in place");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec2 = (D % 2 == 0) && (((uintptr_t)x | (uintptr_t)out) % 16 == 0);
    if (vec2) gather_rows_kernel<2><<<stride_grid(M * (D / 2), 256, 8), 256, 0, st>>>(x, idx, M, D, out);
    else gather_rows_kernel<1><<<stride_grid(M * D, 256, 8), 256, 0, st>>>(x, idx, M, D, out);
    return check_launch("gather_rows_kernel");
}

}  // extern "C"
