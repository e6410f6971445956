// K2/K3 -- one lane of the batched iterative NUTS transition.
//
// Re-derivation (not a translation) of the recursive sampler in
//   /root/reference/smcnuts/proposal/nuts.py:58-112 (generate_nuts_samples), :114-150 (build_tree),
//   :152-160 (stop_criterion), :162-175 (NUTSLeapfrog) and the endpoint MH step of
//   /root/reference/smcnuts/proposal/nuts_acc_rej.py:42-49 + utils.py:22-34
// as an explicit state machine that advances ONE leapfrog per step, so that 32 lanes holding 32
// different particles at different tree positions stay converged on the expensive model evaluation
// and a finished lane can be refilled from the particle work queue.
//
// Tree bookkeeping without recursion (per lane):
//   * registers hold the ACTIVE edge (x, r, g) = the end of the trajectory that the current doubling
//     extends (side `dir`); the other edge, the U-turn checkpoints and the candidate store live in the
//     lane's workspace record in global memory (L1/L2 resident; touched O(1) times per leapfrog).
//   * U-turn checkpoints: the first state of every pending sub-tree is written once, at the even
//     0-based leaf f that starts it, into slot popc(f); the sub-tree of 2^(l+1) leaves that ends at
//     0-based leaf i reads slot popc(i - 2^(l+1) + 1).  Slots of live sub-trees never collide.
//   * candidates: pending first children keep (count, slot-reference) per level, packed into two
//     64-bit registers; a candidate is materialised in the store only when it survives its merges.
//   * merges happen in post-order after leaf i for levels 0..ctz(i)-1, consuming one uniform each,
//     exactly the draw order of the recursion (nuts.py:142); any stop (divergence nuts.py:125 or
//     sub-tree U-turn nuts.py:148) ends the transition at once, because the reference then never
//     reads the sub-tree's candidate nor draws again from this particle's stream.
//
// Wide records (more than 4 coordinates per lane: the Gaussian tensor-core kernels, 26 per lane at D = 100) cannot
// afford registers for the stored edge of a U-turn test -- with 255 registers in use the compiler chops its 52 loads
// into short dependent batches, and the kernel spent 45 % of its warp time on their L2 round trips (ncu, round 2).
// For those (M::STAGE) every lane owns a staging row in shared memory: the first leaf of a two-leaf sub-tree is copied
// there when it is stored (the level-0 test at the next leaf, half of all tests, then never touches the workspace), and
// deeper checkpoints / the other edge arrive by cp.async -- all 26 16-byte copies in flight at once, no registers --
// issued before the Philox draw and the merge bookkeeping that precede the test.
//
// The selected sample: when a doubling is accepted and its candidate sits in a leaf slot, the lane only remembers the
// slot (it stays allocated across doublings) and copies it into the caller's x_new/r_new row once, at the end of the
// transition; a candidate that is the leaf in registers, and the start point, are written to the row at once.
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "philox.cuh"

namespace smcb {

struct NutsArgs {
    ModelDesc model;
    const double* x;      // [N, D] current positions
    const double* r;      // [N, D] current momenta
    long long N;
    double eps, phi;
    int max_depth;        // reference: MAX_TREE_DEPTH = 10 (nuts.py:4) -> at most max_depth+1 doublings
    int accrej;           // fuse the endpoint MH step (NUTSProposalWithAccRej)
    uint64_t seed;
    uint32_t iteration;
    uint64_t particle0;   // global index of local particle 0 (multi-GPU sharding)
    double* x_new;        // [N, D]
    double* r_new;        // [N, D]
    double* A_old;        // [N] log prior + Jacobian at x           (nullable)
    double* B_old;        // [N] log likelihood at x                 (nullable)
    double* A_new;        // [N] same at the returned x_new          (nullable)
    double* B_new;        // [N]                                     (nullable)
    double* ke_old;       // [N] 0.5*|r|^2                           (nullable)
    double* ke_new;       // [N] 0.5*|r_new|^2                       (nullable)
    int* n_leapfrog;      // [N] leapfrog steps = gradient evaluations excluding the initial one (nullable)
    int* accepted;        // [N] MH outcome (1 when accrej == 0)     (nullable)
    int* depth;           // [N] number of doublings                 (nullable)
    double* accept_stat;  // [N] mean over the leaves of min(1, exp(joint_leaf - joint_0)): the NUTS acceptance statistic
                          //     that drives dual-averaging step-size adaptation (nullable: not computed)
    // gradient carry-over (optional, accrej == 0 only): when the caller hands back the split log density and the
    // gradient of the current positions (outputs A_new, B_new, g_new of the previous transition at the same phi), the
    // initial evaluation of every transition (nuts.py:66,72) is skipped -- same numbers, one model evaluation less.
    const double* A_in;   // [N]      (nullable; all three or none)
    const double* B_in;   // [N]
    const double* g_in;   // [N, D]   gradient of A + phi*B at x
    double* g_new;        // [N, D]   gradient at the returned x_new (nullable)
    double* ws;           // workspace: lanes * ws_doubles(D, max_depth)
    unsigned long long* queue;  // work-queue head, zeroed before launch
    void* exchange;             // tail compaction: one Lane<M> per thread of the grid (nuts_launch.cuh)
    int blocks_per_sm;          // cap on resident CTAs per SM (0: as many as fit)
};

// 16-byte pair, the unit of every access to the per-lane workspace record (LDG.128 / STG.128)
struct alignas(16) D2 {
    double x, y;
};

// Workspace record of one LANE (global memory, 128-byte aligned; nlp = nl rounded up to even so that every vector is
// 16-byte aligned):
//   header[8]: A0 B0 ke0 | As Bs kes | sum of leaf acceptance probabilities, joint_0 (only with accept_stat) -- the split log densities and kinetic energies of the start point and of
//              the selected sample: written once or twice per transition, read at its end, so they live here and not in
//              registers that are precious across the model evaluation
//   other edge: x[nlp] r[nlp] g[nlp] | 2L+3 leaf slots, each x[nlp] r[nlp] A B [g[nlp]]
// A leaf state is stored at most ONCE: the first leaf of a sub-tree (U-turn checkpoint) and a pending candidate are
// the same record when they are the same leaf (every odd leaf of a doubling is both); checkpoints and candidates
// are slot references handed out from one free mask.  Candidates carry their gradient only when the caller asked
// for g_new.
SMCB_HD int nuts_nlp(int nl) { return (nl + 1) & ~1; }
// doubles per lane of the shared-memory staging row (x[nlp] r[nlp] + 2 of padding: lanes then start 4 banks apart in
// every quarter-warp phase of an LDS.128, i.e. conflict-free)
SMCB_HD int nuts_stage_stride(int nl) { return 2 * nuts_nlp(nl) + 2; }
SMCB_HD int nuts_slot_stride(int nl, bool carry) { return 2 * nuts_nlp(nl) + 2 + (carry ? nuts_nlp(nl) : 0); }
constexpr int kNutsHdr = 8;
SMCB_HD int nuts_ws_doubles(int nl, int L, bool carry = true) {
    return ((kNutsHdr + 3 * nuts_nlp(nl) + nuts_slot_stride(nl, carry) * (2 * L + 3)) + 15) & ~15;
}

enum LanePhase : int { kIdle = 0, kInit = 1, kLeaf = 2 };

// A particle is owned by M::GROUP adjacent lanes (1 for arma / PRMwCD; 4 for the tensor-core Gaussian, where lane
// `sub` of the group holds coordinates sub, sub+4, sub+8, ...).  All lanes of a group execute the same control flow
// (same particle, same Philox stream); only dot products need the group fold `gsum`.
template <class M>
struct Lane {
    static constexpr int G = M::GROUP;
    static constexpr int DM = M::NLOC;
    double xa[DM], ra[DM], ga[DM];  // active edge (this lane's coordinates)
    double logu;
    long long pid;
    double* ws;       // per-lane record in global memory
    double* stg;      // per-lane staging row in shared memory (M::STAGE only): x[nlp] r[nlp] of a stored edge
    int nlp_, slot_stride;
    int phase, dir, depth, D, L, nl, sub;
    uint32_t leaf, n_tot, n_leapfrog;
    uint32_t ck_used, cand_used, ck_valid;   // slots referenced by live checkpoints / candidates; valid checkpoint ids
    uint32_t samp_used;                      // slot holding the currently selected sample (one bit, or 0: it is in the row)
    int samp_ref;
    uint64_t pend_n, pend_ref, ck_ref;       // packed per-level counts, candidate slot refs, checkpoint slot refs (5 bits)
    StreamReader rng;

#define SMCB_LOCAL(i) for (int i = 0; i < (M::STATIC_NL ? M::STATIC_NL : nl); ++i)
#define SMCB_PAIRS(i) for (int i = 0; i < (M::STATIC_NL ? M::STATIC_NL : nl); i += 2)

    SMCB_HD int gd(int i) const { return G == 1 ? i : sub + G * i; }   // global coordinate of local slot i
    // ---- draws of this particle's NUTS stream, in the reference's order (nuts.py:69,91,99,142)
    SMCB_HD uint64_t draw_bits(const NutsArgs& a) {
        return rng.next_bits(a.seed, (a.iteration << 8) | (uint32_t)kStreamNuts, a.particle0 + (uint64_t)pid);
    }
    SMCB_HD double draw(const NutsArgs& a) { return (double)draw_bits(a) * 0x1.0p-53; }
    // u < num/den for integers 0 <= num, 1 <= den < 2^11, decided in exact integer arithmetic: k * den < num * 2^53.
    // The reference compares u with the ROUNDED quotient (nuts.py:142, :99); the two can only differ when u is the
    // 2^-53-grid neighbour of num/den, i.e. with probability ~2^-53 per draw -- and no FP64 division is needed.
    SMCB_HD bool draw_below_ratio(const NutsArgs& a, uint32_t num, uint32_t den) {
        const uint64_t k = draw_bits(a);
        return k * (uint64_t)den < ((uint64_t)num << 53);
    }
    // padded coordinates per lane: a compile-time constant for the static models (no register)
    SMCB_HD int nlp_get() const { return M::STATIC_NL ? ((M::STATIC_NL + 1) & ~1) : nlp_; }
#define nlp nlp_get()
    SMCB_HD double gsum(double v) const {
#if defined(SMCB_WARP_CODE)
        if (G > 1) {
            const unsigned mask = ((1u << G) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
#pragma unroll
            for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(mask, v, o);
        }
#endif
        return v;
    }

    // ---- vector <-> record, two doubles (16 bytes) per access
    // The wide records of the staged group kernels (8-26 coordinates per lane: the Gaussian tensor-core kernels) go to
    // L2 only (ld/st.global.cg): 256 lanes x ~10 KB of records cannot live in an SM's L1 anyway, and streaming them
    // through it evicted the kernel's register spills, whose reloads then showed up as long-scoreboard stalls on the
    // bookkeeping scalars (ncu, round 2).  MEASURED (B200, D = 100): 78.6 -> 71.4 ms at N = 2^20, 20.5 -> 18.7 ms at 2^18.
    // Not for the one-lane models: their ~1 KB records DO live in L1 (PRMwCD one-lane kernel, 6000 particles: 6x slower).
    // (With these accesses AND tail compaction the one-lane PRMwCD kernel did not terminate: a lane state that moves to
    // another warp is followed by L2-only loads of its record, and without a __threadfence() in front of the exchange
    // barrier those could overtake the previous owner's L2-only stores.  The fence is there now -- nuts_launch.cuh --,
    // found by bisecting debug builds on the GPU, gpurun_out/r2ak; with ordinary L1-cached accesses the same-SM L1 keeps
    // the two coherent.  The staged models have no tail mode anyway.)
#ifndef SMCB_WIDE_RECORDS_CG
#define SMCB_WIDE_RECORDS_CG 1
#endif
    static constexpr bool kCg = (SMCB_WIDE_RECORDS_CG != 0) && M::STAGE;
    SMCB_HD static void st2(double* p, const D2& t) {
#if defined(__CUDA_ARCH__)
        if constexpr (kCg) { __stcg(reinterpret_cast<double2*>(p), make_double2(t.x, t.y)); return; }
#endif
        *reinterpret_cast<D2*>(p) = t;
    }
    SMCB_HD static D2 ld2(const double* p) {
#if defined(__CUDA_ARCH__)
        if constexpr (kCg) { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); D2 t; t.x = v.x; t.y = v.y; return t; }
#endif
        return *reinterpret_cast<const D2*>(p);
    }
    SMCB_HD void stv(double* p, const double (&v)[DM]) const {
        const int n_ = (M::STATIC_NL ? M::STATIC_NL : nl);
#pragma unroll
        SMCB_PAIRS(i) {
            D2 t;
            t.x = v[i];
            t.y = (i + 1 < n_) ? v[i + 1 < DM ? i + 1 : i] : 0.0;
            st2(p + i, t);
        }
    }
    SMCB_HD void ldv(const double* p, double (&v)[DM]) const {
        const int n_ = (M::STATIC_NL ? M::STATIC_NL : nl);
#pragma unroll
        SMCB_PAIRS(i) {
            const D2 t = ld2(p + i);
            v[i] = t.x;
            if (i + 1 < n_) v[i + 1 < DM ? i + 1 : i] = t.y;
        }
    }

    // ---- record views
    SMCB_HD double* other_x() const { return ws + kNutsHdr; }
    SMCB_HD double* other_r() const { return ws + kNutsHdr + nlp; }
    SMCB_HD double* other_g() const { return ws + kNutsHdr + 2 * nlp; }
    SMCB_HD double* slotp(int slot) const { return ws + kNutsHdr + 3 * nlp + slot_stride * slot; }   // x[nlp] r[nlp] A B [g[nlp]]
    SMCB_HD int alloc_slot() {
        const uint32_t free_ = ~(ck_used | cand_used | samp_used);
        return ctz32(free_);   // 2L+3 <= 23 slots: at most L checkpoints + L+1 candidates + the selected sample are live
    }
    // store the leaf in registers (active edge) into a fresh slot
    SMCB_HD int store_leaf(const NutsArgs& a, double A, double B) {
        const int sl = alloc_slot();
        double* c = slotp(sl);
        stv(c, xa); stv(c + nlp, ra);
        D2 ab; ab.x = A; ab.y = B;
        *reinterpret_cast<D2*>(c + 2 * nlp) = ab;
        if (a.g_new) stv(c + 2 * nlp + 2, ga);
        return sl;
    }

    // ---- packed per-level pending counts: level l occupies bits [l(l+1)/2, +l+1)
    SMCB_HD uint32_t get_n(int l) const { return (uint32_t)(pend_n >> (l * (l + 1) / 2)) & ((2u << l) - 1u); }
    SMCB_HD void set_n(int l, uint32_t v) {
        const int sh = l * (l + 1) / 2;
        const uint64_t mask = (uint64_t)((2u << l) - 1u) << sh;
        pend_n = (pend_n & ~mask) | ((uint64_t)v << sh);
    }
    SMCB_HD int get_ref(int l) const { return (int)((pend_ref >> (5 * l)) & 31u); }
    SMCB_HD void set_ref(int l, int s) { pend_ref = (pend_ref & ~((uint64_t)31 << (5 * l))) | ((uint64_t)s << (5 * l)); }
    SMCB_HD int get_ck(int p) const { return (int)((ck_ref >> (5 * p)) & 31u); }
    SMCB_HD void set_ck(int p, int s) {
        if (ck_valid & (1u << p)) ck_used &= ~(1u << get_ck(p));   // the checkpoint it replaces is dead by construction
        ck_ref = (ck_ref & ~((uint64_t)31 << (5 * p))) | ((uint64_t)s << (5 * p));
        ck_valid |= 1u << p;
        ck_used |= 1u << s;
    }

    SMCB_HD void idle_init(const M& m, int sub_) {
        phase = kIdle; sub = sub_; D = m.dim(); nl = m.nloc(); nlp_ = nuts_nlp(nl); pid = -1;
        SMCB_LOCAL(i) { xa[i] = 0.0; ra[i] = 0.0; ga[i] = 0.0; }
    }

    SMCB_HD void begin(const NutsArgs& a, const M& m, long long p, double* ws_) {
        pid = p; ws = ws_; D = m.dim(); nl = m.nloc(); nlp_ = nuts_nlp(nl); L = a.max_depth;
        const int d_ = D;
#pragma unroll
        SMCB_LOCAL(i) {
            const bool ok = gd(i) < d_;
            xa[i] = ok ? a.x[p * d_ + gd(i)] : 0.0;
            ra[i] = ok ? a.r[p * d_ + gd(i)] : 0.0;
        }
        rng.reset();
        n_leapfrog = 0;
        slot_stride = nuts_slot_stride(nl, a.g_new != nullptr);
        phase = kInit;
        if (a.g_in) {   // carried-over evaluation: initialise the tree right away, the first trip is already a leapfrog
#pragma unroll
            SMCB_LOCAL(i) ga[i] = gd(i) < d_ ? a.g_in[p * d_ + gd(i)] : 0.0;
            init_tree(a, a.A_in[p], a.B_in[p]);
        }
    }

    // first half of the leapfrog (nuts.py:169-170); nothing to do before the initial evaluation
    SMCB_HD void pre_eval(const NutsArgs& a) {
        if (phase != kLeaf) return;
        const double half = dir * a.eps / 2, full = dir * a.eps;
#pragma unroll
        SMCB_LOCAL(i) {
            ra[i] = ra[i] + half * ga[i];
            xa[i] = xa[i] + full * ra[i];
        }
    }

    SMCB_HD void start_doubling(const NutsArgs& a, bool first) {
        const int nd = (draw_bits(a) < (1ull << 52)) ? 1 : -1;  // u < 0.5, nuts.py:91
        if (!first && nd != dir) {                   // bring the other edge into registers
            double tx[DM], tr[DM], tg[DM];
            ldv(other_x(), tx); ldv(other_r(), tr); ldv(other_g(), tg);
            stv(other_x(), xa); stv(other_r(), ra); stv(other_g(), ga);
#pragma unroll
            SMCB_LOCAL(i) { xa[i] = tx[i]; ra[i] = tr[i]; ga[i] = tg[i]; }
        }
        dir = nd;
        leaf = 0; pend_n = 0; pend_ref = 0; ck_ref = 0;
        ck_used = cand_used = ck_valid = 0;
    }

    // start_doubling(false) when x and r of the other edge are already in registers (only its gradient is loaded)
    SMCB_HD void start_doubling_with(const NutsArgs& a, const double (&xo)[DM], const double (&ro)[DM]) {
        const int nd = (draw_bits(a) < (1ull << 52)) ? 1 : -1;  // u < 0.5, nuts.py:91
        if (nd != dir) {
            double tg[DM];
            ldv(other_g(), tg);
            stv(other_x(), xa); stv(other_r(), ra); stv(other_g(), ga);
#pragma unroll
            SMCB_LOCAL(i) { xa[i] = xo[i]; ra[i] = ro[i]; ga[i] = tg[i]; }
        }
        dir = nd;
        leaf = 0; pend_n = 0; pend_ref = 0; ck_ref = 0;
        ck_used = cand_used = ck_valid = 0;
    }

    // ---- shared-memory staging of a stored edge (M::STAGE)
    static constexpr bool kStage = M::STAGE;
    SMCB_HD void stage_from_regs() {   // the active edge (the leaf just stored) -> staging row (shared memory: plain stores)
#pragma unroll
        SMCB_PAIRS(i) {
            const int j = i + 1 < DM ? i + 1 : i;
            D2 tx, tr;
            tx.x = xa[i]; tx.y = (i + 1 < DM) ? xa[j] : 0.0;
            tr.x = ra[i]; tr.y = (i + 1 < DM) ? ra[j] : 0.0;
            *reinterpret_cast<D2*>(stg + i) = tx;
            *reinterpret_cast<D2*>(stg + nlp + i) = tr;
        }
    }
    // start copying x[nlp] r[nlp] at c (workspace, contiguous) into the staging row; nothing waits here
    SMCB_HD void stage_issue(const double* c) {
#if defined(__CUDA_ARCH__)
        const unsigned sa = (unsigned)__cvta_generic_to_shared(stg);
#pragma unroll
        for (int i = 0; i < 2 * DM; i += 2)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + 8u * i), "l"(c + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
#else
        for (int i = 0; i < 2 * nlp; ++i) stg[i] = c[i];
#endif
    }
    SMCB_HD void stage_wait() const {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    }
    // U-turn test of the staged edge against the active edge, operands read pair by pair from shared memory
    SMCB_HD bool uturn_stage() const {
        // two partial sums per dot product: the 26-term chains of dependent DFMAs are latency, not throughput
        double s1 = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
        SMCB_PAIRS(i) {
            const D2 xc = *reinterpret_cast<const D2*>(stg + i);
            const D2 rc = *reinterpret_cast<const D2*>(stg + nlp + i);
            const double dx0 = xa[i] - xc.x;
            s1 += dx0 * rc.x;
            s2 += dx0 * ra[i];
            if (i + 1 < DM) {
                const double dx1 = xa[i + 1 < DM ? i + 1 : i] - xc.y;
                t1 += dx1 * rc.y;
                t2 += dx1 * ra[i + 1 < DM ? i + 1 : i];
            }
        }
        s1 = gsum(s1 + t1); s2 = gsum(s2 + t2);
        return (dir * s1 < 0) || (dir * s2 < 0);
    }
    // sum of squares of the active momentum: split into partial sums for wide records (a group kernel's summation order
    // differs from the oracle's sequential one anyway; the one-lane kernels keep the oracle's order for bit-parity)
    SMCB_HD double ra_sqnorm() const {
        if constexpr (G > 1 && DM > 4) {
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
            for (int i = 0; i < DM; i += 4) {
                p0 += ra[i] * ra[i];
                if (i + 1 < DM) p1 += ra[i + 1 < DM ? i + 1 : i] * ra[i + 1 < DM ? i + 1 : i];
                if (i + 2 < DM) p2 += ra[i + 2 < DM ? i + 2 : i] * ra[i + 2 < DM ? i + 2 : i];
                if (i + 3 < DM) p3 += ra[i + 3 < DM ? i + 3 : i] * ra[i + 3 < DM ? i + 3 : i];
            }
            return (p0 + p1) + (p2 + p3);
        } else {
            double rr = 0.0;
#pragma unroll
            SMCB_LOCAL(i) rr += ra[i] * ra[i];
            return rr;
        }
    }
    // start_doubling(false) when x and r of the other edge sit in the staging row: exchange the edges pair by pair
    SMCB_HD void start_doubling_from_stage(const NutsArgs& a) {
        const int nd = (draw_bits(a) < (1ull << 52)) ? 1 : -1;  // u < 0.5, nuts.py:91
        if (nd != dir) {
            double* og = other_g();
#pragma unroll
            SMCB_PAIRS(i) {
                const D2 tg = ld2(og + i);
                const D2 tx = *reinterpret_cast<const D2*>(stg + i);
                const D2 tr = *reinterpret_cast<const D2*>(stg + nlp + i);
                D2 mx, mr, mg;
                mx.x = xa[i]; mr.x = ra[i]; mg.x = ga[i];
                const int j = i + 1 < DM ? i + 1 : i;
                mx.y = xa[j]; mr.y = ra[j]; mg.y = ga[j];
                st2(other_x() + i, mx);
                st2(other_r() + i, mr);
                st2(og + i, mg);
                xa[i] = tx.x; ra[i] = tr.x; ga[i] = tg.x;
                if (i + 1 < DM) { xa[j] = tx.y; ra[j] = tr.y; ga[j] = tg.y; }
            }
        }
        dir = nd;
        leaf = 0; pend_n = 0; pend_ref = 0; ck_ref = 0;
        ck_used = cand_used = ck_valid = 0;
    }

    // U-turn test between a stored edge (x at c, r at c + nlp) and the active edge (nuts.py:152-160); the edge order
    // (minus, plus) is restored through `dir`.
    SMCB_HD bool uturn(const double* c) const {
        double xc[DM], rc[DM];
        ldv(c, xc); ldv(c + nlp, rc);
        return uturn_regs(xc, rc);
    }
    SMCB_HD bool uturn_regs(const double (&xc)[DM], const double (&rc)[DM]) const {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        SMCB_LOCAL(i) {
            const double dx = xa[i] - xc[i];
            s1 += dx * rc[i];
            s2 += dx * ra[i];
        }
        s1 = gsum(s1); s2 = gsum(s2);
        return (dir * s1 < 0) || (dir * s2 < 0);
    }
    // Small records (<= 4 coordinates per lane: arma, the PRMwCD group kernel): the stored edge of a U-turn test is
    // loaded BEFORE the Philox draw and the merge bookkeeping that precede the test, so the load latency (L2 more
    // often than L1: the records of a CTA exceed its L1 share) is covered by ~80 independent instructions.  Larger
    // records cannot afford the 4*DM extra live registers.
#ifndef SMCB_NUTS_EARLY_LOADS
#define SMCB_NUTS_EARLY_LOADS 1
#endif
    static constexpr bool kEarly = (SMCB_NUTS_EARLY_LOADS != 0) && DM <= 4;
    // ... and the level-0 checkpoint (the operand of half of all U-turn tests: the first leaf of the two-leaf sub-tree the
    // next leaf completes) is known BEFORE the model evaluation: its 2*DM doubles are loaded then and ride through the
    // evaluation in registers, so that test never waits on memory.  MEASURED (B200, same call): arma N = 2^16 0.602 ->
    // 0.553 ms, 2^17 0.717 -> 0.673 ms, 2^20 2.910 -> 2.864 ms (the gain is latency: it grows as the shard shrinks); the
    // 4-lane PRMwCD kernel loses 2 % (124 registers of its 128), so one-lane models only.
#ifndef SMCB_NUTS_PREFETCH_CK
#define SMCB_NUTS_PREFETCH_CK 1
#endif
    static constexpr bool kPrefetchCk = (SMCB_NUTS_PREFETCH_CK != 0) && kEarly && G == 1;
    double pxc[kPrefetchCk ? DM : 1], prc[kPrefetchCk ? DM : 1];
    SMCB_HD void prefetch_ck() {
        if constexpr (kPrefetchCk) {
            if (phase == kLeaf && depth > 0 && (leaf & 1u)) {   // the coming leaf has an odd 0-based index: a level-0 merge follows it
                const double* ck = slotp(get_ck(popc32(leaf - 1u)));
                ldv(ck, pxc); ldv(ck + nlp, prc);
            }
        }
    }

    // rows of the caller's [N, D] arrays: 16-byte accesses when the row layout allows it
    SMCB_HD void write_row(double* base, const double (&v)[DM]) const {
        const int d_ = D;
        if (G == 1 && (d_ & 1) == 0) {
            stv(base + pid * d_, v);   // rows of an even number of doubles are 16-byte aligned (torch allocations are)
        } else {
            double* row = base + pid * d_ + (G == 1 ? 0 : sub);
            const int n_ok = G == 1 ? d_ : (d_ - sub + G - 1) / G;   // local slots that hold a real coordinate
#pragma unroll
            SMCB_LOCAL(i) if (i < n_ok) row[G * i] = v[i];
        }
    }

    SMCB_HD void write_sample_from_active(const NutsArgs& a, double A, double B, double ke) {
        write_row(a.x_new, xa);
        write_row(a.r_new, ra);
        if (a.g_new) write_row(a.g_new, ga);
        ws[3] = A; ws[4] = B; ws[5] = ke;   // every lane of a group keeps the (identical) scalars in its own record
        samp_ref = -1; samp_used = 0u;
    }

    // The deferred copy of the selected sample into the caller's row; returns its (A, B, kinetic energy).
    //   samp_ref >= 0: it sits in a leaf slot;  -1: it is already in the row (values in the header);
    //   -2: it is still the start point -- the row is copied from the inputs here, once, instead of being written at the
    //       start of every transition and usually overwritten
    SMCB_HD void flush_sample(const NutsArgs& a, double& As, double& Bs, double& kes) {
        const int d_ = D;
        if (samp_ref == -2) {
            const double* xin = a.x + pid * d_ + (G == 1 ? 0 : sub);
            const double* rin = a.r + pid * d_ + (G == 1 ? 0 : sub);
            double* xo = a.x_new + pid * d_ + (G == 1 ? 0 : sub);
            double* ro = a.r_new + pid * d_ + (G == 1 ? 0 : sub);
            const int n_ok = G == 1 ? d_ : (d_ - sub + G - 1) / G;
#pragma unroll
            SMCB_LOCAL(i) if (i < n_ok) { xo[G * i] = xin[G * i]; ro[G * i] = rin[G * i]; }
            As = ws[0]; Bs = ws[1]; kes = ws[2];
        } else if (samp_ref < 0) {
            As = ws[3]; Bs = ws[4]; kes = ws[5];
        } else {
            const double* c = slotp(samp_ref);
            double t[DM];
            ldv(c, t); write_row(a.x_new, t);
            ldv(c + nlp, t); write_row(a.r_new, t);
            double k2 = 0.0;
#pragma unroll
            SMCB_LOCAL(i) k2 += t[i] * t[i];
            kes = 0.5 * gsum(k2);
            const D2 ab = *reinterpret_cast<const D2*>(c + 2 * nlp);
            As = ab.x; Bs = ab.y;
            if (a.g_new) { ldv(c + 2 * nlp + 2, t); write_row(a.g_new, t); }
        }
        samp_ref = -1; samp_used = 0u;
    }

    // nuts.py:66-87 given logp = A + phi*B and its gradient (already in `ga`) at the start point
    SMCB_HD void init_tree(const NutsArgs& a, double A, double B) {
        double lp = A + a.phi * B;
        if (!is_finite(lp)) lp = neg_inf();
        const double ke0 = 0.5 * gsum(ra_sqnorm());
        ws[0] = A; ws[1] = B; ws[2] = ke0;
        const double H0 = lp - ke0;
        if (a.accept_stat) { ws[6] = 0.0; ws[7] = H0; }
#if SMCB_TABLE_MATH
        logu = H0 + fast_log(1.0 - draw(a));      // 1 - u is exact; table-driven log (common.cuh), <= 2.2e-16 absolute
#else
        logu = H0 - (-log1p(-draw(a)));
#endif
        if (a.g_new) {   // gradient carry-over wants the start point's gradient in the row: write it now
            write_sample_from_active(a, A, B, ke0);
        } else {
            samp_ref = -2; samp_used = 0u;
        }
        stv(other_x(), xa); stv(other_r(), ra); stv(other_g(), ga);
        n_tot = 1; depth = 0;
        start_doubling(a, true);
        phase = kLeaf;
    }

    // Unconditional hand-over of the fresh gradient (also on idle lanes, so that `ga` is dead across the model
    // evaluation and its registers can hold the accumulators).
    SMCB_HD void take_grad(const double (&gn)[DM]) {
#pragma unroll
        SMCB_LOCAL(i) ga[i] = gn[i];
    }

    // Consume the model evaluation at xa (gradient already in `ga`).  Returns true when the transition is complete.
    SMCB_HD bool post_eval(const NutsArgs& a, double A, double B) {
        const int d_ = D;
        double lp = A + a.phi * B;
        const bool bad = !is_finite(lp);  // bridgestan.py:47-49,79-80: failure -> logp = -inf, grad = -inf
        if (bad) {
            lp = neg_inf();
#pragma unroll
            SMCB_LOCAL(i) if (gd(i) < d_) ga[i] = neg_inf();
        }
        if (phase == kInit) {
            init_tree(a, A, B);
            return false;
        }

        // ---- leaf: second half-kick, slice and divergence tests (nuts.py:121-125,173)
        const double half = dir * a.eps / 2;
#pragma unroll
        SMCB_LOCAL(i) ra[i] = ra[i] + half * ga[i];
        const double rr = gsum(ra_sqnorm());
        ++n_leapfrog;
        ++leaf;
        const double joint = lp - 0.5 * rr;
        if (a.accept_stat) {   // min(1, exp(joint - joint_0)); a NaN or -inf joint counts 0
            const double dj = joint - ws[7];
            ws[6] += (dj >= 0.0) ? 1.0 : ((dj == dj) ? fast_exp(dj) : 0.0);
        }
        // (MEASURED, round 2: funnelling the stop sites below into ONE finish() call shrinks the D = 100 kernel from 10.5 k to
        //  6.0 k instructions -- its four inlined copies of the sample flush / MH epilogue are 40 % of the SASS and
        //  instruction-fetch stalls 14 % of the samples -- but costs registers: 74.4 vs 71.4 ms; arma / PRMwCD neutral.)
        if ((logu - 100.) >= joint) { ++depth; return finish(a); }
        uint32_t run_n = (logu < joint) ? 1u : 0u;
        int run_ref = -1;  // -1: the candidate is the leaf in registers
        const uint32_t nleaves = 1u << depth;
        if (nleaves > 1u) {
            const uint32_t i0 = leaf - 1u;
            if ((i0 & 1u) == 0u) {
                // first leaf of (at least) a two-leaf sub-tree: one record serves as its U-turn checkpoint and as
                // the pending level-0 candidate
                run_ref = store_leaf(a, A, B);
                set_ck(popc32(i0), run_ref);
                if constexpr (kStage) stage_from_regs();   // it is the level-0 checkpoint of the next leaf
            } else {
                const int tz = ctz32(leaf);
                for (int l = 0; l < tz; ++l) {  // nuts.py:136-148, second child = running node
                    const double* ck = slotp(get_ck(popc32(i0 - (2u << l) + 1u)));
                    double xc[kEarly ? DM : 1], rc[kEarly ? DM : 1];
                    if constexpr (kPrefetchCk) {
                        if (l == 0) {   // loaded before the evaluation (prefetch_ck)
#pragma unroll
                            SMCB_LOCAL(i) { xc[i] = pxc[i]; rc[i] = prc[i]; }
                        } else {
                            ldv(ck, xc); ldv(ck + nlp, rc);
                        }
                    } else if constexpr (kEarly) {
                        ldv(ck, xc); ldv(ck + nlp, rc);
                    }
                    if constexpr (kStage) { if (l > 0) stage_issue(ck); }   // l == 0: staged when the previous leaf was stored
                    const uint32_t n1 = get_n(l);
                    const int ref1 = get_ref(l);
                    const uint32_t tot = n1 + run_n;
                    if (draw_below_ratio(a, run_n, tot > 1u ? tot : 1u)) {   // u < n''/max(n'+n'', 1), nuts.py:142
                        cand_used &= ~(1u << ref1);
                    } else {
                        if (run_ref >= 0) cand_used &= ~(1u << run_ref);
                        run_ref = ref1;
                    }
                    run_n = tot;
                    bool stop;
                    if constexpr (kEarly) stop = uturn_regs(xc, rc);
                    else if constexpr (kStage) { stage_wait(); stop = uturn_stage(); }
                    else stop = uturn(ck);
                    if (stop) { ++depth; return finish(a); }
                }
            }
        }
        if (leaf == nleaves) {  // doubling complete and not stopped: nuts.py:99-110
            double xo[kEarly ? DM : 1], ro[kEarly ? DM : 1];   // the other edge: needed by the trajectory U-turn test
            if constexpr (kEarly) { ldv(other_x(), xo); ldv(other_r(), ro); }   // and, on a direction flip, as the new active edge
            if constexpr (kStage) stage_issue(other_x());
            // u < min(1, n'/n), nuts.py:99; u < 1 always, so only n' < n needs the comparison (the draw is consumed anyway)
            const bool take = draw_below_ratio(a, run_n, n_tot) || run_n >= n_tot;
            if (take) {
                if (run_ref < 0) {
                    write_sample_from_active(a, A, B, 0.5 * rr);
                } else {   // keep the slot, copy it out once at the end of the transition (flush_sample)
                    samp_ref = run_ref; samp_used = 1u << run_ref;
                }
            }
            n_tot += run_n;
            bool stop;
            if constexpr (kEarly) stop = uturn_regs(xo, ro);
            else if constexpr (kStage) { stage_wait(); stop = uturn_stage(); }
            else stop = uturn(other_x());
            ++depth;
            if (stop || depth > L) return finish(a);
            if constexpr (kEarly) start_doubling_with(a, xo, ro);
            else if constexpr (kStage) start_doubling_from_stage(a);
            else start_doubling(a, false);
            return false;
        }
        // park the running node as the pending first child of level ctz(leaf)
        const int lv = ctz32(leaf);
        if (run_ref < 0) run_ref = store_leaf(a, A, B);
        cand_used |= 1u << run_ref;
        set_n(lv, run_n);
        set_ref(lv, run_ref);
        return false;
    }

    // End of transition: optional endpoint MH step (nuts_acc_rej.py:42-49, utils.py:22-34) and outputs.
    SMCB_HD bool finish(const NutsArgs& a) {
        const int d_ = D;
        double As, Bs, kes;
        flush_sample(a, As, Bs, kes);
        const double A0 = ws[0], B0 = ws[1], ke0 = ws[2];
        double ken = kes;
        int anyinf = 0;
        if (a.accrej) {   // np.any(np.isinf(x_prime)), utils.py:32
            SMCB_LOCAL(i) {
                if (gd(i) < d_) {
                    const double xv = a.x_new[pid * d_ + gd(i)];
                    anyinf |= (xv == -neg_inf()) || (xv == neg_inf());
                }
            }
            if (G > 1) anyinf = gsum((double)anyinf) > 0.0;
        }
        int acc = 1;
        if (a.accrej) {
            double lps = As + a.phi * Bs, lp0 = A0 + a.phi * B0;
            if (!is_finite(lps)) lps = neg_inf();
            if (!is_finite(lp0)) lp0 = neg_inf();
            const double H1 = lps - ken, H0 = lp0 - ke0;
            const double ratio = fast_exp(H1 - H0);
            const double prob = (ratio < 1.) ? ratio : 1.;  // python min(1., ratio): nan -> 1.
            const double u = stream_uniform(a.seed, a.iteration, kStreamAccRej, a.particle0 + (uint64_t)pid, 0);
            if (u > prob || anyinf) {
                acc = 0;
                SMCB_LOCAL(i) {
                    if (gd(i) < d_) {
                        a.x_new[pid * d_ + gd(i)] = a.x[pid * d_ + gd(i)];
                        a.r_new[pid * d_ + gd(i)] = a.r[pid * d_ + gd(i)];
                    }
                }
                As = A0; Bs = B0; ken = ke0;
            }
        }
        if (sub == 0) {
            if (a.A_old) a.A_old[pid] = A0;
            if (a.B_old) a.B_old[pid] = B0;
            if (a.A_new) a.A_new[pid] = As;
            if (a.B_new) a.B_new[pid] = Bs;
            if (a.ke_old) a.ke_old[pid] = ke0;
            if (a.ke_new) a.ke_new[pid] = ken;
            if (a.n_leapfrog) a.n_leapfrog[pid] = (int)n_leapfrog;
            if (a.accepted) a.accepted[pid] = acc;
            if (a.depth) a.depth[pid] = depth;
            if (a.accept_stat) a.accept_stat[pid] = n_leapfrog ? ws[6] / (double)n_leapfrog : 0.0;
        }
        phase = kIdle;
        return true;
    }
#undef SMCB_LOCAL
#undef SMCB_PAIRS
#undef nlp
};

}  // namespace smcb
