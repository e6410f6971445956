// TEST TOOL ONLY: a single-threaded, fiber-based emulator of ONE 32-lane warp, so that the device code that uses
// warp-level primitives (the 4-lanes-per-particle NUTS kernels: PrmModelG / GaussModelG in csrc/models.cuh and the
// group fold of csrc/nuts_lane.cuh) can run on the CPU in the `not gpu` test-suite.
//
// Every lane is a ucontext fiber running the same function.  A warp-level primitive deposits its operand in an exchange
// buffer and waits at a barrier: the warp barrier (32 participants) for full-mask calls, the barrier of the lane's
// particle group (4 participants) for the group-masked shuffles of Lane::gsum -- groups run independent control flow
// between the warp-wide points, exactly as on the device.  A fiber that cannot proceed yields to the scheduler, which
// resumes the lanes round-robin; a full round without progress is a deadlock and aborts.
//
// mma.m8n8k4 is emulated with a sequential FMA-free sum over k (the hardware's internal order is not specified): results
// agree with the device to rounding, not bit for bit.  Never compiled into, nor reachable from, the product library.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt_emu {

constexpr int kLanes = 32, kGroup = 4, kBarriers = kLanes / kGroup + 1, kWarpBarrier = kLanes / kGroup;

struct Warp {
    ucontext_t main_ctx, ctx[kLanes];
    std::vector<char> stacks;
    bool done[kLanes];
    int cur = 0;
    bool progressed = false;
    int arrived[kBarriers] = {0};
    unsigned generation[kBarriers] = {0};
    double dbuf[kLanes], abuf[kLanes], bbuf[kLanes];
    unsigned long long ubuf[kLanes];
    int pbuf[kLanes];
    std::function<void(int)> body;
};

inline Warp*& current() {
    static Warp* w = nullptr;
    return w;
}

inline int lane_id() { return current()->cur; }

inline void yield() {
    Warp* w = current();
    swapcontext(&w->ctx[w->cur], &w->main_ctx);
}

inline void barrier(int id, int participants) {
    Warp* w = current();
    const unsigned gen = w->generation[id];
    w->progressed = true;
    if (++w->arrived[id] == participants) {
        w->arrived[id] = 0;
        ++w->generation[id];
    } else {
        while (w->generation[id] == gen) yield();
    }
}

// barrier for a shuffle/vote mask issued by the current lane: full warp, or the lane's particle group
inline void sync_mask(unsigned mask) {
    if (mask == 0xffffffffu) {
        barrier(kWarpBarrier, kLanes);
    } else {
        const int g = lane_id() / kGroup;
        if (mask != (((1u << kGroup) - 1u) << (g * kGroup))) {
            std::fprintf(stderr, "simt_emu: unsupported mask %08x on lane %d\n", mask, lane_id());
            std::abort();
        }
        barrier(g, kGroup);
    }
}

inline double shfl_double(unsigned mask, double v, int src) {
    Warp* w = current();
    w->dbuf[lane_id()] = v;
    sync_mask(mask);
    const double r = w->dbuf[src & (kLanes - 1)];
    sync_mask(mask);
    return r;
}
inline unsigned long long shfl_u64(unsigned mask, unsigned long long v, int src) {
    Warp* w = current();
    w->ubuf[lane_id()] = v;
    sync_mask(mask);
    const unsigned long long r = w->ubuf[src & (kLanes - 1)];
    sync_mask(mask);
    return r;
}
inline unsigned ballot(unsigned mask, bool pred) {
    Warp* w = current();
    w->pbuf[lane_id()] = pred ? 1 : 0;
    sync_mask(mask);
    unsigned r = 0;
    for (int l = 0; l < kLanes; ++l)
        if (((mask >> l) & 1u) && w->pbuf[l]) r |= 1u << l;
    sync_mask(mask);
    return r;
}

// C[8x8] += A[8x4] B[4x8]; lane l holds A[l/4][l%4], B[l%4][l/4] and C[l/4][2(l%4) + {0,1}]
inline void dmma(double& c0, double& c1, double a, double b) {
    Warp* w = current();
    const int l = lane_id();
    w->abuf[l] = a;
    w->bbuf[l] = b;
    barrier(kWarpBarrier, kLanes);
    const int row = l / 4, col = 2 * (l % 4);
    for (int k = 0; k < 4; ++k) {
        c0 += w->abuf[row * 4 + k] * w->bbuf[col * 4 + k];
        c1 += w->abuf[row * 4 + k] * w->bbuf[(col + 1) * 4 + k];
    }
    barrier(kWarpBarrier, kLanes);
}

inline void fiber_entry() {
    Warp* w = current();
    const int l = w->cur;
    w->body(l);
    w->done[l] = true;
    w->progressed = true;
    swapcontext(&w->ctx[l], &w->main_ctx);
}

// run body(lane) on 32 fibers until all return
inline void run_warp(const std::function<void(int)>& body, size_t stack_bytes = 1 << 20) {
    Warp w;
    w.body = body;
    w.stacks.resize(stack_bytes * kLanes);
    current() = &w;
    for (int l = 0; l < kLanes; ++l) {
        w.done[l] = false;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks.data() + stack_bytes * l;
        w.ctx[l].uc_stack.ss_size = stack_bytes;
        w.ctx[l].uc_link = &w.main_ctx;
        makecontext(&w.ctx[l], fiber_entry, 0);
    }
    for (;;) {
        bool all = true;
        w.progressed = false;
        for (int l = 0; l < kLanes; ++l) {
            if (w.done[l]) continue;
            all = false;
            w.cur = l;
            swapcontext(&w.main_ctx, &w.ctx[l]);
        }
        if (all) break;
        if (!w.progressed) {
            std::fprintf(stderr, "simt_emu: deadlock (no lane made progress)\n");
            std::abort();
        }
    }
    current() = nullptr;
}

}  // namespace simt_emu

// ---- the CUDA spellings the device headers use, for the emulated build only
struct EmuThreadIdx {
    struct X {
        operator unsigned() const { return (unsigned)simt_emu::lane_id(); }
    } x;
};
static EmuThreadIdx threadIdx;

inline double __shfl_xor_sync(unsigned mask, double v, int lane_mask) {
    return simt_emu::shfl_double(mask, v, simt_emu::lane_id() ^ lane_mask);
}
inline unsigned __shfl_xor_sync(unsigned mask, unsigned v, int lane_mask) {
    return (unsigned)simt_emu::shfl_u64(mask, v, simt_emu::lane_id() ^ lane_mask);
}
inline double __shfl_sync(unsigned mask, double v, int src) { return simt_emu::shfl_double(mask, v, src); }
inline unsigned long long __shfl_sync(unsigned mask, unsigned long long v, int src) { return simt_emu::shfl_u64(mask, v, src); }
inline unsigned __ballot_sync(unsigned mask, bool pred) { return simt_emu::ballot(mask, pred); }
inline bool __any_sync(unsigned mask, bool pred) { return simt_emu::ballot(mask, pred) != 0u; }
inline bool __all_sync(unsigned mask, bool pred) { return simt_emu::ballot(mask, pred) == mask; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)simt_emu::ballot(mask, true); }
inline int __double2hiint(double v) {
    uint64_t b;
    std::memcpy(&b, &v, 8);
    return (int)(b >> 32);
}
namespace smcb {
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
}  // namespace smcb
