// K8 -- device-resident replay of scipy.optimize.bisect for the adaptive temperature
// (/root/reference/smcnuts/tempering/adaptive_tempering.py:58-63: phi = 1 if ESS(1) >= alpha N, else
//  bisect(ess_minus_target, old_phi, 1.0) with scipy's defaults xtol = 2e-12, rtol = 4 eps, maxiter = 100).
//
// scipy's iteration (Zeros/bisect.c) is   dm *= .5; xm = xa + dm; fm = f(xm); if (fm*f1 >= 0) xa = xm;
// stop when fm == 0 or |dm| < xtol + rtol |xm|.  Its iterates are a deterministic function of the bracket and of the
// SIGNS of f at dyadic points, so the next `depth` levels can be evaluated speculatively: one pass over the particles
// computes the objective at all 2^depth - 1 midpoints of the next levels (heap order), then one thread walks the
// evaluated tree.  Both steps live on the device; the host enqueues a fixed schedule of passes and reads the result
// once -- no round trip per pass (round 1 walked the tree on the host: 25 % of a PRMwCD iteration on a 2^17 shard).
//
// The functions are SMCB_HD so that tests/hostsim compiles the very same walk with g++ and checks it against scipy
// bit for bit on the CPU.
#pragma once
#include "common.cuh"

namespace smcb {

constexpr int kBisectMaxCand = 16;     // candidates per pass: 2 + 7 in the first pass, 15 afterwards
constexpr int kBisectFirstDepth = 3;   // levels evaluated together with the bracket ends
constexpr int kBisectDepth = 4;        // levels per later pass
constexpr int kBisectPasses = 11;      // 3 + 10*4 = 43 levels >= the 39-41 halvings xtol = 2e-12 can need on [0, 1]

enum BisectStatus : int {
    kBisectRunning = 0,
    kBisectDone = 1,           // root found (or phi = 1 accepted)
    kBisectNaN = 2,            // "The function value at x=... is NaN; solver cannot continue."
    kBisectSameSign = 3,       // "f(a) and f(b) must have different signs"
    kBisectNoConvergence = 4   // maxiter exceeded
};

struct BisectState {
    double xa, dm, f1;         // bracket start, bracket width, f(start)
    double result;             // the root (valid when status == kBisectDone)
    double nan_at;             // abscissa of the NaN (status == kBisectNaN)
    double xtol, rtol;
    double target;             // alpha * N_total: f = ESS - target
    int status, iterations, maxiter, first;
    int n_cand, depth;         // candidates of the pending pass and the number of tree levels they cover
    double cand[kBisectMaxCand];
};

// midpoint held by heap node i (1 = root) of the bisection tree over [xa, xa + dm): follow the bits of i below its
// leading one; a set bit moves the bracket start to the midpoint.  Same operations, same order as the sequential loop.
SMCB_HD double bisect_node(double xa, double dm, int i) {
    int top = 0;
    while ((i >> (top + 1)) != 0) ++top;
    double xm = xa;
    for (int b = top; b >= 0; --b) {
        dm *= 0.5;
        xm = xa + dm;
        if (b > 0 && ((i >> (b - 1)) & 1)) xa = xm;
    }
    return xm;
}

SMCB_HD void bisect_fill_candidates(BisectState& s) {
    int n = 0;
    if (s.first) {
        s.cand[n++] = s.result;      // the bracket end xb (parked in `result`, which is also the answer when f(xb) >= 0)
        s.cand[n++] = s.xa;
    }
    const int nodes = (1 << s.depth) - 1;
    for (int i = 1; i <= nodes; ++i) s.cand[n++] = bisect_node(s.xa, s.dm, i);
    s.n_cand = n;
}

SMCB_HD void bisect_init(BisectState& s, double xa, double xb, double target, double xtol, double rtol, int maxiter) {
    s.xa = xa; s.dm = xb - xa; s.f1 = 0.0; s.result = xb; s.nan_at = 0.0;
    s.xtol = xtol; s.rtol = rtol; s.target = target;
    s.status = kBisectRunning; s.iterations = 0; s.maxiter = maxiter; s.first = 1;
    s.depth = kBisectFirstDepth;
    bisect_fill_candidates(s);
}

// f[j] = objective at s.cand[j].  Walks the evaluated levels and prepares the next pass.
SMCB_HD void bisect_advance(BisectState& s, const double* f) {
    if (s.status != kBisectRunning) return;
    const double* fn = f;            // node values in heap order (fn[i - 1] belongs to node i)
    if (s.first) {
        const double f2 = f[0], f1 = f[1];
        const double xb = s.cand[0];
        fn = f + 2;
        s.first = 0;
        if (f2 >= 0) { s.result = xb; s.status = kBisectDone; return; }            // adaptive_tempering.py:58-59
        if (f1 != f1 || f2 != f2) { s.nan_at = (f1 != f1) ? s.xa : xb; s.status = kBisectNaN; return; }
        if (f1 == 0) { s.result = s.xa; s.status = kBisectDone; return; }
        if ((f1 > 0) == (f2 > 0)) { s.status = kBisectSameSign; return; }
        s.f1 = f1;
    }
    int i = 1;
    for (int lvl = 0; lvl < s.depth; ++lvl) {
        s.dm *= 0.5;
        const double xm = s.xa + s.dm;
        const double fm = fn[i - 1];
        if (fm != fm) { s.nan_at = xm; s.status = kBisectNaN; return; }
        ++s.iterations;
        int nxt = 2 * i;
        if (fm * s.f1 >= 0) { s.xa = xm; nxt = 2 * i + 1; }
        const double axm = xm < 0 ? -xm : xm, adm = s.dm < 0 ? -s.dm : s.dm;
        if (fm == 0 || adm < s.xtol + s.rtol * axm) { s.result = xm; s.status = kBisectDone; return; }
        if (s.iterations >= s.maxiter) { s.status = kBisectNoConvergence; return; }
        i = nxt;
    }
    s.depth = kBisectDepth;
    bisect_fill_candidates(s);
}

}  // namespace smcb
