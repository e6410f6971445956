"""ESSTempering -- adaptive temperature by ESS bisection (reference: smcnuts/tempering/adaptive_tempering.py).

The reference calls scipy.optimize.bisect on a Python closure that does three full model sweeps and ~41
O(N) numpy passes.  Here the per-particle quantities come from the split log density the NUTS kernel already
emitted, and each pass evaluates the ESS objective at up to 16 candidate temperatures at once: the candidates are
the nodes of the next 4 levels of the bisection tree, which is a deterministic function of the bracket
(dm *= .5; xm = xa + dm), so walking the evaluated tree reproduces scipy's iterates exactly.

The whole search is DEVICE-RESIDENT (csrc/bisect.cuh): the bracket, the candidates and the outcome live in a small
device state; the host enqueues a fixed schedule of `eval` (one sweep over the particles) -> all-gather of the
per-rank partial states when sharded (stream-ordered NCCL) -> `step` (one thread walks the tree and writes the next
candidates), and reads phi once at the end.  Round 1 walked the tree on the host -- ~10 device->host round trips per
SMC iteration, 25 % of a PRMwCD iteration on a 2^17-particle shard.  `bisect_batched` below is that host walk, kept
as the CPU-testable statement of the algorithm (tests/test_host_logic.py checks it and the device walk, compiled for
the host, against scipy bit for bit).
"""
import math

import numpy as np
import torch

from .. import _cabi, _device as dev
from ..parallel import ShardContext

XTOL, RTOL, MAXITER = 2e-12, 8.881784197001252e-16, 100  # scipy.optimize.bisect defaults


def _merge_triples(t):
    """Merge [P, m, 3] online log-sum-exp states over ranks -> [m, 3] (same rule as the device lse_merge)."""
    t = np.asarray(t, dtype=np.float64)
    out = t[0].copy()
    for p in range(1, t.shape[0]):
        b = t[p]
        with np.errstate(all="ignore"):
            m = np.where(np.isnan(out[:, 0]) | np.isnan(b[:, 0]), np.nan, np.maximum(out[:, 0], b[:, 0]))
            fa = np.where((out[:, 1] == 0) & np.isneginf(out[:, 0]), 0.0, np.exp(out[:, 0] - m))
            fb = np.where((b[:, 1] == 0) & np.isneginf(b[:, 0]), 0.0, np.exp(b[:, 0] - m))
            out = np.stack([m, out[:, 1] * fa + b[:, 1] * fb, out[:, 2] * fa * fa + b[:, 2] * fb * fb], axis=1)
    return out


class ESSTempering:
    def __init__(self, N, target, alpha=0.5, shard: ShardContext = None):
        self.N = N            # GLOBAL number of particles
        self.target = target
        self.alpha = alpha
        self.shard = shard or ShardContext()
        self.passes = 0       # objective launches of the last calculate_phi (diagnostic)

    # ---- reference entry point (adaptive_tempering.py:18-63)
    def calculate_phi(self, args):
        x_new, _p_logpdf_x_new_phi_old, old_phi = args
        A, B = self.target.split(x_new)
        return self.calculate_phi_from_split(A, B, float(old_phi))

    # ---- device entry point: reuse the split the NUTS kernel carried
    def calculate_phi_from_split(self, A, B, old_phi):
        """phi_new from (A, B) at x_new; one device->host read (of the result) in total."""
        L, n, st, sh = _cabi.lib(), A.shape[0], dev.stream_ptr(), self.shard
        mc = L.smcb_bisect_max_candidates()
        state = dev.workspace("bisect_state", L.smcb_bisect_state_bytes())
        _cabi.call("smcb_bisect_init", dev.ptr(state), float(old_phi), 1.0, float(self.N * self.alpha), st)
        self.passes = L.smcb_bisect_passes()
        for _ in range(self.passes):
            tri = dev.empty(mc * 3)
            _cabi.call("smcb_bisect_eval", dev.ptr(A), dev.ptr(B), float(old_phi), n, dev.ptr(state), dev.ptr(tri),
                       dev.ptr(dev.reduce_ws()), st)
            tris = sh.all_gather_vec(tri).contiguous()
            _cabi.call("smcb_bisect_step", dev.ptr(tris), sh.world, dev.ptr(state), st)
        out = dev.empty(4)
        _cabi.call("smcb_bisect_read", dev.ptr(state), dev.ptr(out), st)
        phi, status, _iters, nan_at = out.cpu().tolist()
        status = int(status)
        if status == 1:
            return float(phi)
        if status == 2:
            raise ValueError(f"The function value at x={nan_at} is NaN; solver cannot continue.")
        if status == 3:
            raise ValueError("f(a) and f(b) must have different signs")
        raise RuntimeError("bisect failed to converge")

    # ---- the round-1 host walk (one round trip per pass); kept for A/B timing and as the host statement of the search
    def calculate_phi_from_split_host(self, A, B, old_phi):
        n = A.shape[0]
        logpri, loglik, c = dev.empty(n), dev.empty(n), dev.empty(n)
        _cabi.call("smcb_tempering_arrays", dev.ptr(A), dev.ptr(B), float(old_phi), n, dev.ptr(logpri),
                   dev.ptr(loglik), dev.ptr(c), dev.stream_ptr())
        self.passes = 0

        def ess_minus_target(phis):
            """f(phi) = ESS(phi) - N*alpha for a batch of <= 16 candidates (adaptive_tempering.py:41-56)."""
            m = len(phis)
            ph = torch.tensor(phis, dtype=torch.float64).to(dev.device())
            out = dev.empty(m * 3)
            _cabi.call("smcb_ess_multi_phi", dev.ptr(loglik), dev.ptr(logpri), dev.ptr(c), n, dev.ptr(ph), m,
                       dev.ptr(out), dev.ptr(dev.reduce_ws()), dev.stream_ptr())
            self.passes += 1
            tri = self.shard.all_gather_vec(out).cpu().numpy().reshape(self.shard.world, m, 3)
            tri = _merge_triples(tri)
            with np.errstate(all="ignore"):
                ess = tri[:, 1] * tri[:, 1] / tri[:, 2]
            return ess - self.N * self.alpha

        return bisect_batched(ess_minus_target, old_phi, 1.0)


def _tree_nodes(xa, dm, depth):
    """Heap-ordered midpoints of the next `depth` bisection levels: node i has children 2i (bracket start kept)
    and 2i+1 (bracket start moved to the midpoint)."""
    nodes = {}

    def rec(i, xa_, dm_, d):
        if d == 0:
            return
        dm2 = dm_ * 0.5
        xm = xa_ + dm2
        nodes[i] = xm
        rec(2 * i, xa_, dm2, d - 1)
        rec(2 * i + 1, xm, dm2, d - 1)
    rec(1, xa, dm, depth)
    return nodes


def bisect_batched(f_batch, xa, xb, xtol=XTOL, rtol=RTOL, maxiter=MAXITER, depth=4):
    """scipy.optimize.bisect semantics (Zeros/bisect.c) with speculative batched evaluation.

    Returns 1.0 straight away when f(1.0) >= 0 (adaptive_tempering.py:58-59)."""
    first = _tree_nodes(xa, xb - xa, 3)
    order = sorted(first)
    vals = f_batch([xb, xa] + [first[i] for i in order])
    f2, f1 = float(vals[0]), float(vals[1])
    if f2 >= 0:
        return 1.0
    if math.isnan(f1) or math.isnan(f2):
        raise ValueError(f"The function value at x={xa if math.isnan(f1) else xb} is NaN; solver cannot continue.")
    if f1 == 0:
        return xa
    if np.sign(f1) == np.sign(f2):
        raise ValueError("f(a) and f(b) must have different signs")
    fvals = {i: float(v) for i, v in zip(order, vals[2:])}
    nodes, dm, it, cur_depth = first, xb - xa, 0, 3
    while True:
        i = 1
        for _ in range(cur_depth):
            dm *= 0.5
            xm = xa + dm
            assert xm == nodes[i]
            fm = fvals[i]
            if math.isnan(fm):
                raise ValueError(f"The function value at x={xm} is NaN; solver cannot continue.")
            it += 1
            nxt = 2 * i
            if fm * f1 >= 0:
                xa = xm
                nxt = 2 * i + 1
            if fm == 0 or abs(dm) < xtol + rtol * abs(xm):
                return xm
            if it >= maxiter:
                raise RuntimeError("bisect failed to converge")
            i = nxt
        cur_depth = depth
        nodes = _tree_nodes(xa, dm, cur_depth)
        order = sorted(nodes)
        fvals = {k: float(v) for k, v in zip(order, f_batch([nodes[k] for k in order]))}
