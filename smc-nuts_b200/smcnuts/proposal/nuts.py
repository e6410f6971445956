"""NUTSProposal -- batched iterative NUTS on the GPU behind the reference's proposal plugin API.

Mirrors smcnuts/proposal/nuts.py of the reference: `NUTSProposal(target, momentum_proposal, step_size, rng)`,
`.rvs(x_cond, r_cond, phi) -> (x_prime, r_prime)`, `.logpdf(r)`.  The per-particle Python loop and the
recursive build_tree (nuts.py:50-53,114-150) are one persistent CUDA kernel (csrc/nuts_kernel.cu).
"""
import math

import torch

from .. import _cabi, _device as dev

# Set max tree depth of the NUTS tree (reference: nuts.py:4)
MAX_TREE_DEPTH = 10


class NUTSProposal:
    accept_reject = False

    def __init__(self, target, momentum_proposal, step_size, rng=None, max_tree_depth=MAX_TREE_DEPTH):
        if not hasattr(target, "handle"):
            raise TypeError("the device NUTS proposal needs a device model (smcnuts.model.device_model / "
                            "smcnuts.model.bridgestan.StanModel); Python targets cannot run on the GPU path")
        self.target = target
        self.momentum_proposal = momentum_proposal
        self.step_size = float(step_size)
        self.rng = rng
        self.seed = dev.seed_from_rng(rng)
        self.max_tree_depth = int(max_tree_depth)
        self.iteration = 0        # Philox iteration key; SMCSampler sets it, standalone calls auto-increment
        self.particle0 = 0        # global index of local particle 0 (multi-GPU shards)
        self.last = None          # per-particle by-products of the last transition (device tensors)
        self.record_events = False  # bench.py: CUDA events tightly around the kernel launch
        self.events = []

    def rvs(self, x_cond, r_cond, phi: float = 1.0):
        """Propagate particles through one NUTS transition each.  numpy in -> numpy out; CUDA tensors stay put."""
        x = dev.to_device(x_cond).reshape(-1, self.target.dim)
        r = dev.to_device(r_cond).reshape(-1, self.target.dim)
        out = self.transition(x, r, phi)
        self.iteration += 1
        return dev.to_host_like(out["x_new"], x_cond, "x_new"), dev.to_host_like(out["r_new"], r_cond, "r_new")

    def transition(self, x, r, phi=1.0, iteration=None, carry=None, want_grad=False):
        """Device entry point: returns dict of device tensors (x_new, r_new, A_old, B_old, A_new, B_new,
        ke_old, ke_new, n_leapfrog, accepted, depth[, g_new]).

        carry = (A, B, g) at `x` (the previous transition's A_new, B_new, g_new at the same phi) skips the initial
        model evaluation of every transition; want_grad=True also returns g_new for the next call."""
        N, D = x.shape
        it = self.iteration if iteration is None else iteration
        h = self.target.handle
        nbytes = _cabi._ll()
        _cabi.call("smcb_nuts_workspace_bytes", h, N, self.max_tree_depth, nbytes)
        ws = dev.workspace("nuts", nbytes.value)
        o = dict(x_new=dev.empty(N, D), r_new=dev.empty(N, D), A_old=dev.empty(N), B_old=dev.empty(N),
                 A_new=dev.empty(N), B_new=dev.empty(N), ke_old=dev.empty(N), ke_new=dev.empty(N),
                 n_leapfrog=dev.empty(N, dtype=torch.int32), accepted=dev.empty(N, dtype=torch.int32),
                 depth=dev.empty(N, dtype=torch.int32))
        if self.accept_reject:
            carry, want_grad = None, False
        if want_grad:
            o["g_new"] = dev.empty(N, D)
        cA, cB, cg = carry if carry is not None else (None, None, None)
        if self.record_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _cabi.call("smcb_nuts_transition", h, dev.ptr(x), dev.ptr(r), N, self.step_size, float(phi),
                   self.max_tree_depth, int(self.accept_reject), self.seed, it, self.particle0,
                   dev.ptr(o["x_new"]), dev.ptr(o["r_new"]), dev.ptr(o["A_old"]), dev.ptr(o["B_old"]),
                   dev.ptr(o["A_new"]), dev.ptr(o["B_new"]), dev.ptr(o["ke_old"]), dev.ptr(o["ke_new"]),
                   dev.ptr(o["n_leapfrog"]), dev.ptr(o["accepted"]), dev.ptr(o["depth"]), dev.ptr(cA), dev.ptr(cB),
                   dev.ptr(cg), dev.ptr(o.get("g_new")), dev.ptr(ws), ws.numel(), dev.stream_ptr())
        if self.record_events:
            e1.record()
            self.events.append((e0, e1))
        self.last = o
        return o

    def logpdf(self, r):
        """Log density of the forward kernel, i.e. of the momentum (nuts.py:177-189)."""
        if dev.is_std_normal(self.momentum_proposal, self.target.dim):
            rd = dev.to_device(r).reshape(-1, self.target.dim)
            out = dev.empty(rd.shape[0])
            _cabi.call("smcb_std_normal_logpdf", dev.ptr(rd), rd.shape[0], self.target.dim, dev.ptr(out),
                       dev.stream_ptr())
            return dev.like_input(out, r)
        return self.momentum_proposal.logpdf(r)
