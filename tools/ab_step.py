"""A/B of host-side per-iteration overheads: whole SMC iterations (arma, forward L-kernel) timed with CUDA events, with and
without (a) the pre-allocation of the transition's outputs before the ESS synchronisation, (b) the fused reweight kernel.

    python tools/ab_step.py [log2 N] [iterations]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts import _cabi, _device as dev  # noqa: E402
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.proposal.nuts import NUTSProposal  # noqa: E402
from smcnuts.samples.samples import Samples  # noqa: E402
from smcnuts.smc_sampler import SMCSampler  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 25
m = make_model("arma")
from smcnuts.estimate.estimate import Estimate  # noqa: E402
new_prepare, new_reweight, new_estimate = NUTSProposal.prepare, Samples._non_asympototic_reweight, Estimate.return_estimate


def old_reweight(self):
    st, n = dev.stream_ptr(), self.n_local
    lp_x = self.target.combine(*self._split_x, 1.0)
    lp_xnew = self.target.combine(*self._split_new, 1.0)
    out = dev.empty(n)
    _cabi.call("smcb_reweight_forward_ke", dev.ptr(self.logw), dev.ptr(lp_x), dev.ptr(lp_xnew), dev.ptr(self._ke[0]),
               dev.ptr(self._ke[1]), n, dev.ptr(out), st)
    return out


for tag, prep, rew, est in (("old", False, False, False), ("prepare + fused reweight", True, True, False),
                            ("+ one-pass moments", True, True, True), ("old", False, False, False),
                            ("prepare + fused reweight", True, True, False), ("+ one-pass moments", True, True, True)):
    NUTSProposal.prepare = new_prepare if prep else (lambda self, N, D, want_grad=False: None)
    Samples._non_asympototic_reweight = new_reweight if rew else old_reweight
    Estimate.return_estimate = new_estimate if est else (lambda self, x, wn, center=None: self._estimate(x, wn, self._constrain))
    s = SMCSampler(K=K, N=1 << lg, target=m, step_size=0.01, sample_proposal=StdNormal(4), momentum_proposal=StdNormal(4),
                   lkernel="forwardsLKernel", tempering=False, rng=10, save_history=False)
    s.reweight_strategy = None
    s.begin()
    for k in range(5):
        s.iterate(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(5, K):
        s.iterate(k)
    b.record(); torch.cuda.synchronize()
    print(f"{tag:28s}: {a.elapsed_time(b) / (K - 5):.4f} ms per iteration", flush=True)
