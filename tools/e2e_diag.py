"""Why is bench.py's e2e call slower than tools/e2e_time.py's?  Same NUTSProposal.rvs call on (a) the tight synthetic cloud of
e2e_time.py, (b) the particle set of a 25-iteration SMC run (what bench.py uses), each with per-call wall times."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.proposal.nuts import NUTSProposal  # noqa: E402
from smcnuts.smc_sampler import SMCSampler  # noqa: E402

N = 1 << 20
m = make_model("arma")
rng = np.random.default_rng(1)
xa = torch.from_numpy(rng.normal(size=(N, 4)) * 0.02 + np.array([0.0068, 0.957, -0.034, np.log(0.1666)])).pin_memory()
ra = torch.from_numpy(rng.normal(size=(N, 4))).pin_memory()
s = SMCSampler(K=25, N=N, target=m, step_size=0.01, sample_proposal=StdNormal(4), momentum_proposal=StdNormal(4),
               lkernel="forwardsLKernel", tempering=False, rng=10)
s.sample(show_progress=False)
xb = s.samples.x.cpu().pin_memory()
rb = torch.empty_like(xb).pin_memory()
rb.copy_(StdNormal(4, seed=11).rvs(N, iteration=0, particle0=0))
for tag, x, r in (("synthetic cloud", xa, ra), ("SMC particle set", xb, rb), ("synthetic cloud", xa, ra)):
    k = NUTSProposal(m, StdNormal(4), 0.01, rng=10)
    ts, nl = [], 0
    for it in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k.rvs(x, r, 1.0)
        nl = int(k.last["n_leapfrog"].sum().item())
        ts.append((time.perf_counter() - t0) * 1e3)
    # kernel-only time of the same state, device resident
    xd, rd = x.cuda(), r.cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k.transition(xd, rd, 1.0, iteration=0); torch.cuda.synchronize()
    e0.record(); k.transition(xd, rd, 1.0, iteration=0); e1.record(); torch.cuda.synchronize()
    print(f"{tag:18s}: per-call ms {[round(t, 2) for t in ts]}  leapfrogs/particle {nl / N:.2f}  device-resident transition {e0.elapsed_time(e1):.3f} ms",
          flush=True)
