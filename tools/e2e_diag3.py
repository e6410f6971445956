"""Per-call wall times of the first 60 host-path calls after a device-resident SMC run (how long is the cold transient?)."""
import sys
import time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.proposal.nuts import NUTSProposal  # noqa: E402
from smcnuts.smc_sampler import SMCSampler  # noqa: E402
N = 1 << 20
m = make_model("arma")
s = SMCSampler(K=25, N=N, target=m, step_size=0.01, sample_proposal=StdNormal(4), momentum_proposal=StdNormal(4),
               lkernel="forwardsLKernel", tempering=False, rng=10)
s.sample(show_progress=False)
x_host = s.samples.x.cpu().pin_memory()
r_host = torch.empty_like(x_host).pin_memory()
r_host.copy_(StdNormal(4, seed=11).rvs(N, iteration=0, particle0=0))
fk = NUTSProposal(m, StdNormal(4), 0.01, rng=10)
ts = []
for i in range(60):
    t0 = time.perf_counter()
    fk.rvs(x_host, r_host, 1.0)
    lf = int(fk.last["n_leapfrog"].sum().item())
    ts.append((time.perf_counter() - t0) * 1e3)
print("per-call ms:", [round(t, 2) for t in ts])
# same with a fresh proposal object and fresh host tensors
x2, r2 = x_host.clone().pin_memory(), r_host.clone().pin_memory()
fk2 = NUTSProposal(m, StdNormal(4), 0.01, rng=10)
ts = []
for i in range(20):
    t0 = time.perf_counter()
    fk2.rvs(x2, r2, 1.0)
    lf = int(fk2.last["n_leapfrog"].sum().item())
    ts.append((time.perf_counter() - t0) * 1e3)
print("fresh object + fresh pinned inputs:", [round(t, 2) for t in ts])
