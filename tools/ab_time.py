"""A/B timing of the NUTS kernel on FIXED inputs (same transition repeated; min / median of CUDA-event times).

    python tools/ab_time.py PRMwCD 17,18,20 [reps]      # honours SMCB_PRM_SCALAR etc.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.proposal.nuts import NUTSProposal  # noqa: E402

CENTRE = {
    "arma": ([0.0068, 0.957, -0.034, float(np.log(0.1666))], 0.02, 0.01, {}),
    "PRMwCD": ([0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014, 1.6721, -0.1868, -0.1491,
                float(np.log(0.3326))], 0.02, 0.01, {}),
    "gauss": ([0.0] * 100, 1.0, 0.1, {"dim": 100}),
}

name = sys.argv[1]
sizes = [int(v) for v in sys.argv[2].split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
centre, spread, eps, kw = CENTRE[name]
m = make_model(name, **kw)
D = m.dim
for lg in sizes:
    N = 1 << lg
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g) * spread + torch.tensor(centre, dtype=torch.float64, device="cuda")
    k = NUTSProposal(m, StdNormal(D), eps, rng=10)
    # two transitions to reach a typical state, then time the third repeatedly
    for it in range(2):
        x = k.transition(x, StdNormal(D, seed=10).rvs(N, iteration=it), 1.0, iteration=it)["x_new"]
    r = StdNormal(D, seed=10).rvs(N, iteration=2)
    ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = k.transition(x, r, 1.0, iteration=2); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = ts[1:]
    nl = int(o["n_leapfrog"].sum().item())
    print(f"{name} N=2^{lg}: min {min(ts):8.3f} ms  median {float(np.median(ts)):8.3f} ms  {nl / min(ts) / 1e6:.3f} G grad-evals/s "
          f"(mean {nl / N:.1f}, max {int(o['n_leapfrog'].max().item())})")
