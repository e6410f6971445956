"""Throughput of GENERATED models in the NUTS kernel against the hand-written arma device function (fixed inputs).

    python tools/gen_time.py [log2N] [reps]

Two phrasings of the same ARMA(1,1) density: rolling scalars (tests/stan/arma11.stan) and whole-series vector locals with
one vectorised likelihood statement at the end (tests/stan/arma_series.stan, the Stan manual's style), which costs
per-thread local arrays of values and sensitivities unless the generator fuses the likelihood into the loop.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.model.generated import GeneratedModel  # noqa: E402
from smcnuts.proposal.nuts import NUTSProposal  # noqa: E402

WHOLE_SERIES = (ROOT / "tests/stan/arma_series.stan").read_text()

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
y = json.loads((ROOT / "smc-nuts_b200/smcnuts/data/arma/arma.json").read_text())["y"]
models = [("built-in arma", make_model("arma")),
          ("generated, rolling scalars", GeneratedModel((ROOT / "tests/stan/arma11.stan").read_text(), {"T": 200, "y": y}, "arma11")),
          ("generated, whole-series vectors", GeneratedModel(WHOLE_SERIES, {"T": 200, "y": y}, "arma_whole"))]
N, D = 1 << lg, 4
g = torch.Generator(device="cuda"); g.manual_seed(1)
x0 = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g) * 0.02 + \
    torch.tensor([0.0068, 0.957, -0.034, float(np.log(0.1666))], dtype=torch.float64, device="cuda")
for name, m in models:
    k = NUTSProposal(m, StdNormal(D), 0.01, rng=10)
    x = x0
    for it in range(2):
        x = k.transition(x, StdNormal(D, seed=10).rvs(N, iteration=it), 1.0, iteration=it)["x_new"]
    r = StdNormal(D, seed=10).rvs(N, iteration=2)
    ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = k.transition(x, r, 1.0, iteration=2); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    nl = int(o["n_leapfrog"].sum().item())
    print(f"{name:34s} N=2^{lg}: min {min(ts[1:]):8.3f} ms  {nl / min(ts[1:]) / 1e6:.3f} G grad-evals/s (mean {nl / N:.1f} leapfrogs)", flush=True)
