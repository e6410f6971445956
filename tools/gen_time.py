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
prm_data = json.loads((ROOT / "smc-nuts_b200/smcnuts/data/PRMwCD/PRMwCD.json").read_text())
prm_data.pop("phi")
ARMA_C = [0.0068, 0.957, -0.034, float(np.log(0.1666))]
PRM_C = [0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014, 1.6721, -0.1868, -0.1491, float(np.log(0.3326))]
cases = [("built-in arma", lambda: make_model("arma"), ARMA_C),
         ("generated arma, rolling scalars", lambda: GeneratedModel((ROOT / "tests/stan/arma11.stan").read_text(), {"T": 200, "y": y}, "arma11"), ARMA_C),
         ("generated arma, whole-series vectors", lambda: GeneratedModel(WHOLE_SERIES, {"T": 200, "y": y}, "arma_whole"), ARMA_C),
         ("built-in PRMwCD (FP64 tensor cores)", lambda: make_model("PRMwCD"), PRM_C),
         ("generated PRM kernel program", lambda: GeneratedModel((ROOT / "tests/stan/prm_kernel.stan").read_text(), prm_data, "prm_kernel"), PRM_C)]
N = 1 << lg
for name, make, centre in cases:
    m = make()
    D = m.dim
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g) * 0.02 + torch.tensor(centre, dtype=torch.float64, device="cuda")
    k = NUTSProposal(m, StdNormal(D), 0.01, rng=10)
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
    print(f"{name:38s} N=2^{lg}: min {min(ts[1:]):8.3f} ms  {nl / min(ts[1:]) / 1e6:.3f} G grad-evals/s (mean {nl / N:.1f} leapfrogs)", flush=True)
