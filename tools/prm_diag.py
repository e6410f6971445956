"""Diagnostic: PRMwCD tensor-core group kernel vs one-lane-per-particle kernel vs C oracle."""
import math
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from oracle import smc_oracle as O  # noqa: E402
from smcnuts import _device as dev  # noqa: E402
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.proposal.nuts_acc_rej import NUTSProposalWithAccRej  # noqa: E402

m, t = make_model("PRMwCD"), O.COracleTarget("PRMwCD")
rng = np.random.default_rng(21)
N = 6000
centre = np.array([0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014, 1.6721, -0.1868,
                   -0.1491, math.log(0.3326)])
x = centre + rng.normal(size=(N, 13)) * 0.05
x[:40] = rng.normal(size=(40, 13)) * 3.0
x[40:60, 0] = -800.0
r = rng.normal(size=(N, 13))
outs = {}
for mode in ("1", "0"):
    os.environ["SMCB_PRM_SCALAR"] = mode
    k = NUTSProposalWithAccRej(m, StdNormal(13), 0.01, rng=3)
    o = k.transition(dev.to_device(x), dev.to_device(r), 0.6, iteration=1)
    outs[mode] = {kk: v.cpu().numpy() for kk, v in o.items()}
ref = t.nuts_batch(x, r, 0.01, 0.6, 10, seed=3, iteration=1, accrej=True, nthreads=8)
a, b = outs["1"], outs["0"]
print("scalar==group", (a["n_leapfrog"] == b["n_leapfrog"]).mean(), "scalar==oracle", (a["n_leapfrog"] == ref["n_leapfrog"]).mean(),
      "group==oracle", (b["n_leapfrog"] == ref["n_leapfrog"]).mean())
bad = np.nonzero(a["n_leapfrog"] != b["n_leapfrog"])[0]
print("mismatch idx (first 30)", bad[:30], "count<60:", (bad < 60).sum(), "of", len(bad))
for i in bad[:12]:
    print(i, "scalar", a["n_leapfrog"][i], a["depth"][i], "group", b["n_leapfrog"][i], b["depth"][i], "oracle", ref["n_leapfrog"][i],
          "A_old", a["A_old"][i], b["A_old"][i], "B_old", a["B_old"][i], b["B_old"][i])
fin = np.isfinite(a["B_old"]) & np.isfinite(b["B_old"])
print("max rel A_old", np.max(np.abs(a["A_old"][fin] - b["A_old"][fin]) / np.abs(a["A_old"][fin])),
      "max rel B_old", np.max(np.abs(a["B_old"][fin] - b["B_old"][fin]) / np.abs(a["B_old"][fin])))
print("isfinite mismatch", np.nonzero(np.isfinite(a["B_old"]) != np.isfinite(b["B_old"]))[0][:20])
same = a["n_leapfrog"] == b["n_leapfrog"]
d = np.abs(a["x_new"][same] - b["x_new"][same]).max(axis=1)
print("x_new diff quantiles (same trees)", np.quantile(d[np.isfinite(d)], [0.5, 0.9, 0.99, 1.0]))
