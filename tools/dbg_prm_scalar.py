import math, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts import _device as dev
from smcnuts.distributions import StdNormal
from smcnuts.model.device_model import make_model
from smcnuts.proposal.nuts_acc_rej import NUTSProposalWithAccRej
from smcnuts.proposal.nuts import NUTSProposal
m = make_model("PRMwCD")
rng = np.random.default_rng(21)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
wild = len(sys.argv) > 2 and sys.argv[2] == "wild"
centre = np.array([0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014, 1.6721, -0.1868, -0.1491, math.log(0.3326)])
x = centre + rng.normal(size=(N, 13)) * 0.05
if wild:
    x[:40] = rng.normal(size=(40, 13)) * 3.0
    x[40:60, 0] = -800.0
r = rng.normal(size=(N, 13))
k = NUTSProposalWithAccRej(m, StdNormal(13), 0.01, rng=3)
t0 = time.time()
o = k.transition(dev.to_device(x), dev.to_device(r), 0.6, iteration=1)
torch.cuda.synchronize()
print("ok", N, wild, round(time.time() - t0, 3), "s; leapfrogs", int(o["n_leapfrog"].sum().item()), flush=True)
