"""Ad-hoc timing of the host-facing NUTSProposal.rvs path (pinned / numpy inputs) for different chunkings."""
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

N = 1 << 20
m = make_model("arma")
rng = np.random.default_rng(1)
x = rng.normal(size=(N, 4)) * 0.02 + np.array([0.0068, 0.957, -0.034, np.log(0.1666)])
r = rng.normal(size=(N, 4))
xp, rp = torch.from_numpy(x).pin_memory(), torch.from_numpy(r).pin_memory()
SWEEP_BPS = [(((1,) * 8, 4), 1), (((1,) * 12, 4), 1), (((1,) * 16, 4), 1), (((1,) * 8, 2), 2), (((1,) * 12, 2), 2), (((1, 2, 2, 2, 2, 2, 1), 4), 1),
             (((1,) * 24, 4), 1), (((1,) * 8, 3), 1), (((1,) * 12, 3), 1)]
for (chunks, streams), bps in SWEEP_BPS:
    k = NUTSProposal(m, StdNormal(4), 0.01, rng=10)
    k.PIPELINE_FRACTIONS, k.PIPELINE_STREAMS, k.PIPELINE_BLOCKS_PER_SM = chunks, streams, bps
    ts = []
    for it in range(8):
        k.iteration = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k.rvs(xp, rp, 1.0)
        nl = int(k.last["n_leapfrog"].sum().item())
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts[2:]))
    print(f"chunks={len(chunks)} {str(chunks)[:24]:24s} streams={streams} blocks/SM={bps}: {t * 1e3:6.2f} ms  {nl / t / 1e9:.3f} G grad-evals/s", flush=True)
for chunks, streams in (((1,), 1), ((1, 1, 1, 1), 3), ((1, 3, 3, 1), 3), ((1, 3, 3, 1), 4), ((1, 2, 2, 2, 1), 3), ((1, 4, 4, 1), 3),
                        ((1, 7, 7, 1), 3), ((1, 6, 1), 3), ((1, 3, 3, 3, 1), 3), ((1, 5, 5, 4, 1), 4), ((2, 3, 2, 1), 3),
                        ((1, 2, 4, 1), 3), ((1, 2, 3, 2, 1), 4), ((1,) * 8, 3), ((1,) * 8, 4), ((1, 2, 2, 2, 2, 2, 2, 2, 1), 3),
                        ((1,) * 12, 4), ((1,) * 16, 4), ((1, 2, 3, 3, 3, 2, 1, 1), 4)):
    k = NUTSProposal(m, StdNormal(4), 0.01, rng=10)
    k.PIPELINE_FRACTIONS, k.PIPELINE_STREAMS = chunks, streams
    for inp, tag in (((xp, rp), "pinned"),):
        ts = []
        for it in range(8):
            k.iteration = 0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k.rvs(*inp, 1.0)
            nl = int(k.last["n_leapfrog"].sum().item())
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts[2:]))
        print(f"chunks={str(chunks):18s} streams={streams} {tag:6s}: {t * 1e3:6.2f} ms  {nl / t / 1e9:.3f} G grad-evals/s")
# old path for comparison: one launch, serial copies
k = NUTSProposal(m, StdNormal(4), 0.01, rng=10)
k.PIPELINE_MIN_PARTICLES = 1 << 40
for inp, tag in (((xp, rp), "pinned"), ((x, r), "numpy")):
    ts = []
    for it in range(6):
        k.iteration = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k.rvs(*inp, 1.0)
        nl = int(k.last["n_leapfrog"].sum().item())
        ts.append(time.perf_counter() - t0)
    t = min(ts[2:])
    print(f"serial {tag:6s}: {t * 1e3:6.2f} ms  {nl / t / 1e9:.3f} G grad-evals/s")
