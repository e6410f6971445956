"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts.distributions import StdNormal
from smcnuts.model.device_model import make_model
from smcnuts.smc_sampler import SMCSampler

for name, kw, eps, lk, temp, N, K, rs in (("arma", {}, 0.01, "forwardsLKernel", False, 700, 3, "multinomial"),
                                          ("arma", {}, 0.01, "asymptoticLKernel", True, 300, 2, "systematic"),
                                          ("PRMwCD", {}, 0.01, "asymptoticLKernel", True, 64, 2, "multinomial"),
                                          ("gauss", {"dim": 8}, 0.1, "GaussianApproxLKernel", False, 500, 2, "systematic"),
                                          ("gauss", {"dim": 100}, 0.1, "GaussianApproxLKernel", False, 400, 2, "systematic"),
                                          ("gauss", {"dim": 110}, 0.1, "forwardsLKernel", False, 64, 1, "multinomial"),
                                          ("gauss", {"dim": 128}, 0.1, "forwardsLKernel", False, 64, 2, "multinomial"),
                                          ("gauss", {"dim": 120}, 0.1, "GaussianApproxLKernel", False, 300, 1, "systematic")):
    m = make_model(name, **kw)
    s = SMCSampler(K=K, N=N, target=m, step_size=eps, sample_proposal=StdNormal(m.dim), momentum_proposal=StdNormal(m.dim),
                   lkernel=lk, tempering=temp, rng=3, resampling=rs)
    s.samples.ess = 0  # force a resample in the first iteration
    s.sample(show_progress=False)
    assert np.all(np.isfinite(s.mean_estimate)), name
    print(name, kw, lk, "ok", s.leapfrogs)
print("SANITIZE_SMOKE_OK")
