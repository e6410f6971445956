"""Time the UNMODIFIED reference (imported from /root/reference) on this machine's CPU: BASELINE config 1
(arma, N = 100, K = 10, forward-proposal L-kernel, single thread, as experiments/run_experiments.py) and a larger N.

The reference evaluates the target through BridgeStan, which is not installable offline; the target handed to it here is
the oracle's C density behind the reference's duck-typed target API (one ctypes call per particle per leapfrog, exactly
the call pattern of smcnuts/model/bridgestan.py:28-90).  Runs only where /root/reference exists (the build container);
the numbers are quoted in DESIGN.md section 7 as the "reference Python" baseline beside the oracle-port baseline that
bench.py measures on the GPU box.

    python tools/reference_cpu_time.py > profiles/r2_reference_python_cpu.log
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")
from scipy.stats import multivariate_normal  # noqa: E402

from oracle.smc_oracle import COracleTarget  # noqa: E402
from smcnuts.smc_sampler import SMCSampler  # noqa: E402  (the reference's)


class CountingTarget:
    """The oracle target with the reference's API; counts gradient evaluations (= leapfrog steps + initial points)."""

    def __init__(self, name):
        self.t = COracleTarget(name)
        self.dim = self.t.dim
        self.constrained_dim = self.t.constrained_dim
        self.constrain = self.t.constrain
        self.grad_evals = 0

    def logpdf(self, x, phi=1.0):
        return self.t.logpdf(x, phi)

    def logpdfgrad(self, x, phi=1.0):
        self.grad_evals += 1 if np.ndim(x) == 1 else len(x)
        return self.t.logpdfgrad(x, phi)


print(f"host: {os.cpu_count()} logical cores; the reference is single-threaded")
for name, N, K in (("arma", 100, 10), ("arma", 1000, 5), ("PRMwCD", 100, 3)):
    target = CountingTarget(name)
    rng = np.random.RandomState(10)
    q0 = multivariate_normal(mean=np.zeros(target.dim), cov=np.eye(target.dim), seed=rng)
    smc = SMCSampler(K=K, N=N, target=target, step_size=0.01, sample_proposal=q0, momentum_proposal=q0,
                     lkernel="forwardsLKernel" if name == "arma" else "asymptoticLKernel", tempering=(name != "arma"), rng=rng)
    t0 = time.perf_counter()
    import io, contextlib
    with contextlib.redirect_stderr(io.StringIO()):
        smc.sample()
    dt = time.perf_counter() - t0
    print(f"{name} N={N} K={K}: {dt:.2f} s, {target.grad_evals} gradient evaluations -> {target.grad_evals / dt:.0f} grad-evals/s, "
          f"{K / dt:.3f} SMC iterations/s (reference Python + oracle C density per call)", flush=True)
