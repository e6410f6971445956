"""bench.py's e2e leg in isolation, with and without its clock-sampler thread (is the Python-driven copy/compute pipeline
sensitive to the second thread?)."""
import sys
import time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
import bench  # noqa: E402
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
for tag in ("no sampler thread", "NVML sampler thread", "no sampler thread", "nvidia-smi sampler thread"):
    clocks = None
    if "NVML" in tag:
        clocks = bench.ClockSampler(0); clocks.start()
    if "nvidia-smi" in tag:
        import os
        os.environ["SMCB_BENCH_NVIDIA_SMI"] = "1"
        clocks = bench.ClockSampler(0); clocks.start()
    fk = NUTSProposal(m, StdNormal(4), 0.01, rng=10)
    for _ in range(3):
        fk.rvs(x_host, r_host, 1.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); lf = 0
    for _ in range(20):
        fk.rvs(x_host, r_host, 1.0)
        lf += int(fk.last["n_leapfrog"].sum().item())
    dt = time.perf_counter() - t0
    if clocks:
        clocks.stop()
    print(f"{tag:28s}: {dt / 20 * 1e3:.3f} ms per call, {lf / dt / 1e9:.3f} G grad-evals/s", flush=True)
