"""Run under torchrun with P ranks: the particle-sharded sampler must reproduce the single-GPU run.

Philox streams are keyed by GLOBAL particle index, so sharding changes only the order of the global
reductions (log-sum-exp triples, moment sums) and routes resampled rows through all-to-all-v.
Rank 0 also runs the same configuration unsharded on its own GPU (in a separate, ungrouped context) and
compares.  Prints 'MULTI_GPU_OK' on success.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))


def run(model, kw, N, K, eps, lk, temp, resampling, extra=None):
    from smcnuts.distributions import StdNormal
    from smcnuts.model.device_model import make_model
    from smcnuts.smc_sampler import SMCSampler
    if model == "generated":     # a Stan-subset program compiled into a model plug-in (every rank builds / loads the same one)
        import json
        from smcnuts.model.generated import GeneratedModel
        y = json.loads((ROOT / "smc-nuts_b200/smcnuts/data/arma/arma.json").read_text())["y"]
        m = GeneratedModel((ROOT / "tests/stan/arma11.stan").read_text(), {"T": 200, "y": y}, "arma11")
    else:
        m = make_model(model, **kw)
    s = SMCSampler(K=K, N=N, target=m, step_size=eps, sample_proposal=StdNormal(m.dim), momentum_proposal=StdNormal(m.dim),
                   lkernel=lk, tempering=temp, rng=10, resampling=resampling, **(extra or {}))
    s.sample(show_progress=False)
    return s


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    cases = [("arma", {}, 1 << 14, 6, 0.01, "forwardsLKernel", False, "multinomial"),
             ("arma", {}, 1 << 14, 6, 0.01, "forwardsLKernel", False, "systematic"),
             ("arma", {}, 1 << 13, 5, 0.01, "asymptoticLKernel", True, "systematic"),
             ("arma", {}, 1 << 13, 5, 0.01, "asymptoticLKernel", True, "multinomial"),
             ("gauss", {"dim": 8}, 1 << 13, 4, 0.1, "GaussianApproxLKernel", False, "systematic"),
             # generated model plug-in; step-size + diagonal mass-matrix adaptation (statistics all-reduced over the shards)
             ("generated", {}, 1 << 13, 5, 0.01, "forwardsLKernel", False, "multinomial"),
             ("gauss", {"dim": 8}, 1 << 13, 8, 0.02, "forwardsLKernel", False, "multinomial",
              {"adapt_step_size": 6, "adapt_mass_matrix": True})]
    # single-GPU references first (no process group yet -> ShardContext is a no-op)
    refs = [run(*c) for c in cases] if rank == 0 else None
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for ci, c in enumerate(cases):
        s = run(*c)
        n_local = c[2] // world
        x_last = s.samples.x.contiguous()
        gathered = [torch.empty_like(x_last) for _ in range(world)] if rank == 0 else None
        dist.gather(x_last, gathered, dst=0)
        migrated = torch.tensor([float(s.samples.resampler.last_migrated_rows)], device="cuda")
        dist.all_reduce(migrated)
        if rank == 0:
            r = refs[ci]
            x_all = torch.cat(gathered).cpu().numpy()
            assert np.array_equal(s.leapfrogs, r.leapfrogs), (c, s.leapfrogs, r.leapfrogs)
            np.testing.assert_allclose(s.ess, r.ess, rtol=1e-9)
            np.testing.assert_allclose(s.log_likelihood, r.log_likelihood, rtol=1e-10, atol=1e-9)
            np.testing.assert_allclose(s.phi, r.phi, rtol=1e-9)
            np.testing.assert_allclose(s.mean_estimate, r.mean_estimate, rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(s.variance_estimate, r.variance_estimate, rtol=1e-7, atol=1e-12)
            np.testing.assert_allclose(s.acceptance_rate, r.acceptance_rate, rtol=1e-12)
            assert list(s.resampled) == list(r.resampled)
            np.testing.assert_allclose(s.step_sizes, r.step_sizes, rtol=1e-9)
            if r.metric_scale is not None:
                np.testing.assert_allclose(s.metric_scale, r.metric_scale, rtol=1e-9)
            np.testing.assert_allclose(x_all, r.samples.x.cpu().numpy(), rtol=1e-9, atol=1e-12)
            print(f"case {ci} {c[0]} {c[5]} {c[7]}: ok (resampled {sum(s.resampled)}x, rows migrated in last resample: "
                  f"{int(migrated.item())}, shard {n_local})", flush=True)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
