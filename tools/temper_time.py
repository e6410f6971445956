"""A/B timing of the adaptive-temperature search: device-resident bisection vs the round-1 host walk.

    python tools/temper_time.py [log2n]      (under torchrun: sharded, all-gather per pass)
"""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
import torch.distributed as dist  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.parallel import ShardContext  # noqa: E402
from smcnuts.tempering.adaptive_tempering import ESSTempering  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 17
n = 1 << lg
m = make_model("PRMwCD")
x = StdNormal(13, seed=3).rvs(n, iteration=0, particle0=n * int(os.environ.get("RANK", "0"))) * 0.3
A, B = m.split(x)
ts = ESSTempering(n * world, m, alpha=0.5, shard=ShardContext())
for name, fn in (("device", ts.calculate_phi_from_split), ("host", ts.calculate_phi_from_split_host)):
    fn(A, B, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        phi = fn(A, B, 0.0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"tempering search ({name} walk), {n} particles/rank x {world} ranks: {dt * 1e3:.3f} ms per call, phi = {phi!r}, "
              f"passes = {ts.passes}", flush=True)
if world > 1:
    dist.destroy_process_group()
