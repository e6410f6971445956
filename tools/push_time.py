"""Kernel-level timing of the fused resample + migration kernel (smcb_resample_systematic_push) under torchrun:
CUDA events tightly around the launch, for weight profiles that make 0 % ... ~50 % of the rows change GPU.

    torchrun --nproc-per-node 2 tools/push_time.py [log2 rows per rank]
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from smcnuts import _cabi, _device as dev  # noqa: E402
from smcnuts.parallel import ShardContext  # noqa: E402
from smcnuts.samples.samples import Resampler, normalise  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 25
n, D = 1 << lg, 16
N = n * world
sh = ShardContext()
x = torch.randn(n, D, dtype=torch.float64, device="cuda")
events = []
orig_call = _cabi.call


def timed_call(name, *a):
    if name == "smcb_resample_systematic_push":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig_call(name, *a); e1.record()
        events.append((e0, e1))
        return r
    return orig_call(name, *a)


import smcnuts.samples.samples as S  # noqa: E402
S._cabi.call = timed_call
for slope in (0.0, 0.5, 1.5, 4.0):
    gidx = (torch.arange(n, device="cuda", dtype=torch.float64) + rank * n) / N - 0.5
    logw = torch.randn(n, dtype=torch.float64, device="cuda") + slope * gidx
    rs = Resampler(N, 10, sh, scheme="systematic")
    rs.keep_idx = False
    events.clear()
    tot = []
    for it in range(6):
        wn, stats, _, scan = normalise(logw, sh, scan=True)
        cdf = rs._cdf(wn, scan)
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = rs.resample_from_cdf(x, cdf, it); b.record()
        torch.cuda.synchronize()
        tot.append(a.elapsed_time(b))
    k = [e0.elapsed_time(e1) for e0, e1 in events][2:]
    t = torch.tensor([min(k), min(tot[2:]), float(rs._migrated)], dtype=torch.float64, device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        ks = [float(v[0]) for v in allt]; ts = [float(v[1]) for v in allt]; mg = [int(v[2]) for v in allt]
        gb_in = max(mg) * D * 8 / 1e9
        print(f"slope {slope}: rows received from peers per rank {mg} ({100 * max(mg) / n:.0f} %), push kernel ms per rank "
              f"{[round(v, 3) for v in ks]}, whole resample ms {[round(v, 3) for v in ts]}; "
              f"max in {gb_in:.2f} GB -> {gb_in / (max(ks) * 1e-3):.0f} GB/s of NVLink in per GPU", flush=True)
    rs.close()
dist.destroy_process_group()
