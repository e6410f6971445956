import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts import _cabi, _device as dev
from smcnuts.parallel import ShardContext
from smcnuts.samples.samples import Resampler, normalise
n, D = 1 << 25, 16
st = dev.stream_ptr()
x = torch.randn(n, D, dtype=torch.float64, device="cuda")
for label, scale in (("mild", 1.0), ("degenerate", 6.0)):
    logw = torch.randn(n, dtype=torch.float64, device="cuda") * scale
    sh = ShardContext()
    wn, stats, _ = normalise(logw, sh)
    print(label, "ess/N", stats[1].item() / n)
    rs = Resampler(n, 10, sh, scheme="systematic"); rs.keep_idx = False
    cdf = rs._cdf(wn)
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = rs.resample_from_cdf(x, cdf, rep); b.record(); torch.cuda.synchronize()
        print(f"  rep {rep}: events {a.elapsed_time(b):.3f} ms, wall {(time.perf_counter() - t0) * 1e3:.3f} ms")
        del out
