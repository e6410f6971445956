"""CUDA-event time of the Gaussian-approximation L-kernel (moments, factor, log density) at D = 100."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts.lkernel.gaussian_lkernel import GaussianApproxLKernel  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N, D = 1 << lg, 100
m = make_model("gauss", dim=D)
g = torch.Generator(device="cuda"); g.manual_seed(3)
x = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
r = 0.3 * x + torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
lk = GaussianApproxLKernel(target=m, N=N)
ts = []
for _ in range(6):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = lk.calculate_L(r, x); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"GaussianApproxLKernel.calculate_L N=2^{lg} D={D}: min {min(ts[1:]):.3f} ms; checksum {float(out.sum()):.10e}")
