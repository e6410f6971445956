"""Host <-> device copy rates of this box with pinned buffers (the bound of the e2e path): one direction alone, both at once."""
import time
import torch
n = 1 << 23   # 64 MB of float64
h_in, h_out = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.float64, device="cuda"), torch.randn(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


for name, fn in (("H2D 64 MB", h2d), ("D2H 64 MB", d2h), ("H2D + D2H 64 MB each, concurrently", both)):
    ms = timed(fn)
    print(f"{name}: {ms:.3f} ms -> {64 * 1.048576 / ms:.1f} GB/s per direction")
