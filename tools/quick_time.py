"""Ad-hoc timing of the hot kernels on the GPU box (development aid; bench.py is the contract)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "smc-nuts_b200"))
from smcnuts import _cabi, _device as dev  # noqa: E402
from smcnuts.distributions import StdNormal  # noqa: E402
from smcnuts.model.device_model import make_model  # noqa: E402
from smcnuts.proposal.nuts import NUTSProposal  # noqa: E402
from smcnuts.smc_sampler import SMCSampler  # noqa: E402


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


def probe():
    sink = dev.zeros(1)
    blocks, threads, iters = 148 * 8, 256, 20000
    t, _ = ev_time(lambda: _cabi.call("smcb_probe_fp64", blocks, threads, iters, dev.ptr(sink), dev.stream_ptr()))
    fl = blocks * threads * iters * 8 * 2
    print(f"FP64 DFMA probe: {fl / t / 1e12:.2f} TFLOP/s ({t * 1e3:.2f} ms)")
    for warps in (1, 2, 4, 8):
        t2, _ = ev_time(lambda: _cabi.call("smcb_probe_dmma", 148 * 4, 32 * warps, iters, dev.ptr(sink), dev.stream_ptr()))
        fl2 = 148 * 4 * warps * iters * 8 * 512
        print(f"FP64 DMMA probe ({warps} warps/CTA x 4 CTA/SM): {fl2 / t2 / 1e12:.2f} TFLOP/s ({t2 * 1e3:.2f} ms)")
    return fl / t


def nuts(name, N, eps, flop_per_eval, spread, centre, iters=3, carry_mode=False):
    m = make_model(name) if name != "gauss" else make_model("gauss", dim=100)
    D = m.dim
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g) * spread + torch.tensor(centre, dtype=torch.float64, device="cuda")
    k = NUTSProposal(m, StdNormal(D), eps, rng=10)
    carry = None
    for it in range(iters):
        r = StdNormal(D, seed=10).rvs(N, iteration=it)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = k.transition(x, r, 1.0, iteration=it, carry=carry, want_grad=carry_mode); b.record(); torch.cuda.synchronize()
        if carry_mode:
            carry = (o["A_new"], o["B_new"], o["g_new"])
        t = a.elapsed_time(b) * 1e-3
        nl = int(o["n_leapfrog"].sum().item())
        mx = int(o["n_leapfrog"].max().item())
        print(f"{name} N={N}: {t * 1e3:.2f} ms, {nl} leapfrogs (mean {nl / N:.1f}, max {mx}) -> {nl / t / 1e9:.3f} G grad-evals/s, "
              f"{(nl + N) * flop_per_eval / t / 1e12:.2f} TFLOP/s")
        x = o["x_new"]


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0))
    if which in ("all", "probe"):
        probe()
    if which in ("all", "arma"):
        nuts("arma", 1 << 20, 0.01, 3900, 0.02, [0.0068, 0.957, -0.034, float(np.log(0.1666))], iters=4)
    if which == "logp":
        for name, flop, D in (("arma", 3900, 4), ("PRMwCD", 5200, 13)):
            m = make_model(name)
            x = torch.randn(1 << 22, D, dtype=torch.float64, device="cuda") * 0.1
            A, B, g = dev.empty(1 << 22), dev.empty(1 << 22), dev.empty(1 << 22, D)
            for N in (1 << 20, 1 << 22):
                t, _ = ev_time(lambda: _cabi.call("smcb_logp_grad", m.handle, dev.ptr(x), N, 1.0, dev.ptr(A), dev.ptr(B), dev.ptr(g),
                                                  dev.stream_ptr()))
                print(f"logp_grad {name} N={N}: {t * 1e6:.1f} us -> {N * flop / t / 1e12:.2f} TFLOP/s")
    if which == "armacarry":
        nuts("arma", 1 << 20, 0.01, 3900, 0.02, [0.0068, 0.957, -0.034, float(np.log(0.1666))], iters=5, carry_mode=True)
    if which == "armaN":
        for lg in (18, 19, 20, 21, 22, 23):
            nuts("arma", 1 << lg, 0.01, 3900, 0.02, [0.0068, 0.957, -0.034, float(np.log(0.1666))], iters=3)
    if which.startswith("prmN"):   # prmN17,20 -> PRMwCD at the listed log2 sizes
        for lg in which[4:].split(","):
            nuts("PRMwCD", 1 << int(lg), 0.01, 5200, 0.02, [0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521,
                                                            -0.3014, 1.6721, -0.1868, -0.1491, float(np.log(0.3326))], iters=2)
    if which == "prm20":
        nuts("PRMwCD", 1 << 20, 0.01, 5200, 0.02, [0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014,
                                                 1.6721, -0.1868, -0.1491, float(np.log(0.3326))], iters=2)
    if which in ("all", "prm"):
        nuts("PRMwCD", 1 << 18, 0.01, 5200, 0.02, [0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014,
                                                 1.6721, -0.1868, -0.1491, float(np.log(0.3326))], iters=2)
    if which in ("all", "gauss"):
        nuts("gauss", 1 << 18, 0.1, 20200, 1.0, [0.0] * 100, iters=3)
    if which in ("all", "smc"):
        m = make_model("arma")
        t0 = time.time()
        s = SMCSampler(K=10, N=1 << 20, target=m, step_size=0.01, sample_proposal=StdNormal(4), momentum_proposal=StdNormal(4),
                       lkernel="forwardsLKernel", tempering=False, rng=10)
        s.sample(show_progress=False)
        print(f"SMC arma N=2^20 K=10: run_time {s.run_time:.3f}s (ctor+run {time.time() - t0:.3f}s), leapfrogs {s.leapfrogs}, "
              f"propose ms {np.round(s.propose_time * 1e3, 2)}")
        print("mean", s.mean_estimate[-1], "ess", s.ess)
