"""CPU: host-side logic of the plugin layer (batched bisection, shard routing, seeds, collectives over gloo)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from scipy.optimize import bisect as scipy_bisect

from oracle import smc_oracle as O
from smcnuts import _device as dev
from smcnuts.parallel import ShardContext, split_counts, systematic_slot_bounds
from smcnuts.tempering.adaptive_tempering import _merge_triples, bisect_batched

ROOT = Path(__file__).resolve().parents[1]


def test_batched_bisect_reproduces_scipy_bit_for_bit():
    rng = np.random.default_rng(0)
    for _ in range(60):
        a, s = rng.uniform(0.02, 0.98), rng.uniform(0.5, 30)
        lo = rng.uniform(0.0, a * 0.9)
        f = lambda p: -(np.tanh(s * (p - a)) + 0.01 * (p - a))   # noqa: E731  positive at lo, negative at 1
        calls = []

        def fb(ps):
            calls.append(len(ps))
            return [f(p) for p in ps]
        got = bisect_batched(fb, lo, 1.0)
        assert got == scipy_bisect(f, lo, 1.0) == O.bisect(f, lo, 1.0)
        assert len(calls) <= 12 and max(calls) <= 16


def test_batched_bisect_edge_cases():
    assert bisect_batched(lambda ps: [1.0 for _ in ps], 0.2, 1.0) == 1.0           # f(1) >= 0 -> 1.0
    with pytest.raises(ValueError):
        bisect_batched(lambda ps: [-1.0 for _ in ps], 0.2, 1.0)                     # same sign
    with pytest.raises(ValueError):
        bisect_batched(lambda ps: [float("nan") if p < 1 else -1.0 for p in ps], 0.2, 1.0)


def test_device_bisection_walk_reproduces_scipy_bit_for_bit(golden):
    """csrc/bisect.cuh (the walk the GPU runs, compiled here with g++) against scipy.optimize.bisect: same roots, same
    iteration counts, same error cases, and the reference's tempering goldens -- within the fixed schedule of passes."""
    from tests.hostsim import sim
    rng = np.random.default_rng(0)
    for _ in range(80):
        a, sc = rng.uniform(0.02, 0.98), rng.uniform(0.5, 30)
        lo = rng.uniform(0.0, a * 0.9)
        f = lambda p: -(np.tanh(sc * (p - a)) + 0.01 * (p - a))   # noqa: E731  positive at lo, negative at 1
        root, status, iters, _, passes = sim.bisect(f, lo, 1.0)
        ref, info = scipy_bisect(f, lo, 1.0, full_output=True)
        assert status == 1 and root == ref and iters == info.iterations and passes <= 11
    # worst case for the schedule: the full interval and a root that is never hit exactly
    root, status, iters, _, passes = sim.bisect(lambda p: 0.3 - p + 1e-13, 0.0, 1.0)
    assert status == 1 and root == scipy_bisect(lambda p: 0.3 - p + 1e-13, 0.0, 1.0) and passes <= 11
    assert sim.bisect(lambda p: np.ones_like(p), 0.2, 1.0)[:2] == (1.0, 1)                 # f(1) >= 0 -> phi = 1
    assert sim.bisect(lambda p: -np.ones_like(p), 0.2, 1.0)[1] == 3                         # same sign -> ValueError
    assert sim.bisect(lambda p: np.where(p < 1, np.nan, -1.0), 0.2, 1.0)[1] == 2           # NaN -> ValueError
    assert sim.bisect(lambda p: np.where(p == 0.2, 0.0, -1.0), 0.2, 1.0)[:2] == (0.2, 1)   # f(a) == 0 -> a
    g = golden("tempering")
    for j in range(4):
        ll, lpri, c = g[f"temper_{j}_loglik"], g[f"temper_{j}_logpri"], g[f"temper_{j}_lp_old"]
        N = len(ll)
        fb = lambda ps: np.array([O.ess_at_phi(p, ll, lpri, c) - N * 0.5 for p in ps])   # noqa: E731
        root, status, _, _, _ = sim.bisect(fb, float(g[f"temper_{j}_old_phi"]), 1.0)
        assert status == 1 and root == float(g[f"temper_{j}_phi"])


def test_tempering_golden_through_batched_bisect(golden):
    g = golden("tempering")
    for j in range(4):
        ll, lpri, c = g[f"temper_{j}_loglik"], g[f"temper_{j}_logpri"], g[f"temper_{j}_lp_old"]
        N = len(ll)
        fb = lambda ps: [O.ess_at_phi(p, ll, lpri, c) - N * 0.5 for p in ps]   # noqa: E731
        assert bisect_batched(fb, float(g[f"temper_{j}_old_phi"]), 1.0) == float(g[f"temper_{j}_phi"])


def test_merge_triples_matches_direct_logsumexp():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(4, 1000)) * 30
    x[2, :] = -np.inf
    tri = []
    for p in range(4):
        fin = x[p][~np.isneginf(x[p])]
        m = fin.max() if len(fin) else -np.inf
        tri.append([[m, np.exp(fin - m).sum() if len(fin) else 0.0, np.exp(2 * (fin - m)).sum() if len(fin) else 0.0]])
    m = _merge_triples(np.array(tri))[0]
    wn, logZ = O.normalise_weights(x.ravel())
    assert np.isclose(m[0] + np.log(m[1]), logZ, rtol=1e-14)
    assert np.isclose(m[1] ** 2 / m[2], O.calculate_ess(wn), rtol=1e-12)


def test_systematic_slot_bounds_match_per_slot_search():
    rng = np.random.default_rng(2)
    for N, P in ((64, 2), (1000, 4), (4096, 8)):
        w = rng.exponential(size=N) ** 2
        w[rng.integers(0, N, N // 4)] = 0
        cdf = O.cdf_of(w / w.sum())
        u0 = rng.uniform()
        idx = O.systematic_ancestors(None, u0, cdf=cdf)
        m = N // P
        bounds = systematic_slot_bounds(np.concatenate([[0.0], cdf[m - 1::m]]), u0, N)
        assert bounds[0] == 0 and bounds[-1] == N and np.all(np.diff(bounds) >= 0)
        for q in range(P):   # rank q serves exactly the slots whose ancestor it owns
            served = np.arange(bounds[q], bounds[q + 1])
            assert np.all(idx[served] // m == q)
            sc = split_counts(int(bounds[q]), int(bounds[q + 1]), m, P)
            assert sum(sc) == len(served)
            assert sc == [int(np.sum(served // m == d)) for d in range(P)]


def test_seed_from_rng_is_reproducible():
    assert dev.seed_from_rng(10) == 10
    assert dev.seed_from_rng(np.random.RandomState(7)) == dev.seed_from_rng(np.random.RandomState(7))
    assert dev.seed_from_rng(np.random.default_rng(7)) == dev.seed_from_rng(np.random.default_rng(7))
    assert dev.seed_from_rng(None) == 0


def test_std_normal_detection():
    from scipy.stats import multivariate_normal
    assert dev.is_std_normal(multivariate_normal(mean=np.zeros(4), cov=np.eye(4)), 4)
    assert not dev.is_std_normal(multivariate_normal(mean=np.zeros(4), cov=2 * np.eye(4)), 4)
    assert not dev.is_std_normal(multivariate_normal(mean=np.ones(4), cov=np.eye(4)), 4)


def test_single_shard_context_is_a_noop():
    import torch
    s = ShardContext()
    assert (s.world, s.rank) == (1, 0) and s.local_count(12) == 12 and s.offset(12) == 0
    t = torch.arange(3.0)
    assert s.all_gather_vec(t).shape == (1, 3) and s.all_reduce_sum_(t) is t


def test_gloo_world2_collectives_and_systematic_routing(tmp_path):
    """world_size-2 gloo run of the N>1 host path: shard partition, all_gather_vec, all_reduce, the systematic
    slot split and the all-to-all-v row migration (rows gathered with numpy in place of the CUDA gather)."""
    script = tmp_path / "w2.py"
    script.write_text(f'''
import sys, os
sys.path.insert(0, {str(ROOT)!r}); sys.path.insert(0, {str(ROOT / "smc-nuts_b200")!r})
import numpy as np, torch, torch.distributed as dist
from oracle import smc_oracle as O
from smcnuts.parallel import ShardContext, split_counts, systematic_slot_bounds
dist.init_process_group("gloo")
s = ShardContext()
assert s.world == 2
N, D = 512, 3
rng = np.random.default_rng(4)
w = rng.exponential(size=N) ** 3; wn = w / w.sum(); x = rng.normal(size=(N, D)); u0 = 0.37
m = s.local_count(N); lo_p = s.offset(N)
# global exclusive scan of rank totals, as Resampler._cdf does
tot = torch.tensor([wn[lo_p:lo_p + m].sum()])
totals = s.all_gather_vec(tot).view(-1).numpy()
off = np.concatenate([[0.0], np.cumsum(totals)])
cdf_local = (off[s.rank] + np.cumsum(wn[lo_p:lo_p + m])) / off[-1]
lasts = s.all_gather_vec(torch.tensor([cdf_local[-1]])).view(-1).numpy()
bounds = systematic_slot_bounds(np.concatenate([[0.0], lasts]), u0, N)
lo, hi = int(bounds[s.rank]), int(bounds[s.rank + 1])
pos = (np.arange(lo, hi, dtype=np.float64) + u0) / N
idx_local = np.minimum(np.searchsorted(cdf_local, pos, side="right"), m - 1)
send = torch.tensor(x[lo_p:lo_p + m][idx_local]).reshape(-1, D)
send_counts = split_counts(lo, hi, m, 2)
mylo, myhi = s.rank * m, (s.rank + 1) * m
recv_counts = [max(0, min(myhi, int(bounds[q + 1])) - max(mylo, int(bounds[q]))) for q in range(2)]
got = s.all_to_all_rows(send, send_counts, recv_counts).numpy()
cdf = np.concatenate([c.numpy() for c in [torch.tensor(cdf_local)]])
full_cdf = s.all_gather_vec(torch.tensor(cdf_local)).view(-1).numpy()
want = x[np.minimum(np.searchsorted(full_cdf, (np.arange(N) + u0) / N, side="right"), N - 1)][mylo:myhi]
assert got.shape == want.shape and np.array_equal(got, want), (s.rank, got.shape)
t = torch.ones(4, dtype=torch.float64) * (s.rank + 1)
assert torch.equal(s.all_reduce_sum_(t), torch.full((4,), 3.0, dtype=torch.float64))
assert s.all_reduce_sum_scalar(s.rank + 1) == 3.0
dist.destroy_process_group()
print("rank", s.rank, "ok")
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def _mma_m8n8k4(c0, c1, a, b):
    """numpy emulation of mma.sync.m8n8k4.f64 on per-lane fragments (arrays of 32 lanes): A[row=l/4][k=l%4],
    B[k=l%4][n=l/4], C[row=l/4][cols 2(l%4), 2(l%4)+1]."""
    lanes = np.arange(32)
    Am = np.zeros((8, 4)); Bm = np.zeros((4, 8))
    Am[lanes // 4, lanes % 4] = a
    Bm[lanes % 4, lanes // 4] = b
    Cm = Am @ Bm
    return c0 + Cm[lanes // 4, 2 * (lanes % 4)], c1 + Cm[lanes // 4, 2 * (lanes % 4) + 1]


def test_prmwcd_tensor_core_fragment_packing_reproduces_the_model():
    """The PRMwCD NUTS kernel evaluates the model as two FP64 tensor-core products on pre-packed fragments
    (csrc/models.cuh, PrmModelG).  Emulate its exact dataflow on the CPU from the packing the library produces and
    compare with the oracle's density: catches any fragment-layout error without a GPU."""
    import ctypes
    import json
    import math
    from smcnuts import _cabi
    from smcnuts.model.device_model import DATA_DIR
    data = json.loads((DATA_DIR / "PRMwCD" / "PRMwCD.json").read_text())
    NO, NT = int(data["N"]), 13
    y = np.asarray(data["y"], dtype=np.float64)
    X = np.asarray(data["Xkernel"], dtype=np.float64).reshape(NO, 11)
    scalar = np.zeros(16 + NO * 12)
    scalar[0] = y.sum(); scalar[1:12] = X.T @ y; scalar[12] = sum(math.lgamma(v + 1.0) for v in y)
    rows = scalar[16:].reshape(NO, 12); rows[:, :11] = X; rows[:, 11] = y
    out = np.zeros(32 + NT * 288)
    rc = _cabi.lib().smcb_debug_pack_prm(scalar.ctypes.data_as(ctypes.c_void_p), NO, NT,
                                         out.ctypes.data_as(ctypes.c_void_p), out.size)
    assert rc == 0
    PF1, PF2 = 32, 32 + NT * 96
    YM = PF2 + NT * 128
    t = O.COracleTarget("PRMwCD")
    rng = np.random.default_rng(5)
    xs = rng.normal(size=(8, 13)) * 0.4
    xs[:, 12] = rng.normal(size=8) * 0.3 - 1.0
    lanes = np.arange(32)
    grp, sub = lanes // 4, lanes % 4
    xl = np.zeros((4, 32))                       # local slot i of every lane: coordinate sub + 4 i of particle grp
    for i in range(4):
        j = sub + 4 * i
        xl[i] = np.where(j < 13, xs[grp, np.minimum(j, 12)], 0.0)
    e = np.zeros((2 * NT, 32))
    for kk in range(3):
        for nt in range(NT):
            e[2 * nt], e[2 * nt + 1] = _mma_m8n8k4(e[2 * nt], e[2 * nt + 1], xl[kk], out[PF1 + (nt * 3 + kk) * 32:][:32])
    # eta of particle g, observation 8 nt + 2 t + h sits in e[2 nt + h] of lane (g, t)
    eta_ref = xs[:, :1] + xs[:, 1:12] @ X.T
    for nt in range(NT):
        for h in range(2):
            o = 8 * nt + 2 * sub + h
            ok = o < NO
            np.testing.assert_allclose(e[2 * nt + h][ok], eta_ref[grp[ok], o[ok]], rtol=1e-13, atol=1e-13)
            assert np.all(e[2 * nt + h][~ok] == 0.0)
            assert np.array_equal(out[YM + (nt * 2 + h) * 32:][:32], np.where(ok, y[np.minimum(o, NO - 1)] > 0, False).astype(float))
    lam = np.exp(e)
    c = np.zeros((4, 32))
    for nt in range(NT):
        for h in range(2):
            for nt2 in range(2):
                c[2 * nt2], c[2 * nt2 + 1] = _mma_m8n8k4(c[2 * nt2], c[2 * nt2 + 1], lam[2 * nt + h],
                                                        out[PF2 + ((nt * 2 + h) * 2 + nt2) * 32:][:32])
    # assemble A, B and the gradient exactly like the device epilogue and compare with the oracle at phi = 0.7
    phi, q = 0.7, float(data["q"])
    gg = xs[:, 12]
    ig = np.exp(-gg)
    grad = np.zeros((8, 13)); ydot = np.zeros(8); ssum = np.zeros(8)
    for i in range(3):
        for l in range(32):
            j, g_ = sub[l] + 4 * i, grp[l]
            hy = out[j]
            ydot[g_] += xl[i][l] * hy
            if j == 0:
                grad[g_, j] = phi * (hy - c[i][l])
            else:
                aq = math.sqrt(abs(xl[i][l]) * ig[g_])
                ssum[g_] += aq
                grad[g_, j] = -q * aq / xl[i][l] + phi * (hy - c[i][l])
    slam = c[0][lanes[sub == 0]]
    grad[:, 12] = -3.0 + 1.3 * ig + 1.0 - 11 + q * ssum
    Bv = ydot - slam - out[16]
    Av = (2.0 * 0.26236426446749105204 - 3.0 * gg - 1.3 * ig) + gg + (-11 * gg - ssum)
    Ao, Bo, _, _ = t.split(xs, grads=False)
    np.testing.assert_allclose(Av, Ao, rtol=1e-12)
    np.testing.assert_allclose(Bv, Bo, rtol=1e-12)
    np.testing.assert_allclose(grad, t.logpdfgrad(xs, phi), rtol=1e-11, atol=1e-11)


def test_pipeline_chunk_bounds_cover_every_row_once():
    """Host-side chunking of NUTSProposal.rvs (copy/compute pipeline): contiguous, ordered, complete for any N."""
    from smcnuts.proposal.nuts import NUTSProposal, chunk_bounds
    for n in (0, 1, 3, 7, 8, 1000, (1 << 17) + 1234, 1 << 20):
        for fr in (NUTSProposal.PIPELINE_FRACTIONS, (1,), (1, 1, 1, 1), (1, 7, 7, 1), (5, 3)):
            b = chunk_bounds(n, fr)
            assert b[0] == 0 and b[-1] == n and len(b) == len(fr) + 1
            assert all(lo <= hi for lo, hi in zip(b, b[1:]))
            if n >= 64:
                sizes = np.diff(b) / n
                np.testing.assert_allclose(sizes, np.array(fr) / sum(fr), atol=2.0 / n)


def _parse_table(text, name):
    import re
    body = re.search(name + r"\[\d+\] = \{(.*?)\};", text, re.S).group(1)
    return [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-fA-F.]+p[+-]?\d+", body)]


def test_fast_exp_and_fast_log_tables_and_algorithms():
    """csrc/common.cuh::fast_exp / fast_log restated in Python with exactly rounded FMAs, using the tables and constants
    parsed from the header: the committed tables are the correctly rounded 2^(j/32), 1/c_j and -log(1/c_j), and the
    algorithms meet the accuracy the device tests assert (exp <= 1 ulp, log <= 2.5e-16 * max(1, |log u|))."""
    import math
    import re
    import struct
    from fractions import Fraction as F
    import mpmath as mp
    mp.mp.dps = 50
    text = (ROOT / "smc-nuts_b200" / "csrc" / "common.cuh").read_text()
    T, INV, LOGC = _parse_table(text, "kExpT"), _parse_table(text, "kLogInvC"), _parse_table(text, "kLogC")
    assert len(T) == 32 and len(INV) == 64 and len(LOGC) == 64
    assert T == [float(mp.mpf(2) ** (mp.mpf(j) / 32)) for j in range(32)]
    assert INV == [float(1 / (1 + (mp.mpf(j) + mp.mpf(1) / 2) / 64)) for j in range(64)]
    assert LOGC == [float(-mp.log(mp.mpf(v))) for v in INV]
    const = {k: float.fromhex(v) for k, v in re.findall(r"constexpr double (kExp\w+) = (0x[0-9a-fp.+-]+);", text)}
    inv_l, l_hi, l_lo = const["kExpInvL"], const["kExpLHi"], const["kExpLLo"]
    assert inv_l == float(32 / mp.log(2)) and abs(float(mp.mpf(l_hi) + mp.mpf(l_lo) - mp.log(2) / 32)) < 1e-28
    magic = 6755399441055744.0

    def fma(a, b, c):
        return float(F(a) * F(b) + F(c))

    def fexp(x):
        t = fma(x, inv_l, magic)
        kp = struct.unpack("<i", struct.pack("<d", t)[:4])[0]
        t -= magic
        r = fma(t, -l_lo, fma(t, -l_hi, x))
        s = r * r
        b1 = fma(1 / 720, s, fma(1 / 120, r, 1 / 24))
        q = fma(s, fma(s, b1, fma(1 / 6, r, 0.5)), r)
        return math.ldexp(fma(T[kp & 31], q, T[kp & 31]), kp >> 5)

    ln2_hi, ln2_lo = (float.fromhex(v) for v in re.search(
        r"fma\(e, (0x[0-9a-fp.+-]+), SMCB_LDG\(&kLogC\[j\]\)\) \+ fma\(e, (0x[0-9a-fp.+-]+), l1\)", text).groups())

    def flog(u):
        bits = struct.unpack("<q", struct.pack("<d", u))[0]
        e, j = ((bits >> 52) & 0x7ff) - 1023, (bits >> 46) & 63
        m = struct.unpack("<d", struct.pack("<q", (bits & ((1 << 52) - 1)) | (1023 << 52)))[0]
        r = fma(m, INV[j], -1.0)
        p = fma(1 / 7, r, -1 / 6)
        for c in (0.2, -0.25, 1 / 3, -0.5):
            p = fma(p, r, c)
        return fma(float(e), ln2_hi, LOGC[j]) + fma(float(e), ln2_lo, fma(r * r, p, r))

    rng = np.random.default_rng(0)
    # the oracle's C restatement (oracle/devmath.h, independent tables from mpmath) is the same function, bit for bit
    from oracle import smc_oracle as O
    xs = np.concatenate([rng.uniform(-707, 707, 300), rng.normal(size=300) * 3])
    assert np.array_equal(O.devmath_exp(xs), np.array([fexp(float(v)) for v in xs]))
    us = np.concatenate([1 + np.exp(rng.uniform(-30, 30, 300)), 1 - rng.random(300), np.exp(rng.uniform(-700, 700, 200))])
    assert np.array_equal(O.devmath_log(us), np.array([flog(float(v)) for v in us]))
    assert np.array_equal(O.devmath_exp(np.array([710.0, -746.0, np.inf, -np.inf, 709.5, -740.0])),
                          np.array([np.inf, 0.0, np.inf, 0.0, float(mp.exp(mp.mpf(709.5))), O.devmath_exp(np.array([-740.0]))[0]]))
    assert np.isnan(O.devmath_exp(np.array([np.nan]))[0])
    worst = 0.0
    for x in np.concatenate([rng.uniform(-707, 707, 700), rng.normal(size=700) * 3]):
        g, ref = fexp(float(x)), mp.exp(mp.mpf(float(x)))
        worst = max(worst, float(abs(mp.mpf(g) - ref) / mp.mpf(float(np.spacing(g)))))
    assert worst <= 1.0, worst
    worst = 0.0
    for u in np.concatenate([1 + np.exp(rng.uniform(-30, 30, 600)), 1 - rng.random(600), np.exp(rng.uniform(-700, 700, 300))]):
        ref = mp.log(mp.mpf(float(u)))
        worst = max(worst, float(abs(mp.mpf(flog(float(u))) - ref)) / max(1.0, abs(float(ref))))
    assert worst <= 2.5e-16, worst


def test_mse_mean_var_matches_the_reference_function(golden):
    """experiments/run_experiments.py::mse_mean_var against outputs of the unmodified plot_experiments.py:60-78."""
    sys.path.insert(0, str(ROOT / "experiments"))
    from run_experiments import mse_mean_var, mse_per_iteration
    g = golden("mse")
    mean, var = mse_mean_var(g["x"], g["truth"])
    np.testing.assert_allclose(mean, g["mse_mean"], rtol=1e-13)
    np.testing.assert_allclose(var, g["mse_var"], rtol=1e-12)
    np.testing.assert_allclose(mse_per_iteration(g["x"], g["truth"]), g["mse_mean"], rtol=1e-13)
