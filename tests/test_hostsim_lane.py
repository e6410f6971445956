"""CPU: the DEVICE lane state machine (csrc/nuts_lane.cuh compiled with g++, tests/hostsim) against the
oracle and the reference golden transitions.  Validates kernel logic without a GPU."""
import numpy as np
import pytest

from oracle import smc_oracle as O
from tests.hostsim import sim

CASES = ["arma", "arma_tempered", "arma_prior", "PRMwCD", "PRMwCD_tempered", "gauss8", "gauss100"]


def _target(case):
    tname = case.split("_")[0]
    kw = {}
    if tname.startswith("gauss"):
        kw, tname = {"dim": int(tname[5:])}, "gauss"
    return tname, O.COracleTarget(tname, **kw)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("accrej", [False, True])
def test_lane_matches_reference_golden(golden, case, accrej):
    g = golden("nuts")
    tname, t = _target(case)
    x0, r0 = g[f"{case}_x0"], g[f"{case}_r0"]
    eps, phi, it, seed = float(g[f"{case}_eps"]), float(g[f"{case}_phi"]), int(g[f"{case}_iteration"]), int(g[f"{case}_seed"])
    o = sim.nuts(tname, t.np_target, x0, r0, eps, phi, 10, accrej, seed, it, 0, lanes=5)
    assert np.array_equal(o["n_leapfrog"], g[f"{case}_n_leapfrog"])
    acc = g[f"{case}_accepted"] if accrej else np.ones(len(x0), dtype=bool)
    assert np.array_equal(o["accepted"].astype(bool), acc)
    np.testing.assert_allclose(o["x_new"][acc], g[f"{case}_x_new"][acc], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(o["r_new"][acc], g[f"{case}_r_new"][acc], rtol=1e-9, atol=1e-9)
    assert np.array_equal(o["x_new"][~acc], x0[~acc]) and np.array_equal(o["r_new"][~acc], r0[~acc])


@pytest.mark.parametrize("tname,kw,eps,N", [("arma", {}, 0.01, 400), ("PRMwCD", {}, 0.01, 40),
                                            ("gauss", {"dim": 8}, 0.1, 300), ("gauss", {"dim": 33}, 0.15, 60)])
@pytest.mark.parametrize("lanes", [1, 32])
@pytest.mark.parametrize("devmath", [False, True])
def test_lane_matches_c_oracle_bitwise(tname, kw, eps, N, lanes, devmath):
    """Same compiler flags, same expression order -> the lane machine must reproduce the recursive C
    oracle bit for bit (x', r', split log-densities, tree sizes, depths, MH outcomes).  devmath: both sides evaluate
    exp / log with the kernels' table-driven algorithms (csrc/common.cuh with -DSMCB_DEVMATH vs oracle/devmath.h) --
    the CPU twin of the parity device build that tests/test_gpu_parity_build.py checks on the B200."""
    t = O.COracleTarget(tname, **kw)
    rng = np.random.default_rng(5)
    x = rng.normal(size=(N, t.dim)) * 0.3
    if tname == "arma":
        x += np.array([0.0, 0.9, 0.0, -1.7])
        x[:20] = rng.normal(size=(20, 4)) * 2.0   # wild starts: divergences, -inf
        x[0, 3] = 800.0
    r = rng.normal(size=(N, t.dim))
    for accrej in (False, True):
        for phi in (1.0, 0.2):
            with O.devmath(devmath):
                ref = t.nuts_batch(x, r, eps, phi, 10, seed=77, iteration=3, particle0=1000, accrej=accrej)
            o = sim.nuts(tname, t.np_target, x, r, eps, phi, 10, accrej, 77, 3, 1000, lanes=lanes, devmath=devmath)
            assert np.array_equal(o["n_leapfrog"], ref["n_leapfrog"])
            assert np.array_equal(o["depth"], ref["depth"])
            assert np.array_equal(o["accepted"], ref["accepted"])
            assert np.array_equal(o["x_new"], ref["x_new"], equal_nan=True)
            assert np.array_equal(o["r_new"], ref["r_new"], equal_nan=True)
            with np.errstate(invalid="ignore"):
                lp_old = o["A_old"] + phi * o["B_old"]
                lp_new = o["A_new"] + phi * o["B_new"]
            lp_old = np.where(np.isfinite(lp_old), lp_old, -np.inf)
            lp_new = np.where(np.isfinite(lp_new), lp_new, -np.inf)
            assert np.array_equal(lp_old, ref["lp_old"]) and np.array_equal(lp_new, ref["lp_new"])
            np.testing.assert_allclose(o["ke_old"], 0.5 * np.sum(r * r, axis=1), rtol=1e-14)


def test_max_depth_cap():
    """depth > MAX_TREE_DEPTH break (nuts.py:109-110): a flat target never U-turns -> 2^(L+1)-1 leapfrogs."""
    t = O.COracleTarget("gauss", dim=4)
    x = np.zeros((3, 4)); r = np.ones((3, 4)) * 1e-3
    for L in (3, 10):
        ref = t.nuts_batch(x, r, 1e-6, 1.0, L, seed=1)
        o = sim.nuts("gauss", t.np_target, x, r, 1e-6, 1.0, L, False, 1, 0, 0, lanes=2)
        assert np.all(ref["n_leapfrog"] == 2 ** (L + 1) - 1) and np.array_equal(o["n_leapfrog"], ref["n_leapfrog"])
        assert np.array_equal(o["x_new"], ref["x_new"])


@pytest.mark.parametrize("tname,kw,eps", [("arma", {}, 0.01), ("gauss", {"dim": 8}, 0.1)])
def test_gradient_carry_over_skips_the_initial_evaluation_without_changing_anything(tname, kw, eps):
    """Two consecutive transitions: handing (A_new, B_new, g_new) of the first to the second must give bit-identical
    results to re-evaluating the model at the start point (nuts.py:66,72)."""
    t = O.COracleTarget(tname, **kw)
    rng = np.random.default_rng(9)
    N = 200
    x = rng.normal(size=(N, t.dim)) * 0.2 + (np.array([0.0, 0.9, 0.0, -1.7]) if tname == "arma" else 0.0)
    r1, r2 = rng.normal(size=(2, N, t.dim))
    a = sim.nuts(tname, t.np_target, x, r1, eps, 1.0, 10, False, 5, 0, 0, lanes=7, want_grad=True)
    np.testing.assert_allclose(a["g_new"], t.logpdfgrad(a["x_new"], 1.0), rtol=1e-12, atol=1e-12)
    plain = sim.nuts(tname, t.np_target, a["x_new"], r2, eps, 1.0, 10, False, 5, 1, 0, lanes=7, want_grad=True)
    carried = sim.nuts(tname, t.np_target, a["x_new"], r2, eps, 1.0, 10, False, 5, 1, 0, lanes=7,
                       carry=(a["A_new"], a["B_new"], a["g_new"]), want_grad=True)
    for k in ("x_new", "r_new", "A_old", "B_old", "A_new", "B_new", "n_leapfrog", "depth", "g_new", "ke_old", "ke_new"):
        assert np.array_equal(plain[k], carried[k]), k


# ------------------------------------------------------------------------------------------------ group kernels
# The 4-lanes-per-particle kernels (GaussModelG, PrmModelG: FP64 tensor-core products, group folds by shuffle, the
# warp-level work queue) run on the CPU through the fiber-based warp emulator tests/hostsim/simt_emu.h.

@pytest.mark.parametrize("dim,eps,N", [(8, 0.1, 150), (33, 0.15, 50), (100, 0.1, 20)])
def test_group_kernel_gauss_matches_c_oracle_bitwise(dim, eps, N):
    """The emulated mma accumulates k = 0..3 in order, so -Px is summed in the oracle's order: trees, MH outcomes and the
    returned states of the tensor-core group kernel must equal the recursive C oracle bit for bit (ragged work queue: N
    is not a multiple of the 8 particles a warp holds)."""
    t = O.COracleTarget("gauss", dim=dim)
    rng = np.random.default_rng(5)
    x = rng.normal(size=(N, dim)) * 0.3
    r = rng.normal(size=(N, dim))
    for accrej, phi in ((False, 1.0), (True, 0.7)):
        ref = t.nuts_batch(x, r, eps, phi, 10, seed=77, iteration=3, particle0=1000, accrej=accrej)
        o = sim.nuts_simt("gauss", t.np_target, x, r, eps, phi, 10, accrej, 77, 3, 1000)
        for k in ("n_leapfrog", "depth", "accepted", "x_new", "r_new"):
            assert np.array_equal(o[k], ref[k]), k
        np.testing.assert_allclose(o["ke_old"], 0.5 * np.sum(r * r, axis=1), rtol=1e-14)
        # the quadratic form is folded over the 4 lanes of a group by shuffles (a tree), the oracle sums it in order
        np.testing.assert_allclose(o["A_new"] + phi * o["B_new"], ref["lp_new"], rtol=1e-13)


def test_group_kernel_prmwcd_matches_c_oracle():
    """PrmModelG: eta and the gradient are tensor-core products over tiles of 8 observations, so sums are associated
    differently from the oracle's observation loop: values agree to rounding (1e-12), and the 250-leapfrog trajectories
    amplify that to a few changed tree sizes (>= 90 % equal, as on the GPU)."""
    import math
    t = O.COracleTarget("PRMwCD")
    rng = np.random.default_rng(21)
    N = 44                                                   # ragged: 5.5 warps-full of particles
    centre = np.array([0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014, 1.6721, -0.1868,
                       -0.1491, math.log(0.3326)])
    x = centre + rng.normal(size=(N, 13)) * 0.05
    x[:6] = rng.normal(size=(6, 13)) * 3.0                   # wild starts
    x[6:9, 0] = -800.0                                       # lambda underflows to 0 with y > 0 -> logp = -inf
    r = rng.normal(size=(N, 13))
    phi = 0.6
    ref = t.nuts_batch(x, r, 0.01, phi, 10, seed=3, iteration=1, accrej=True)
    o = sim.nuts_simt("PRMwCD", t.np_target, x, r, 0.01, phi, 10, True, 3, 1, 0)
    Ao, Bo, _, _ = t.split(x, grads=False)
    assert np.all(np.isneginf(o["B_old"][6:9])) and np.array_equal(np.isfinite(o["B_old"]), np.isfinite(Bo))
    fin = np.isfinite(Bo)
    np.testing.assert_allclose(o["A_old"][fin], Ao[fin], rtol=1e-13)
    np.testing.assert_allclose(o["B_old"][fin], Bo[fin], rtol=1e-12)
    same = o["n_leapfrog"] == ref["n_leapfrog"]
    assert same.mean() >= 0.9, same.mean()
    assert np.array_equal(o["depth"][same], ref["depth"][same])
    ok = same & (o["accepted"] == ref["accepted"]) & np.isfinite(ref["x_new"]).all(axis=1)
    close = np.isclose(o["x_new"][ok], ref["x_new"][ok], rtol=1e-4, atol=1e-6).all(axis=1)
    assert close.mean() >= 0.85, close.mean()
    rej = o["accepted"] == 0
    assert np.array_equal(o["x_new"][rej], x[rej]) and np.array_equal(o["r_new"][rej], r[rej])


def test_group_kernel_max_depth_cap_and_slot_pressure():
    """A flat target never U-turns: the group kernel must run to the depth cap (2^(L+1) - 1 leapfrogs), which is also
    the case with the most live checkpoints, candidates and the deferred sample slot at once."""
    t = O.COracleTarget("gauss", dim=4)
    x = np.zeros((11, 4)); r = np.ones((11, 4)) * 1e-3          # 11 particles: one full and one ragged octet
    for L in (3, 10):
        ref = t.nuts_batch(x, r, 1e-6, 1.0, L, seed=1)
        o = sim.nuts_simt("gauss", t.np_target, x, r, 1e-6, 1.0, L, False, 1, 0, 0)
        assert np.all(ref["n_leapfrog"] == 2 ** (L + 1) - 1) and np.array_equal(o["n_leapfrog"], ref["n_leapfrog"])
        assert np.array_equal(o["x_new"], ref["x_new"]) and np.array_equal(o["r_new"], ref["r_new"])
        assert np.array_equal(o["depth"], ref["depth"])
