"""CPU: the oracle (oracle/smc_oracle.c + oracle/smc_oracle.py) against fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import philox
from oracle import smc_oracle as O
from oracle.models import make_target

RTOL = 1e-10


def test_philox_kat():
    for ctr, key, exp in philox.KAT:
        out = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in out] == list(exp)
        import ctypes
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        O.lib().orc_philox(c, k, o)
        assert list(o) == list(exp)


def test_c_streams_match_numpy_definition():
    p = np.arange(5, 37)
    for draw in (0, 1, 2, 7):
        assert np.array_equal(O.uniforms(10, 3, 2, 5, 32, draw), philox.uniform(10, 3, 2, p, draw))
    np.testing.assert_allclose(O.normals(10, 3, 1, 5, 32, 13), philox.normals(10, 3, 1, p, 13), rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("name,tname,kw", [("arma", "arma", {}), ("PRMwCD", "PRMwCD", {}),
                                           ("gauss", "gauss", {"dim": 8}), ("gauss100", "gauss", {"dim": 100})])
def test_c_models_match_numpy_and_mpmath(golden, name, tname, kw):
    g = golden("models")
    t = O.COracleTarget(tname, **kw)
    X = g[f"{name}_X"]
    A, B, gA, gB = t.split(X)
    ok = np.isfinite(g[f"{name}_A"]) & np.isfinite(g[f"{name}_B"])
    np.testing.assert_allclose(A[ok], g[f"{name}_A"][ok], rtol=1e-13)
    np.testing.assert_allclose(B[ok], g[f"{name}_B"][ok], rtol=1e-12)
    np.testing.assert_allclose(gA[ok], g[f"{name}_gA"][ok], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(gB[ok], g[f"{name}_gB"][ok], rtol=1e-11, atol=1e-9)
    for phi in (0.0, 0.37, 1.0):
        lp, gr = t.logpdf(X, phi), t.logpdfgrad(X, phi)
        ref_lp, ref_g = g[f"{name}_lp_{phi}"], g[f"{name}_grad_{phi}"]
        assert np.array_equal(np.isneginf(lp), np.isneginf(ref_lp))
        fin = np.isfinite(ref_lp)
        np.testing.assert_allclose(lp[fin], ref_lp[fin], rtol=1e-12)
        np.testing.assert_allclose(gr[fin], ref_g[fin], rtol=1e-10, atol=1e-9)
        assert np.all(np.isneginf(gr[~fin]))
    if f"{name}_mp_lp_0.37" in g:
        np.testing.assert_allclose(t.logpdf(X[:8], 0.37), g[f"{name}_mp_lp_0.37"], rtol=1e-13)


def test_tempering_identity():
    """logpdf(x,phi) = logpdf(x,0) + phi*(logpdf(x,1)-logpdf(x,0))  (adaptive_tempering.py:38-43)."""
    for name in ("arma", "PRMwCD"):
        t = O.COracleTarget(name)
        x = np.random.default_rng(0).normal(size=(64, t.dim)) * 0.3
        l0, l1 = t.logpdf(x, 0.0), t.logpdf(x, 1.0)
        np.testing.assert_allclose(t.logpdf(x, 0.3), l0 + 0.3 * (l1 - l0), rtol=1e-12)


NUTS_CASES = ["arma", "arma_tempered", "arma_prior", "PRMwCD", "PRMwCD_tempered", "gauss8", "gauss100"]


@pytest.mark.parametrize("case", NUTS_CASES)
def test_c_nuts_matches_reference_transitions(golden, case):
    """One NUTS transition per particle: same x0, r0, same Philox draws -> same (x', r'), same tree size."""
    g = golden("nuts")
    tname = case.split("_")[0]
    kw = {}
    if tname.startswith("gauss"):
        kw, tname = {"dim": int(tname[5:])}, "gauss"
    t = O.COracleTarget(tname, **kw)
    x0, r0 = g[f"{case}_x0"], g[f"{case}_r0"]
    out = t.nuts_batch(x0, r0, float(g[f"{case}_eps"]), float(g[f"{case}_phi"]), 10, int(g[f"{case}_seed"]),
                       int(g[f"{case}_iteration"]), 0, accrej=False)
    assert np.array_equal(out["n_leapfrog"], g[f"{case}_n_leapfrog"])
    np.testing.assert_allclose(out["x_new"], g[f"{case}_x_new"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(out["r_new"], g[f"{case}_r_new"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out["lp_new"], t.logpdf(out["x_new"], float(g[f"{case}_phi"])), rtol=1e-12)
    out2 = t.nuts_batch(x0, r0, float(g[f"{case}_eps"]), float(g[f"{case}_phi"]), 10, int(g[f"{case}_seed"]),
                        int(g[f"{case}_iteration"]), 0, accrej=True)
    acc = g[f"{case}_accepted"]
    assert np.array_equal(out2["accepted"].astype(bool), acc)
    np.testing.assert_allclose(out2["x_new"][acc], g[f"{case}_x_new"][acc], rtol=1e-9, atol=1e-11)
    assert np.array_equal(out2["x_new"][~acc], x0[~acc]) and np.array_equal(out2["r_new"][~acc], r0[~acc])


def test_choice_restatement(golden):
    g = golden("choice")
    for flavour in ("RandomState", "Generator"):
        for N in (1, 2, 7, 100, 4096):
            idx = O.multinomial_ancestors(g[f"{flavour}_{N}_wn"], g[f"{flavour}_{N}_u"])
            assert np.array_equal(idx, g[f"{flavour}_{N}_idx"])
    assert np.array_equal(O.multinomial_ancestors(g["dyadic_wn"], g["dyadic_u"]), g["dyadic_idx"])


def test_lkernels_weights_estimates(golden):
    g = golden("lkernel_weights")
    for D in (4, 13, 100):
        r_new, x_new = g[f"gaussL_{D}_r_new"], g[f"gaussL_{D}_x_new"]
        np.testing.assert_allclose(O.gaussian_lkernel(r_new, x_new), g[f"gaussL_{D}_L"], rtol=1e-9)
        np.testing.assert_allclose(O.forward_lkernel(r_new), g[f"fwdL_{D}"], rtol=1e-13)
        np.testing.assert_allclose(O.std_normal_logpdf(r_new), g[f"qlogpdf_{D}"], rtol=1e-13)
    for tag, tname in (("a", "arma"), ("b", "PRMwCD")):
        wn, logZ = O.normalise_weights(g[f"weights_{tag}_logw"])
        assert np.array_equal(wn, g[f"weights_{tag}_wn"]) and logZ == g[f"weights_{tag}_logZ"]
        assert O.calculate_ess(wn) == g[f"weights_{tag}_ess"]
        x = g[f"weights_{tag}_x"]
        m, v = O.estimate(make_target(tname).constrain(x), wn)
        np.testing.assert_allclose(m, g[f"weights_{tag}_mean_c"], rtol=1e-14)
        np.testing.assert_allclose(v, g[f"weights_{tag}_var_c"], rtol=1e-14)
        m, v = O.estimate(x, wn)
        np.testing.assert_allclose(m, g[f"weights_{tag}_mean_u"], rtol=1e-14)


def test_tempering_bisect(golden):
    g = golden("tempering")
    for j in range(4):
        phi = O.calculate_phi(g[f"temper_{j}_loglik"], g[f"temper_{j}_logpri"], g[f"temper_{j}_lp_old"],
                              float(g[f"temper_{j}_old_phi"]), len(g[f"temper_{j}_loglik"]))
        assert phi == float(g[f"temper_{j}_phi"])          # bit-exact restatement of scipy bisect
    for a, root in zip(g["bisect_a"], g["bisect_root"]):
        assert O.bisect(lambda p: np.tanh(3 * (a - p)) + 0.1 * (a - p), 0.0, 1.0) == root


RUNS = [("arma_forward", "arma", {}, "forwardsLKernel"), ("arma_gauss", "arma", {}, "GaussianApproxLKernel"),
        ("arma_asymptotic", "arma", {}, "asymptoticLKernel"), ("arma_forward_tempered", "arma", {}, "forwardsLKernel"),
        ("PRMwCD_asymptotic", "PRMwCD", {}, "asymptoticLKernel"),
        ("gauss8_gaussL", "gauss", {"dim": 8}, "GaussianApproxLKernel")]


@pytest.mark.parametrize("name,tname,kw,lk", RUNS)
def test_full_run_matches_reference(golden, name, tname, kw, lk):
    """Whole SMCSampler runs: oracle loop vs the unmodified reference driven with the same Philox streams."""
    g = golden("runs")
    N, K, eps, temp = g[f"{name}_cfg"]
    s = O.OracleSMC(int(K), int(N), tname, float(eps), lk, bool(temp), seed=10, target_kw=kw, nthreads=4).run()
    assert np.array_equal(s.n_leapfrog, g[f"{name}_leapfrogs"])
    np.testing.assert_allclose(s.x_saved[0], g[f"{name}_x_first"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(s.phi, g[f"{name}_phi"], rtol=1e-9)
    np.testing.assert_allclose(s.ess, g[f"{name}_ess"], rtol=1e-7)
    np.testing.assert_allclose(s.log_likelihood, g[f"{name}_log_likelihood"], rtol=1e-8)
    np.testing.assert_allclose(s.acceptance_rate, g[f"{name}_acceptance_rate"])
    np.testing.assert_allclose(s.x_saved[int(K)], g[f"{name}_x_final"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(s.logw_saved[int(K)], g[f"{name}_logw_final"], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(s.mean_estimate, g[f"{name}_mean_estimate"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(s.variance_estimate, g[f"{name}_variance_estimate"], rtol=1e-5, atol=1e-10)


def test_oracle_densities_equal_an_independent_scipy_restatement_of_the_stan_programs():
    """BridgeStan cannot be installed here, so the oracle's model arithmetic is pinned twice over by independent
    restatements: mpmath (golden fixture, 50 digits) and -- this test -- the `.stan` programs written out with
    scipy.stats' lpdf/lpmf implementations (all normalising constants kept, as `target +=` does) plus the log-Jacobian
    of the `<lower=0>` transform (Stan reference manual; BridgeStan's default jacobian=True).  Gradients are checked
    against Richardson-extrapolated central differences of the scipy form."""
    import json
    from scipy import stats
    from smcnuts.model.device_model import DATA_DIR          # the shipped arma.json / repaired PRMwCD.json
    arma_json, prm_json = DATA_DIR / "arma" / "arma.json", DATA_DIR / "PRMwCD" / "PRMwCD.json"
    y_arma = np.asarray(json.loads(arma_json.read_text())["y"], dtype=float)
    prm = json.loads(prm_json.read_text())
    y_prm = np.asarray(prm["y"], dtype=float)
    Xk = np.asarray(prm["Xkernel"], dtype=float).reshape(int(prm["N"]), int(prm["Clength"]))
    q = float(prm["q"])

    def arma_logp(x, phi):                                   # arma.stan:14-31
        mu, beta, theta, s = x
        sigma = np.exp(s)
        lp = stats.norm.logpdf(mu, 0, 10) + stats.norm.logpdf(beta, 0, 2) + stats.norm.logpdf(theta, 0, 2)
        lp += stats.cauchy.logpdf(sigma, 0, 2.5) + s         # + log-Jacobian of sigma = exp(s)
        err = np.empty_like(y_arma)
        err[0] = y_arma[0] - (mu + beta * mu)
        for t in range(1, len(y_arma)):
            err[t] = y_arma[t] - (mu + beta * y_arma[t - 1] + theta * err[t - 1])
        return lp + phi * stats.norm.logpdf(err, 0, sigma).sum()

    def prm_logp(x, phi):                                    # PRMwCD.stan:17-39
        B, g = x[:12], x[12]
        Gam = np.exp(g)
        lp = stats.invgamma.logpdf(Gam, 2, scale=1.3) + g
        eta = B[0] + Xk @ B[1:]
        lp += phi * stats.poisson.logpmf(y_prm, np.exp(eta)).sum()
        return lp + np.sum(-np.log(Gam) - np.abs(B[1:] / Gam) ** q)

    def num_grad(f, x):
        g = np.empty_like(x)
        for i in range(len(x)):
            def d(h):
                e = np.zeros_like(x); e[i] = h
                return (f(x + e) - f(x - e)) / (2 * h)
            h = 1e-4 * max(1.0, abs(x[i]))
            g[i] = (4 * d(h / 2) - d(h)) / 3
        return g

    rng = np.random.default_rng(8)
    t = O.COracleTarget("arma")
    X = rng.normal(size=(12, 4)) * 0.2 + np.array([0.0, 0.9, 0.0, -1.7])
    for phi in (0.0, 0.37, 1.0):
        np.testing.assert_allclose(t.logpdf(X, phi), [arma_logp(x, phi) for x in X], rtol=1e-12)
        G = t.logpdfgrad(X, phi)
        for x, gr in zip(X[:4], G[:4]):
            np.testing.assert_allclose(gr, num_grad(lambda v: arma_logp(v, phi), x), rtol=2e-6, atol=1e-6)
    t = O.COracleTarget("PRMwCD")
    X = rng.normal(size=(12, 13)) * 0.3
    X[:, 12] = rng.normal(size=12) * 0.3 - 1.0
    for phi in (0.0, 0.37, 1.0):
        np.testing.assert_allclose(t.logpdf(X, phi), [prm_logp(x, phi) for x in X], rtol=1e-12)
        G = t.logpdfgrad(X, phi)
        for x, gr in zip(X[:4], G[:4]):
            np.testing.assert_allclose(gr, num_grad(lambda v: prm_logp(v, phi), x), rtol=2e-6, atol=1e-5)


@pytest.mark.parametrize("case", ["arma", "arma_tempered", "arma_prior", "PRMwCD", "PRMwCD_tempered", "gauss8", "gauss100"])
def test_c_nuts_devmath_mode_matches_reference_devmath_fixtures(golden, case):
    """tests/golden/nuts_devmath.npz: the UNMODIFIED reference transitions with the devmath target (exp / log = the
    kernels' table-driven algorithms, oracle/devmath.h).  The C oracle in the same mode reproduces them exactly; these
    are the fixtures the parity device build is held to on the GPU (tests/test_gpu_parity_build.py)."""
    g = golden("nuts_devmath")
    tname, kw = case.split("_")[0], {}
    if tname.startswith("gauss"):
        kw, tname = {"dim": int(tname[5:])}, "gauss"
    t = O.COracleTarget(tname, **kw)
    x0, r0 = g[f"{case}_x0"], g[f"{case}_r0"]
    eps, phi, it, seed = float(g[f"{case}_eps"]), float(g[f"{case}_phi"]), int(g[f"{case}_iteration"]), int(g[f"{case}_seed"])
    with O.devmath():
        o = t.nuts_batch(x0, r0, eps, phi, 10, seed=seed, iteration=it, accrej=True)
    assert np.array_equal(o["n_leapfrog"], g[f"{case}_n_leapfrog"])
    acc = g[f"{case}_accepted"]
    assert np.array_equal(o["accepted"].astype(bool), acc)
    np.testing.assert_allclose(o["x_new"][acc], g[f"{case}_x_new"][acc], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(o["r_new"][acc], g[f"{case}_r_new"][acc], rtol=1e-12, atol=1e-14)


def test_devmath_tables_and_density_agreement():
    """oracle/devmath_tables.h is what its generator writes, equals the tables of csrc/common.cuh, and the devmath mode
    changes the model densities by rounding only (<= 1e-13 relative on A, B and the gradient)."""
    import re
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "oracle"))
    import gen_devmath_tables as G
    assert (root / "oracle" / "devmath_tables.h").read_text() == G.render()
    text = (root / "smc-nuts_b200" / "csrc" / "common.cuh").read_text()

    def table(name):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", text, re.S).group(1)
        return [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-fA-F.]+p[+-]?\d+", body)]
    exp_t, inv_c, log_c = G.tables()
    assert table("kExpT") == exp_t and table("kLogInvC") == inv_c and table("kLogC") == log_c
    rng = np.random.default_rng(2)
    for name in ("arma", "PRMwCD"):
        t = O.COracleTarget(name)
        x = rng.normal(size=(200, t.dim)) * 0.5
        a = t.split(x)
        with O.devmath():
            b = t.split(x)
        for u, v in zip(a, b):
            fin = np.isfinite(u) & np.isfinite(v)
            assert np.array_equal(np.isfinite(u), np.isfinite(v))
            np.testing.assert_allclose(u[fin], v[fin], rtol=1e-13, atol=1e-13)
