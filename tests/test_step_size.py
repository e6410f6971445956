"""Step-size adaptation (README.md:66-67 "future updates" of the reference): the host-side dual
averaging against the oracle restatement (CPU), the kernel's acceptance statistic against the C oracle and an adaptive
SMC run that reaches the target acceptance (GPU)."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import smc_oracle as O

ROOT = Path(__file__).resolve().parents[1]


def _adapter_cls():
    # step_size.py has no device dependency: load the file itself so that the CPU suite does not import the package's
    # CUDA plumbing
    spec = importlib.util.spec_from_file_location("smcb_step_size", ROOT / "smc-nuts_b200/smcnuts/proposal/step_size.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DualAveragingStepSize


def test_dual_averaging_matches_oracle_restatement():
    cls = _adapter_cls()
    rng = np.random.default_rng(5)
    for eps0, delta in ((0.01, 0.8), (0.3, 0.65)):
        acc = rng.uniform(0.2, 1.0, size=25)
        d = cls(eps0, target_accept=delta)
        got = [d.step_size] + [d.update(a) for a in acc]
        ref, ref_bar = O.dual_averaging(eps0, acc, delta=delta)
        np.testing.assert_allclose(got, ref, rtol=1e-13)
        np.testing.assert_allclose(d.averaged(), ref_bar, rtol=1e-13)


def test_dual_averaging_converges_on_a_synthetic_acceptance_curve():
    """acceptance(eps) = exp(-(eps/0.2)^2): the fixed point of delta = 0.8 is eps* = 0.2 sqrt(-log 0.8)."""
    d = _adapter_cls()(0.001, target_accept=0.8)
    for _ in range(200):
        d.update(np.exp(-(d.step_size / 0.2) ** 2))
    assert abs(d.averaged() - 0.2 * np.sqrt(-np.log(0.8))) < 0.01
    assert d.update(float("nan")) < d.step_size * 1.0000001     # NaN statistic counts as 0 -> never grows the step
    with pytest.raises(ValueError):
        _adapter_cls()(0.1, target_accept=1.5)


def test_oracle_accept_stat_is_a_probability_and_falls_with_step_size():
    t = O.COracleTarget("arma")
    rng = np.random.default_rng(0)
    x = rng.normal(size=(256, 4)) * 0.05 + np.array([0.0, 0.9, 0.0, -1.7])
    r = rng.normal(size=(256, 4))
    means = []
    for eps in (0.005, 0.02, 0.06):
        a = t.nuts_batch(x, r, eps, seed=3)["accept_stat"]
        assert np.all((a >= 0) & (a <= 1))
        means.append(a.mean())
    assert means[0] > means[1] > means[2] and means[0] > 0.97


def test_oracle_diagonal_metric_is_nuts_on_the_rescaled_model():
    """Diagonal metric = identity-metric NUTS on z = x / scale.  For the Gaussian the change of variables is another
    Gaussian (precision S P S), so the oracle's metric path can be checked against its plain path on that model."""
    D = 8
    idx = np.arange(D)
    P = np.linalg.inv(0.9 ** np.abs(idx[:, None] - idx[None, :]))
    P = 0.5 * (P + P.T)
    s = np.linspace(0.4, 2.5, D)
    t = O.COracleTarget("gauss", precision=P)
    tz = O.COracleTarget("gauss", precision=(s[:, None] * P) * s[None, :])
    rng = np.random.default_rng(2)
    x, r = rng.normal(size=(300, D)), rng.normal(size=(300, D))
    a = t.nuts_batch(x, r, 0.1, 0.7, 8, seed=5, scale=s)
    b = tz.nuts_batch(x / s, r, 0.1, 0.7, 8, seed=5)
    same = a["n_leapfrog"] == b["n_leapfrog"]
    assert same.mean() > 0.99                       # two roundings of the same trajectory
    np.testing.assert_allclose(a["x_new"][same], b["x_new"][same] * s, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(a["lp_new"][same], b["lp_new"][same], rtol=1e-9, atol=1e-9)
    ident = t.nuts_batch(x, r, 0.1, 0.7, 8, seed=5, scale=np.ones(D))
    plain = t.nuts_batch(x, r, 0.1, 0.7, 8, seed=5)
    assert np.array_equal(ident["x_new"], plain["x_new"]) and np.array_equal(ident["n_leapfrog"], plain["n_leapfrog"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,eps", [("arma", 0.02), ("PRMwCD", 0.01), ("gauss8", 0.1), ("gauss100", 0.1)])
def test_kernel_diagonal_metric_matches_oracle(name, eps):
    from smcnuts.distributions import StdNormal
    from smcnuts.model.device_model import make_model
    from smcnuts.proposal.nuts import NUTSProposal
    if name.startswith("gauss"):
        d = int(name[5:])
        m, t = make_model("gauss", dim=d), O.COracleTarget("gauss", dim=d)
    else:
        m, t = make_model(name), O.COracleTarget(name)
    rng = np.random.default_rng(4)
    N, D = 2048, m.dim
    x = rng.normal(size=(N, D)) * 0.1
    if name == "arma":
        x += np.array([0.0, 0.9, 0.0, -1.7])
    r = rng.normal(size=(N, D))
    s = np.exp(rng.uniform(-0.7, 0.7, size=D))
    k = NUTSProposal(m, StdNormal(D), eps, rng=7, max_tree_depth=6)
    plain, _ = k.rvs(x, r, 0.8)
    m.set_metric_scale(s)
    try:
        k.iteration = 0
        xn, rn = k.rvs(x, r, 0.8)
        ref = t.nuts_batch(x, r, eps, 0.8, 6, seed=7, iteration=0, nthreads=4, scale=s)
        same = k.last["n_leapfrog"].cpu().numpy() == ref["n_leapfrog"]
        assert same.mean() > 0.97, same.mean()
        np.testing.assert_allclose(xn[same], ref["x_new"][same], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(rn[same], ref["r_new"][same], rtol=1e-6, atol=1e-8)
        # A / B by-products are the x-space split at the returned point
        A, B = (v.cpu().numpy() for v in m.split(xn))
        np.testing.assert_allclose(k.last["A_new"].cpu().numpy()[same], A[same], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(k.last["B_new"].cpu().numpy()[same], B[same], rtol=1e-9, atol=1e-7)
    finally:
        m.set_metric_scale(None)
    k.iteration = 0
    again, _ = k.rvs(x, r, 0.8)
    assert np.array_equal(again, plain)            # identity metric restored: the reference path, bit for bit


@pytest.mark.gpu
def test_mass_matrix_adaptation_recovers_the_scales_of_an_anisotropic_gaussian():
    from smcnuts.distributions import StdNormal
    from smcnuts.model.device_model import make_model
    from smcnuts.smc_sampler import SMCSampler
    D = 16
    sd = np.exp(np.linspace(-2.0, 1.0, D))                    # standard deviations from 0.14 to 2.7
    m = make_model("gauss", precision=np.diag(1.0 / sd ** 2))
    kw = dict(K=24, N=8192, target=m, sample_proposal=StdNormal(D), momentum_proposal=StdNormal(D),
              lkernel="forwardsLKernel", tempering=False, rng=10)
    s1 = SMCSampler(step_size=0.02, adapt_step_size=16, adapt_mass_matrix=True, **kw)
    s1.sample(show_progress=False)
    assert m.metric_scale is None                             # the model object is handed back with the identity metric
    np.testing.assert_allclose(s1.metric_scale, sd, rtol=0.25)
    np.testing.assert_allclose(np.sqrt(s1.variance_estimate[-1]), sd, rtol=0.35)   # importance-sampling estimate at N = 8192: noisy
    s0 = SMCSampler(step_size=0.02, adapt_step_size=16, **kw)
    s0.sample(show_progress=False)
    # with the metric the step size is not held down by the narrowest coordinate: far fewer leapfrogs per transition
    print('leapfrogs per iteration, last four: metric', s1.leapfrogs[20:], 'identity', s0.leapfrogs[20:], 'eps', s1.step_sizes[-1], s0.step_sizes[-1])
    assert s1.leapfrogs[20:].mean() < 0.5 * s0.leapfrogs[20:].mean()


@pytest.mark.gpu
@pytest.mark.parametrize("name,eps", [("arma", 0.02), ("PRMwCD", 0.01), ("gauss100", 0.1)])
def test_kernel_accept_stat_matches_oracle(name, eps):
    from smcnuts.distributions import StdNormal
    from smcnuts.model.device_model import make_model
    from smcnuts.proposal.nuts import NUTSProposal
    if name.startswith("gauss"):
        m, t = make_model("gauss", dim=100), O.COracleTarget("gauss", dim=100)
    else:
        m, t = make_model(name), O.COracleTarget(name)
    rng = np.random.default_rng(1)
    N, D = 2048, m.dim
    x = rng.normal(size=(N, D)) * 0.1
    if name == "arma":
        x += np.array([0.0, 0.9, 0.0, -1.7])
    r = rng.normal(size=(N, D))
    k = NUTSProposal(m, StdNormal(D), eps, rng=7, max_tree_depth=6)
    k.want_accept_stat = True
    k.rvs(x, r, 0.8)
    ref = t.nuts_batch(x, r, eps, 0.8, 6, seed=7, iteration=0, nthreads=4)
    got = k.last["accept_stat"].cpu().numpy()
    same = k.last["n_leapfrog"].cpu().numpy() == ref["n_leapfrog"]
    assert same.mean() > 0.97
    assert np.all((got >= 0) & (got <= 1))
    # same trees -> the same mean of min(1, exp(dH)) up to the rounding of the leaves' energies
    np.testing.assert_allclose(got[same], ref["accept_stat"][same], rtol=1e-6, atol=1e-9)
    # and switching the statistic on does not change the transition
    k2 = NUTSProposal(m, StdNormal(D), eps, rng=7, max_tree_depth=6)
    xn2, _ = k2.rvs(x, r, 0.8)
    assert np.array_equal(xn2, k.last["x_new"].cpu().numpy())


@pytest.mark.gpu
def test_adaptive_run_reaches_target_acceptance():
    from smcnuts.distributions import StdNormal
    from smcnuts.model.device_model import make_model
    from smcnuts.smc_sampler import SMCSampler
    m = make_model("gauss", dim=16)
    kw = dict(K=16, N=4096, target=m, sample_proposal=StdNormal(16), momentum_proposal=StdNormal(16),
              lkernel="forwardsLKernel", tempering=False, rng=10)
    s = SMCSampler(step_size=0.01, adapt_step_size=12, target_accept=0.8, **kw)
    s.sample(show_progress=False)
    assert s.step_sizes[0] == 0.01 and np.all(s.step_sizes[12:] == s.step_sizes[12])
    assert s.step_sizes[12] > 0.05                      # 0.01 is far too small for this target: the step grew
    assert abs(np.nanmean(s.accept_stat[8:12]) - 0.8) < 0.1
    assert np.all(np.isnan(s.accept_stat[12:]))
    assert s.leapfrogs[12:].mean() < 0.5 * s.leapfrogs[0]    # and the trees got shorter
    # default: the reference's constant step size
    s0 = SMCSampler(step_size=0.01, **kw)
    s0.sample(show_progress=False)
    assert np.all(s0.step_sizes == 0.01) and s0.step_size_adapter is None
