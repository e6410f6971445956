"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference on sys.path).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

What is pinned, and how the reference is driven:
  * The reference classes are imported unmodified.  The `target` they receive is duck-typed
    (oracle.smc_oracle.COracleTarget: the restated Stan densities; BridgeStan is not installable
    offline, so the model arithmetic itself is PARITY UNPINNED and is cross-checked against mpmath
    below instead).
  * The `rng` they receive is oracle.philox.ReplayRNG: the per-particle Philox4x32-10 streams the
    B200 kernels use, served through numpy's `.uniform()/.exponential()` names, so reference and
    device consume identical draws in the reference's own order.
  * `np.random.RandomState.choice` / `Generator.choice` are run for real and compared with the
    cdf/searchsorted restatement.

Fixtures are small (KBs) and committed; tests/test_oracle_golden.py checks the C/numpy oracle against
them on CPU, tests/test_gpu_*.py check the CUDA path against them on the B200.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

import mpmath as mp  # noqa: E402
from scipy.optimize import bisect as scipy_bisect  # noqa: E402
from scipy.stats import multivariate_normal  # noqa: E402

from oracle import philox  # noqa: E402
from oracle.models import make_target  # noqa: E402
from oracle.smc_oracle import COracleTarget  # noqa: E402

from smcnuts.estimate.estimate import Estimate  # noqa: E402  (reference)
from smcnuts.lkernel.forward_lkernel import ForwardLKernel  # noqa: E402
from smcnuts.lkernel.gaussian_lkernel import GaussianApproxLKernel  # noqa: E402
from smcnuts.proposal.nuts import NUTSProposal  # noqa: E402
from smcnuts.proposal.utils import hmc_accept_reject  # noqa: E402
from smcnuts.samples.samples import Samples  # noqa: E402
from smcnuts.smc_sampler import SMCSampler  # noqa: E402
from smcnuts.tempering.adaptive_tempering import ESSTempering  # noqa: E402

OUT = Path(__file__).resolve().parent
SEED = 10


# ------------------------------------------------------------------ helpers
class CountingTarget:
    """Pass-through target that counts gradient calls (one per leapfrog + one per transition)."""

    def __init__(self, t):
        self.t, self.dim, self.ngrad = t, t.dim, 0
        if hasattr(t, "constrained_dim"):
            self.constrained_dim, self.constrain = t.constrained_dim, t.constrain

    def logpdf(self, x, phi=1.0):
        return self.t.logpdf(x, phi)

    def logpdfgrad(self, x, phi=1.0):
        self.ngrad += 1
        return self.t.logpdfgrad(x, phi)


class StdNormal:
    """q0 / momentum N(0, I) whose .rvs(N) serves the Philox streams (run_experiments.py:110-111)."""

    def __init__(self, dim, stream):
        self.dim, self.stream, self.calls = dim, stream, 0
        self._mvn = multivariate_normal(mean=np.zeros(dim), cov=np.eye(dim))

    def rvs(self, N):
        z = philox.normals(SEED, self.calls, self.stream, np.arange(N), self.dim)
        self.calls += 1
        return z

    def logpdf(self, x):
        return self._mvn.logpdf(x)


class PerParticleKernel:
    """forward_kernel plugin that calls the reference's own per-particle methods, switching the
    ReplayRNG to particle i's stream first (the reference loops particles in order, nuts.py:50-53)."""

    def __init__(self, ref_kernel, momentum, accrej):
        self.k, self.momentum_proposal, self.accrej = ref_kernel, momentum, accrej
        self.step_size, self.target, self.rng = ref_kernel.step_size, ref_kernel.target, ref_kernel.rng
        self.iteration = 0
        self.leapfrogs = []

    def logpdf(self, r):
        return self.k.logpdf(r)

    def rvs(self, x_cond, r_cond, phi=1.0):
        rng = self.k.rng
        rng.iteration = self.iteration
        x_prime, r_prime = np.zeros_like(x_cond), np.zeros_like(r_cond)
        g0 = self.target.ngrad
        for i in range(len(x_cond)):
            rng.set_particle(i, stream=philox.STREAM_NUTS)
            x_prime[i], r_prime[i] = self.k.generate_nuts_samples(x_cond[i], r_cond[i], phi=phi)
        self.leapfrogs.append(self.target.ngrad - g0 - len(x_cond))
        if self.accrej:  # nuts_acc_rej.py:42-49
            accepted = np.array([False] * len(x_prime))
            for i in range(len(x_prime)):
                rng.set_particle(i, stream=philox.STREAM_ACCREJ)
                accepted[i] = hmc_accept_reject(self.target.logpdf, x_cond[i], x_prime[i], r_cond[i], r_prime[i],
                                                phi, rng=rng)
            x_prime[~accepted] = x_cond[~accepted]
            r_prime[~accepted] = r_cond[~accepted]
        self.iteration += 1
        return x_prime, r_prime


class ChoiceRNG:
    """Serves `rng.choice(i, N, p=wn)` (samples.py:139, estimate_from_tempered.py:43) from the Philox
    resampling stream through numpy's own algorithm (cdf + searchsorted 'right')."""

    def __init__(self, stream, kernel=None):
        self.stream, self.kernel, self.calls = stream, kernel, 0

    def choice(self, a, size, p):
        it = self.kernel.iteration if self.kernel is not None else self.calls
        self.calls += 1
        u = philox.uniform(SEED, it, self.stream, np.arange(size), 0)
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        return a[np.searchsorted(cdf, u, side="right")]


# ------------------------------------------------------------------ A. model values (numpy vs mpmath)
def golden_models():
    rng = np.random.default_rng(1)
    out = {}
    for name, kw in (("arma", {}), ("PRMwCD", {}), ("gauss", {"dim": 8}), ("gauss100", {"dim": 100})):
        t = make_target("gauss" if name.startswith("gauss") else name, **kw)
        D = t.dim
        X = rng.normal(size=(48, D)) * 0.7
        if name == "arma":
            X[:8] = np.array([0.0068, 0.957, -0.034, np.log(0.1666)]) + 0.02 * rng.normal(size=(8, 4))
            X[8] = [0.3, -1.2, 0.8, 4.0]
            X[9] = [0.3, -1.2, 0.8, -6.0]
            X[10] = [1e3, 2.0, 1.5, 0.1]      # exploding recurrence
            X[11] = [0.0, 0.0, 0.0, 800.0]    # sigma = inf  -> -inf
            X[12] = [0.0, 0.0, 0.0, -800.0]   # sigma = 0    -> -inf
        if name == "PRMwCD":
            X[8, :12] *= 6.0
            X[9, 12] = -4.0
            X[10, 0] = 800.0                  # exp(eta) = inf -> -inf
            X[11, 12] = 800.0                 # Gamma = inf -> -inf
        A, B, gA, gB = t.split(X)
        out[f"{name}_X"], out[f"{name}_A"], out[f"{name}_B"] = X, A, B
        out[f"{name}_gA"], out[f"{name}_gB"] = gA, gB
        for phi in (0.0, 0.37, 1.0):
            out[f"{name}_lp_{phi}"] = t.logpdf(X, phi)
            out[f"{name}_grad_{phi}"] = t.logpdfgrad(X, phi)
    # mpmath 50-digit values of arma / PRMwCD at the first 8 rows, phi = 0.37
    mp.mp.dps = 50
    ta, tp = make_target("arma"), make_target("PRMwCD")

    def arma_mp(x, phi):
        mu, beta, theta, s = [mp.mpf(float(v)) for v in x]
        sigma = mp.e ** s
        y = [mp.mpf(float(v)) for v in ta.y]

        def nl(v, sd):
            return -mp.log(2 * mp.pi) / 2 - mp.log(sd) - v * v / (2 * sd * sd)
        lp = nl(mu, 10) + nl(beta, 2) + nl(theta, 2) + s
        lp += -mp.log(mp.pi) - mp.log(mp.mpf("2.5")) - mp.log(1 + (sigma / mp.mpf("2.5")) ** 2)
        e = y[0] - (mu + beta * mu)
        S = e * e
        for k in range(1, len(y)):
            e = y[k] - (mu + beta * y[k - 1] + theta * e)
            S += e * e
        return lp + phi * (-len(y) * mp.log(2 * mp.pi) / 2 - len(y) * s - S / (2 * sigma ** 2))

    def prm_mp(x, phi):
        Bv = [mp.mpf(float(v)) for v in x[:12]]
        g = mp.mpf(float(x[12]))
        G = mp.e ** g
        lp = 2 * mp.log(mp.mpf("1.3")) - mp.loggamma(2) - 3 * mp.log(G) - mp.mpf("1.3") / G + g
        for i in range(100):
            eta = Bv[0] + sum(Bv[j + 1] * mp.mpf(float(tp.X[i, j])) for j in range(11))
            lam = mp.e ** eta
            lp += phi * (int(tp.y[i]) * mp.log(lam) - lam - mp.loggamma(int(tp.y[i]) + 1))
        for i in range(1, 12):
            lp += -mp.log(G) - abs(Bv[i] / G) ** mp.mpf(str(tp.q))
        return lp

    phi = mp.mpf("0.37")
    out["arma_mp_lp_0.37"] = np.array([float(arma_mp(x, phi)) for x in out["arma_X"][:8]])
    out["PRMwCD_mp_lp_0.37"] = np.array([float(prm_mp(x, phi)) for x in out["PRMwCD_X"][:8]])
    np.savez_compressed(OUT / "models.npz", **out)
    print("models.npz", len(out), "arrays;",
          "arma mp rel", np.max(np.abs(out["arma_mp_lp_0.37"] / out["arma_lp_0.37"][:8] - 1)),
          "prm mp rel", np.max(np.abs(out["PRMwCD_mp_lp_0.37"] / out["PRMwCD_lp_0.37"][:8] - 1)))


# ------------------------------------------------------------------ B. single NUTS transitions
def golden_nuts(devmath=False):
    """devmath=True: the duck-typed target (and nothing else) evaluates exp / log with the table-driven device
    algorithms (oracle/devmath.h) -> nuts_devmath.npz, the fixtures the PARITY device build must reproduce exactly."""
    from oracle import smc_oracle as _O
    _O.lib().orc_set_devmath(int(devmath))
    cases = [
        # name, target, kwargs, eps, phi, centre, spread, P, iteration
        ("arma", "arma", {}, 0.01, 1.0, [0.0068, 0.957, -0.034, np.log(0.1666)], 0.02, 96, 3),
        ("arma_tempered", "arma", {}, 0.01, 0.05, [0.0, 0.5, 0.0, -1.0], 0.3, 48, 4),
        ("arma_prior", "arma", {}, 0.01, 1.0, [0.0, 0.0, 0.0, 0.0], 1.0, 64, 0),
        ("PRMwCD", "PRMwCD", {}, 0.01, 1.0, None, 0.05, 24, 5),
        ("PRMwCD_tempered", "PRMwCD", {}, 0.01, 0.01, None, 0.3, 24, 1),
        ("gauss8", "gauss", {"dim": 8}, 0.1, 1.0, [0.0] * 8, 1.0, 64, 2),
        ("gauss100", "gauss", {"dim": 100}, 0.1, 1.0, [0.0] * 100, 1.0, 16, 2),
    ]
    prm_centre = np.array([0.8925, 0.0946, 1.3969, 0.1151, -1.4883, -0.0898, 0.6766, -1.7521, -0.3014, 1.6721,
                           -0.1868, -0.1491, np.log(0.3326)])
    out = {}
    gen = np.random.default_rng(7)
    for name, tname, kw, eps, phi, centre, spread, P, it in cases:
        ct = CountingTarget(COracleTarget(tname, **kw))
        D = ct.dim
        centre = prm_centre if centre is None else np.asarray(centre, dtype=float)
        x0 = centre[None, :] + spread * gen.normal(size=(P, D))
        r0 = philox.normals(SEED, it, philox.STREAM_MOMENTUM, np.arange(P), D)
        rng = philox.ReplayRNG(SEED, iteration=it)
        prop = NUTSProposal(target=ct, momentum_proposal=None, step_size=eps, rng=rng)
        xn, rn = np.zeros_like(x0), np.zeros_like(r0)
        nleap, ndraw = np.zeros(P, dtype=np.int64), np.zeros(P, dtype=np.int64)
        acc = np.zeros(P, dtype=bool)
        for i in range(P):
            rng.set_particle(i, stream=philox.STREAM_NUTS)
            g0 = ct.ngrad
            with np.errstate(all="ignore"):
                xn[i], rn[i] = prop.generate_nuts_samples(x0[i], r0[i], phi=phi)
            nleap[i], ndraw[i] = ct.ngrad - g0 - 1, rng.pos
            rng.set_particle(i, stream=philox.STREAM_ACCREJ)
            acc[i] = hmc_accept_reject(ct.logpdf, x0[i], xn[i], r0[i], rn[i], phi, rng=rng)
        for k, v in dict(x0=x0, r0=r0, x_new=xn, r_new=rn, n_leapfrog=nleap, n_draws=ndraw, accepted=acc,
                         eps=eps, phi=phi, iteration=it, seed=SEED).items():
            out[f"{name}_{k}"] = np.asarray(v)
        print(f"nuts {name}: leapfrogs min/mean/max = {nleap.min()}/{nleap.mean():.1f}/{nleap.max()}, "
              f"moved {np.mean(np.any(xn != x0, axis=1)):.2f}, accepted {acc.mean():.2f}")
    np.savez_compressed(OUT / ("nuts_devmath.npz" if devmath else "nuts.npz"), **out)
    _O.lib().orc_set_devmath(0)


# ------------------------------------------------------------------ C. rng.choice == cdf/searchsorted
def golden_choice():
    out = {}
    gen = np.random.default_rng(3)
    for j, N in enumerate((1, 2, 7, 100, 4096)):
        w = gen.exponential(size=N) ** 3
        if N > 4:
            w[gen.integers(0, N, size=N // 5)] = 0.0
        wn = w / w.sum()
        a = np.linspace(0, N - 1, N, dtype=int)
        for flavour in ("RandomState", "Generator"):
            mk = (lambda: np.random.RandomState(123 + j)) if flavour == "RandomState" else \
                 (lambda: np.random.default_rng(123 + j))
            idx = mk().choice(a, N, p=wn)
            g2 = mk()
            u = g2.random_sample(N) if flavour == "RandomState" else g2.random(N)
            cdf = np.cumsum(wn)
            cdf /= cdf[-1]
            assert np.array_equal(idx, np.searchsorted(cdf, u, side="right")), (flavour, N)
            out[f"{flavour}_{N}_wn"], out[f"{flavour}_{N}_u"], out[f"{flavour}_{N}_idx"] = wn, u, idx
    # dyadic weights: every partial sum is exact in fp64, so any summation order gives the same cdf
    N = 1 << 12
    k = gen.integers(0, 6, size=N)
    w = np.ldexp(1.0, -k.astype(int))
    w[gen.integers(0, N, size=N // 8)] = 0.0
    tot = w.sum()
    # make the total a power of two by topping up the last weight with a dyadic value
    target_total = 2.0 ** np.ceil(np.log2(tot))
    w[-1] += target_total - tot
    wn = w / target_total
    u = np.random.RandomState(5).random_sample(N)
    out["dyadic_wn"], out["dyadic_u"] = wn, u
    out["dyadic_idx"] = np.random.RandomState(5).choice(np.arange(N), N, p=wn)
    np.savez_compressed(OUT / "choice.npz", **out)
    print("choice.npz ok (numpy choice == searchsorted(cumsum) for RandomState and Generator)")


# ------------------------------------------------------------------ D. L-kernels, weights, estimates
def golden_lkernel_weights():
    out = {}
    gen = np.random.default_rng(11)
    for D, N in ((4, 64), (13, 200), (100, 300)):
        class T:  # noqa: N801
            dim = D
        Lc = np.linalg.cholesky(0.5 ** np.abs(np.arange(D)[:, None] - np.arange(D)[None, :]))
        x_new = gen.normal(size=(N, D)) @ Lc.T
        r_new = 0.7 * gen.normal(size=(N, D)) + 0.3 * x_new
        out[f"gaussL_{D}_L"] = GaussianApproxLKernel(T, N).calculate_L(r_new, x_new)
        out[f"gaussL_{D}_r_new"], out[f"gaussL_{D}_x_new"] = r_new, x_new
        mom = multivariate_normal(mean=np.zeros(D), cov=np.eye(D))
        out[f"fwdL_{D}"] = ForwardLKernel(T, mom).calculate_L(r_new, None)
        out[f"qlogpdf_{D}"] = NUTSProposal(None, mom, 0.1).logpdf(r_new)
    # normalise / ESS / estimate through the reference Samples + Estimate
    for tag, N, D in (("a", 100, 4), ("b", 1000, 13)):
        logw = gen.normal(size=N) * 3.0
        logw[gen.integers(0, N, size=N // 10)] = -np.inf
        x = gen.normal(size=(N, D))
        s = Samples(N, D, None, None, None, "asymptoticLKernel", False, None)
        s.logw = logw
        s.normalise_weights()
        s.calculate_ess()
        ta = make_target("arma") if D == 4 else make_target("PRMwCD")
        mean_c, var_c = Estimate(ta).return_estimate(x, s.wn)
        mean_u, var_u = Estimate(ta).return_estimate_unconstrained(x, s.wn)
        for k, v in dict(logw=logw, x=x, wn=s.wn, logZ=s.log_likelihood, ess=s.ess, mean_c=mean_c, var_c=var_c,
                         mean_u=mean_u, var_u=var_u).items():
            out[f"weights_{tag}_{k}"] = np.asarray(v)
    np.savez_compressed(OUT / "lkernel_weights.npz", **out)
    print("lkernel_weights.npz", len(out))


# ------------------------------------------------------------------ D2. Gaussian-approx L-kernel with a singular C_xx
def golden_lkernel_degenerate():
    """Cases where cov(x_new) is rank deficient, so that np.linalg.pinv (gaussian_lkernel.py:64-75) is not an inverse:
    N <= D, and N > D with the particles collapsed onto few distinct rows (heavy duplication after resampling)."""
    out = {}
    gen = np.random.default_rng(17)
    cases = {"n_le_d_4": (4, 4), "n_le_d_13": (13, 10), "dup_13": (13, 64), "dup_4": (4, 50)}
    for tag, (D, N) in cases.items():
        class T:  # noqa: N801
            dim = D
        if tag.startswith("dup"):
            distinct = gen.normal(size=(3 if D == 4 else 6, D))
            x_new = distinct[gen.integers(0, len(distinct), size=N)]
        else:
            x_new = gen.normal(size=(N, D))
        r_new = gen.normal(size=(N, D))
        with np.errstate(all="ignore"):
            out[f"{tag}_L"] = GaussianApproxLKernel(T, N).calculate_L(r_new, x_new)
        out[f"{tag}_r_new"], out[f"{tag}_x_new"] = r_new, x_new
        print(f"degenerate {tag}: L in [{out[f'{tag}_L'].min():.6g}, {out[f'{tag}_L'].max():.6g}]")
    np.savez_compressed(OUT / "lkernel_degenerate.npz", **out)


# ------------------------------------------------------------------ E. tempering / bisect
def golden_tempering():
    out = {}
    gen = np.random.default_rng(21)

    class ArrTarget:
        """logpdf(x, phi) = logpri + phi*loglik for stored arrays (x is ignored)."""
        def __init__(self, logpri, loglik):
            self.logpri, self.loglik = logpri, loglik

        def logpdf(self, x, phi=1.0):
            with np.errstate(invalid="ignore"):
                return self.logpri + phi * self.loglik

    for j, (N, scale, old_phi) in enumerate(((100, 50.0, 0.0), (1000, 400.0, 0.013), (4096, 3.0, 0.4), (512, 0.05, 0.7))):
        logpri = gen.normal(size=N) * 2.0 - 5.0
        loglik = -np.abs(gen.normal(size=N)) * scale
        t = ArrTarget(logpri, loglik)
        lp_old = t.logpdf(None, old_phi)
        phi = ESSTempering(N, t, alpha=0.5).calculate_phi([None, lp_old, old_phi])
        for k, v in dict(logpri=logpri, loglik=loglik, lp_old=lp_old, old_phi=old_phi, phi=phi).items():
            out[f"temper_{j}_{k}"] = np.asarray(v)
        print(f"tempering case {j}: old_phi={old_phi} -> phi={phi!r}")
    # scipy bisect itself on smooth functions (restatement check)
    roots = []
    for a in np.linspace(0.05, 0.95, 19):
        roots.append(scipy_bisect(lambda p: np.tanh(3 * (a - p)) + 0.1 * (a - p), 0.0, 1.0))
    out["bisect_a"], out["bisect_root"] = np.linspace(0.05, 0.95, 19), np.array(roots)
    np.savez_compressed(OUT / "tempering.npz", **out)


# ------------------------------------------------------------------ F. full reference runs
def golden_runs():
    out = {}
    cfgs = [
        ("arma_forward", "arma", {}, 0.01, "forwardsLKernel", False, 100, 10),
        ("arma_gauss", "arma", {}, 0.01, "GaussianApproxLKernel", False, 100, 10),
        ("arma_asymptotic", "arma", {}, 0.01, "asymptoticLKernel", True, 100, 10),
        ("arma_forward_tempered", "arma", {}, 0.01, "forwardsLKernel", True, 64, 6),
        ("PRMwCD_asymptotic", "PRMwCD", {}, 0.01, "asymptoticLKernel", True, 48, 8),
        ("gauss8_gaussL", "gauss", {"dim": 8}, 0.1, "GaussianApproxLKernel", False, 128, 8),
    ]
    for name, tname, kw, eps, lk, temp, N, K in cfgs:
        ct = CountingTarget(COracleTarget(tname, **kw))
        D = ct.dim
        q0, mom = StdNormal(D, philox.STREAM_INIT), StdNormal(D, philox.STREAM_MOMENTUM)
        nuts_rng = philox.ReplayRNG(SEED)
        # SMCSampler hands `rng` to the proposal, to Samples (for .choice) and to the estimator.  One object
        # must serve all three: ReplayRNG for uniform/exponential, ChoiceRNG.choice for resampling.
        with np.errstate(all="ignore"):
            smc = SMCSampler(K=K, N=N, target=ct, step_size=eps, sample_proposal=q0, momentum_proposal=mom,
                             lkernel=lk, tempering=temp, rng=nuts_rng)
            kern = PerParticleKernel(smc.samples.forward_kernel, mom, accrej=(lk == "asymptoticLKernel"))
            smc.samples.forward_kernel = kern
            smc.samples.rng = ChoiceRNG(philox.STREAM_RESAMPLE, kern)
            if lk == "asymptoticLKernel":
                smc.estimator.rng = ChoiceRNG(philox.STREAM_ESTIMATE)
            smc.sample(show_progress=False)
        for k in ("ess", "log_likelihood", "phi", "acceptance_rate", "mean_estimate", "variance_estimate"):
            out[f"{name}_{k}"] = np.asarray(getattr(smc, k))
        out[f"{name}_x_final"], out[f"{name}_logw_final"] = smc.x_saved[K], smc.logw_saved[K]
        out[f"{name}_x_first"], out[f"{name}_logw_first"] = smc.x_saved[0], smc.logw_saved[0]
        out[f"{name}_leapfrogs"] = np.array(kern.leapfrogs)
        out[f"{name}_cfg"] = np.array([N, K, eps, float(temp)])
        print(f"run {name}: {smc.run_time:.1f}s, leapfrogs/iter {kern.leapfrogs}, phi {np.round(smc.phi, 4)}")
        print("    mean_estimate[K] =", smc.mean_estimate[K])
    np.savez_compressed(OUT / "runs.npz", **out)


# ------------------------------------------------------------------ G. experiments/plot_experiments.py::mse_mean_var
def golden_mse():
    """The reference's plot_experiments.py imports seaborn/matplotlib (absent here), so the function is lifted out of the
    unmodified file by its AST and executed as is."""
    import ast
    src = Path("/root/reference/experiments/plot_experiments.py").read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "mse_mean_var")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "plot_experiments.py", "exec"), ns)
    gen = np.random.default_rng(4)
    x = gen.normal(size=(7, 16, 4)) * 0.1 + np.array([0.0, 0.95, -0.03, 0.17])
    truth = np.array([0.00678443, 0.95700831, -0.03407898, 0.16660982])
    mean, var = ns["mse_mean_var"](x, truth)
    np.savez_compressed(OUT / "mse.npz", x=x, truth=truth, mse_mean=mean, mse_var=var)
    print("mse.npz", mean[:3], var[:3])


if __name__ == "__main__":
    which = sys.argv[1:] or ["models", "nuts", "choice", "lkernel", "tempering", "runs"]
    if "models" in which:
        golden_models()
    if "nuts" in which:
        golden_nuts()
    if "nuts_devmath" in which or not sys.argv[1:]:
        golden_nuts(devmath=True)
    if "choice" in which:
        golden_choice()
    if "lkernel" in which:
        golden_lkernel_weights()
    if "tempering" in which:
        golden_tempering()
    if "runs" in which:
        golden_runs()
    if "lkernel_degenerate" in which or not sys.argv[1:]:
        golden_lkernel_degenerate()
    if "mse" in which or not sys.argv[1:]:
        golden_mse()
