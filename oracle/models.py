"""TEST INFRASTRUCTURE ONLY (oracle).  numpy restatement of the target densities.

PARITY UNPINNED for the model arithmetic: the reference evaluates these densities through
BridgeStan (/root/reference/smcnuts/model/bridgestan.py:46,78), which is neither vendored
nor installable offline, and the reference ships no pointwise logp/grad vectors.  These
classes follow the `.stan` program text line by line and are cross-checked by
(i) central finite differences, (ii) mpmath 50-digit evaluation, (iii) posterior recovery
of the `.params` gold means through the UNMODIFIED reference sampler (tests/golden/).

Each target is duck-typed exactly as the reference expects
(/root/reference/smcnuts/model/bridgestan.py:23-25,28,60,93):
    .dim, .constrained_dim, .param_names, .logpdf(x, phi=1.0), .logpdfgrad(x, phi=1.0), .constrain(x)
and additionally exposes the split  logp(x, phi) = A(x) + phi * B(x)
(A = log prior + log Jacobian, B = log likelihood; adaptive_tempering.py:38-43 relies on it).
Failure semantics follow bridgestan.py:47-49,79-80: an invalid point gives logp = -inf and
a gradient row of -inf.
"""
import json
from math import lgamma
from pathlib import Path

import numpy as np

DATA_DIR = Path(__file__).resolve().parents[1] / "smc-nuts_b200" / "smcnuts" / "data"

LOG_2PI = float(np.log(2.0 * np.pi))
LOG_PI = float(np.log(np.pi))


class _Target:
    dim = 0

    def split(self, x):
        """x: (N, D) -> A (N,), B (N,), gradA (N, D), gradB (N, D)."""
        raise NotImplementedError

    def _eval(self, x, phi, want_grad):
        x = np.asarray(x, dtype=np.float64)
        single = x.ndim == 1
        X = x[None, :] if single else x
        with np.errstate(all="ignore"):
            A, B, gA, gB = self.split(X)
            lp = A + phi * B
            g = gA + phi * gB
        bad = ~np.isfinite(lp)
        lp = np.where(bad, -np.inf, lp)
        g = np.where(bad[:, None], -np.inf, g)
        if want_grad:
            return g[0] if single else g
        return float(lp[0]) if single else lp

    def logpdf(self, x, phi=1.0, adjust_transform=True):
        return self._eval(x, phi, False)

    def logpdfgrad(self, x, phi=1.0, adjust_transform=True):
        return self._eval(x, phi, True)


class ArmaTarget(_Target):
    """ARMA(1,1), /root/reference/stan_models/arma/arma.stan:3-31.  x = (mu, beta, theta, s), sigma = exp(s)."""

    dim = 4
    constrained_dim = 4
    param_names = ["mu", "beta", "theta", "sigma"]

    def __init__(self, y=None):
        if y is None:
            y = json.loads((DATA_DIR / "arma" / "arma.json").read_text())["y"]
        self.y = np.asarray(y, dtype=np.float64)
        self.T = len(self.y)

    def device_data(self):
        return self.y.copy()

    def split(self, X):
        mu, beta, theta, s = X[:, 0], X[:, 1], X[:, 2], X[:, 3]
        y, T = self.y, self.T
        sigma = np.exp(s)
        sig2 = sigma * sigma
        # priors (arma.stan:18-21) + Jacobian of sigma = exp(s)
        q = sig2 / 6.25
        A = (-0.5 * LOG_2PI - np.log(10.0) - mu * mu / 200.0) \
            + (-0.5 * LOG_2PI - np.log(2.0) - beta * beta / 8.0) \
            + (-0.5 * LOG_2PI - np.log(2.0) - theta * theta / 8.0) \
            + (-LOG_PI - np.log(2.5) - np.log1p(q)) + s
        gA = np.stack([-mu / 100.0, -beta / 4.0, -theta / 4.0, 1.0 - 2.0 * q / (1.0 + q)], axis=1)
        # likelihood recurrence (arma.stan:23-28) with forward sensitivities
        e = y[0] - (mu + beta * mu)
        dm, db, dt = -(1.0 + beta), -mu, np.zeros_like(mu)
        S = e * e
        Sm, Sb, St = e * dm, e * db, e * dt
        for t in range(1, T):
            e_new = y[t] - (mu + beta * y[t - 1] + theta * e)
            dm_new = -1.0 - theta * dm
            db_new = -y[t - 1] - theta * db
            dt_new = -e - theta * dt
            e, dm, db, dt = e_new, dm_new, db_new, dt_new
            S = S + e * e
            Sm, Sb, St = Sm + e * dm, Sb + e * db, St + e * dt
        inv = 1.0 / sig2
        B = -0.5 * T * LOG_2PI - T * s - 0.5 * S * inv
        gB = np.stack([-Sm * inv, -Sb * inv, -St * inv, -T + S * inv], axis=1)
        # Stan rejects a non-finite or zero scale (normal_lpdf / cauchy_lpdf argument checks)
        bad = ~np.isfinite(sigma) | (sigma <= 0.0)
        A = np.where(bad, -np.inf, A)
        return A, B, gA, gB

    def constrain(self, x, include_tparams=True, include_gqs=True):
        c = np.array(x, dtype=np.float64, copy=True)
        c[..., 3] = np.exp(c[..., 3])
        return c


class PRMwCDTarget(_Target):
    """Poisson regression with exponential-power prior, /root/reference/stan_models/PRMwCD/PRMwCD.stan:1-39.

    x = (Beta_1..Beta_12, g), Gamma = exp(g).
    """

    dim = 13
    constrained_dim = 13
    param_names = [f"Beta.{i}" for i in range(1, 13)] + ["Gamma"]

    def __init__(self, data=None):
        if data is None:
            data = json.loads((DATA_DIR / "PRMwCD" / "PRMwCD.json").read_text())
        self.Nobs, self.M, self.q, self.C = int(data["N"]), int(data["M"]), float(data["q"]), int(data["Clength"])
        self.y = np.asarray(data["y"], dtype=np.float64)
        self.X = np.asarray(data["Xkernel"], dtype=np.float64).reshape(self.Nobs, self.C)  # PRMwCD.stan:28
        self.lgam = np.array([lgamma(v + 1.0) for v in self.y])
        assert self.M == self.C + 1

    def device_data(self):
        """[q, y(100), lgamma(y+1)(100), X(100x11 row-major)]"""
        return np.concatenate([[self.q], self.y, self.lgam, self.X.ravel()])

    def split(self, X):
        Bt, g = X[:, : self.M], X[:, self.M]
        q = self.q
        eta = Bt[:, :1] + Bt[:, 1:] @ self.X.T                      # (N, Nobs)  PRMwCD.stan:24-30
        lam = np.exp(eta)
        term = self.y[None, :] * eta - lam - self.lgam[None, :]     # poisson_lpmf, PRMwCD.stan:32
        # Stan: lambda == inf -> -inf ; lambda == 0 with y != 0 -> -inf
        term = np.where((lam == 0.0) & (self.y[None, :] > 0), -np.inf, term)
        B = term.sum(axis=1)
        d = self.y[None, :] - lam
        gB = np.zeros_like(X)
        gB[:, 0] = d.sum(axis=1)
        gB[:, 1: self.M] = d @ self.X
        # priors: inv_gamma(Gamma | 2, 1.3) (PRMwCD.stan:21), EP prior on Beta[2:M] (:36-38), Jacobian g
        inv_gam = np.exp(-g)
        a = np.abs(Bt[:, 1:]) * inv_gam[:, None]                      # |Beta_i / Gamma|
        aq = a ** q
        A = (2.0 * np.log(1.3) - lgamma(2.0) - 3.0 * g - 1.3 * inv_gam) + g \
            + (-(self.M - 1) * g - aq.sum(axis=1))
        gA = np.zeros_like(X)
        gA[:, 1: self.M] = -q * aq / Bt[:, 1:]
        gA[:, self.M] = -3.0 + 1.3 * inv_gam + 1.0 - (self.M - 1) + q * aq.sum(axis=1)
        Gam = np.exp(g)
        bad = ~np.isfinite(Gam) | (Gam <= 0.0)
        A = np.where(bad, -np.inf, A)
        return A, B, gA, gB

    def constrain(self, x, include_tparams=True, include_gqs=True):
        c = np.array(x, dtype=np.float64, copy=True)
        c[..., self.M] = np.exp(c[..., self.M])
        return c


class GaussTarget(_Target):
    """Synthetic correlated Gaussian (BASELINE.json config 4; defined in SURVEY.md §8d, not in the reference).

    Sigma_ij = rho^|i-j|, P = Sigma^-1 stored dense; A = 0, B = -0.5 x'Px, grad B = -P x.  No
    `constrained_dim`, so Estimate takes the unconstrained branch (estimate.py:25-28).
    """

    def __init__(self, dim=100, rho=0.9, precision=None):
        if precision is not None:          # any symmetric positive definite precision matrix
            P = np.asarray(precision, dtype=np.float64)
            self.dim = P.shape[0]
            self.P = 0.5 * (P + P.T)
            self.Sigma = np.linalg.inv(self.P)
            return
        self.dim = dim
        idx = np.arange(dim)
        self.Sigma = rho ** np.abs(idx[:, None] - idx[None, :])
        P = np.linalg.inv(self.Sigma)
        self.P = 0.5 * (P + P.T)

    def device_data(self):
        return self.P.ravel().copy()

    def split(self, X):
        Px = X @ self.P
        B = -0.5 * np.einsum("nd,nd->n", X, Px)
        return np.zeros(len(X)), B, np.zeros_like(X), -Px


def make_target(name, **kw):
    return {"arma": ArmaTarget, "PRMwCD": PRMwCDTarget, "gauss": GaussTarget}[name](**kw)
