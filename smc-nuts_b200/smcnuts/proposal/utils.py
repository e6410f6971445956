"""hmc_accept_reject with the reference's signature (smcnuts/proposal/utils.py:3-34), vectorised over
particles.  The sampler does not call it -- the MH step is fused into the NUTS kernel -- it exists so
code written against the reference's helper keeps working."""
import numpy as np

from .. import _device as dev


def hmc_accept_reject(target_lpdf, x, x_prime, r, r_prime, phi=1.0, rng=None, u=None):
    """True where the move x -> x_prime is accepted.  `u`: the uniform(s) to compare against (drawn from
    `rng.uniform()` when omitted, exactly one per particle, as the reference does)."""
    xh, xph = np.atleast_2d(dev.like_input(dev.to_device(x), np.empty(0))), np.atleast_2d(
        dev.like_input(dev.to_device(x_prime), np.empty(0)))
    rh, rph = np.atleast_2d(np.asarray(r, dtype=float)), np.atleast_2d(np.asarray(r_prime, dtype=float))
    with np.errstate(all="ignore"):
        H1 = np.atleast_1d(target_lpdf(xph, phi=phi)) - 0.5 * np.sum(rph * rph, axis=1)
        H0 = np.atleast_1d(target_lpdf(xh, phi=phi)) - 0.5 * np.sum(rh * rh, axis=1)
        ratio = np.exp(H1 - H0)
        prob = np.where(ratio < 1.0, ratio, 1.0)  # python min(1., ratio): nan -> 1.
    if u is None:
        rng = rng if rng is not None else np.random.default_rng()
        u = np.array([rng.uniform() for _ in range(len(xh))])
    acc = ~((np.asarray(u) > prob) | np.any(np.isinf(xph), axis=1))
    return bool(acc[0]) if np.ndim(x) == 1 else acc
