"""Random programs through the Stan-subset generator: every generated value is compared with a numpy evaluation of the
same expression tree and every gradient with central finite differences.  The trees mix scalar arithmetic and functions
with container-valued sub-expressions (elementwise operations, sum / mean / dot_product / dot_self, row_vector * vector,
matrix * vector), so the lowering, hoisting and differentiation passes meet in combinations no hand-written fixture has."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import test_stan_codegen as T  # noqa: E402

HEAD = """
data { vector[3] d; row_vector[3] rw; matrix[3, 3] M; real e; real phi; }
parameters { vector[3] a; real b; real<lower=0> s; }
model {
  vector[3] loc = a .* d + b;
"""


class Gen:
    """Builds the Stan text and the numpy text of one random expression side by side."""

    def __init__(self, rng):
        self.rng = rng

    def pick(self, xs):
        return xs[int(self.rng.integers(len(xs)))]

    def scalar(self, depth):
        r = self.rng.random()
        if depth <= 0 or r < 0.15:
            return self.pick([("b", "b"), ("s", "s"), ("e", "e"), ("a[2]", "a[1]"), ("d[1]", "d[0]"), ("0.7", "0.7"), ("loc[3]", "loc[2]")])
        if r < 0.45:
            op = self.pick(["+", "-", "*"])
            (x, px), (y, py) = self.scalar(depth - 1), self.scalar(depth - 1)
            return f"({x} {op} {y})", f"({px} {op} {py})"
        if r < 0.55:
            (x, px), (y, py) = self.scalar(depth - 1), self.scalar(depth - 1)
            return f"({x} / (1.5 + square({y})))", f"({px} / (1.5 + ({py}) ** 2))"
        if r < 0.75:
            f, pf = self.pick([("tanh", "np.tanh"), ("inv_logit", "expit"), ("log1p_exp", "softplus"), ("sin", "np.sin"),
                               ("cos", "np.cos"), ("square", "np.square")])
            x, px = self.scalar(depth - 1)
            return f"{f}({x})", f"{pf}({px})"
        if r < 0.80:
            x, px = self.scalar(depth - 1)
            return f"log(1.2 + square({x}))", f"np.log(1.2 + ({px}) ** 2)"
        if r < 0.85:
            (x, px), (y, py) = self.scalar(depth - 1), self.scalar(depth - 1)
            f = self.pick(["fmin", "fmax"])
            return f"{f}({x}, {y})", f"np.{f}({px}, {py})"
        # reductions of container expressions
        kind = self.pick(["sum", "mean", "dot_product", "dot_self", "rowcol", "matvec"])
        v, pv = self.vector(depth - 1)
        if kind in ("sum", "mean"):
            return f"{kind}({v})", f"np.{kind}({pv})"
        if kind == "dot_self":
            return f"dot_self({v})", f"np.dot({pv}, {pv})"
        if kind == "dot_product":
            w, pw = self.vector(depth - 1)
            return f"dot_product({v}, {w})", f"np.dot({pv}, {pw})"
        if kind == "rowcol":
            return f"(rw * ({v}))", f"np.dot(rw, {pv})"
        return f"sum(M * ({v}))", f"np.sum(M @ ({pv}))"

    def vector(self, depth):
        r = self.rng.random()
        if depth <= 0 or r < 0.3:
            return self.pick([("a", "a"), ("d", "d"), ("loc", "loc"), ("rep_vector(0.5, 3)", "np.full(3, 0.5)")])
        if r < 0.6:
            op, pop = self.pick([("+", "+"), ("-", "-"), (".*", "*")])
            (x, px), (y, py) = self.vector(depth - 1), self.vector(depth - 1)
            return f"({x} {op} {y})", f"({px} {pop} {py})"
        if r < 0.8:
            (x, px), (y, py) = self.vector(depth - 1), self.scalar(depth - 1)
            op = self.pick(["*", "+"])
            return f"({x} {op} {y})", f"({px} {op} {py})"
        f, pf = self.pick([("tanh", "np.tanh"), ("exp", "np.exp"), ("inv_logit", "expit")])
        x, px = self.vector(depth - 1)
        return f"{f}(0.3 * {x})", f"{pf}(0.3 * {px})"


@pytest.mark.parametrize("seed", range(16))
def test_random_programs_match_numpy_and_finite_differences(tmp_path, seed):
    from scipy.special import expit
    rng = np.random.default_rng(1000 + seed)
    g = Gen(rng)
    (e1, p1), (e2, p2) = g.scalar(4), g.scalar(4)
    (v1, pv1) = g.vector(3)
    text = HEAD + f"  target += {e1};\n  target += phi * ({e2});\n  a ~ normal({v1}, 1 + s);\n}}\n"
    data = {"d": rng.normal(size=3).tolist(), "rw": rng.normal(size=3).tolist(), "M": rng.normal(size=(3, 3)).tolist(),
            "e": float(rng.normal())}
    src = T.SC.generate(text, data)
    h = T.HostModel(src, tmp_path)
    d, rw, M, e = np.array(data["d"]), np.array(data["rw"]), np.array(data["M"]), data["e"]
    env = {"np": np, "expit": expit, "softplus": lambda z: np.logaddexp(0.0, z), "d": d, "rw": rw, "M": M, "e": e}

    def restated(u):
        loc_env = dict(env, a=u[:3], b=u[3], s=np.exp(u[4]))
        loc_env["loc"] = loc_env["a"] * d + loc_env["b"]
        mu = eval(pv1, loc_env) * np.ones(3)
        sd = 1.0 + loc_env["s"]
        A = eval(p1, loc_env) + np.sum(-np.log(sd) - 0.5 * ((loc_env["a"] - mu) / sd) ** 2) + u[4]
        return float(A), float(eval(p2, loc_env))
    x = rng.normal(size=(12, 5)) * 0.5
    A, B, grad = h.split(x, 0.6)
    ref = np.array([restated(u) for u in x])
    np.testing.assert_allclose(A, ref[:, 0], rtol=1e-10, atol=1e-10, err_msg=text)
    np.testing.assert_allclose(B, ref[:, 1], rtol=1e-10, atol=1e-10, err_msg=text)
    eps = 1e-6
    for j in range(5):
        xp, xm = x.copy(), x.copy()
        xp[:, j] += eps; xm[:, j] -= eps
        Ap, Bp, _ = h.split(xp, 0.6)
        Am, Bm, _ = h.split(xm, 0.6)
        fd = ((Ap + 0.6 * Bp) - (Am + 0.6 * Bm)) / (2 * eps)
        scale = 1.0 + np.abs(fd)
        np.testing.assert_allclose(grad[:, j] / scale, fd / scale, rtol=0, atol=2e-5, err_msg=text)
