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


class LoopGen:
    """Random recurrence programs: two model-block vectors filled by pre-loop writes and one loop (each element written
    exactly once, unconditionally or in both branches of an if), read at [t - 1], [1] and -- after this trip's write --
    [t]; consumed by vectorised densities after the loop.  Only defined elements are ever read."""

    def __init__(self, rng):
        self.rng = rng

    def pick(self, xs):
        return xs[int(self.rng.integers(len(xs)))]

    def expr(self, readable, depth=2):
        r = self.rng.random()
        if depth <= 0 or r < 0.35:
            return self.pick(readable + ["a", "b", "y[t]", "0.3"])
        if r < 0.75:
            return f"({self.expr(readable, depth - 1)} {self.pick(['+', '-', '*'])} {self.expr(readable, depth - 1)})"
        return f"{self.pick(['tanh', 'sin', 'inv_logit'])}({self.expr(readable, depth - 1)})"

    def program(self):
        rng = self.rng
        lines = ["vector[T] v;", "vector[T] w;"]
        pre = ["a", "b", "y[1]", "0.3"]
        lines.append("v[1] = " + self.pick(pre) + " * 0.5;")
        lines.append("w[1] = " + self.pick(pre + ["v[1]"]) + " - 0.2;")
        if rng.random() < 0.3:
            lines.append("target += -0.01 * square(v[1] - w[1]);")
        body, written = [], set()
        order = ["v", "w"] if rng.random() < 0.5 else ["w", "v"]
        slots = [order[0], "use", order[1], "use"]
        for what in slots:
            readable = ["v[t - 1]", "w[t - 1]", "v[1]"] + [f"{x}[t]" for x in written]
            readable = [r.replace("y[t]", "y[t]") for r in readable]
            if what == "use":
                if rng.random() < 0.6:
                    body.append(f"target += -0.01 * square({self.expr(readable)});")
                continue
            rhs = f"0.5 * {self.expr(readable)}"
            if rng.random() < 0.3:
                body.append(f"if (t <= 3) {what}[t] = {rhs}; else {what}[t] = 0.4 * {self.expr(readable)};")
            else:
                body.append(f"{what}[t] = {rhs};")
            written.add(what)
        lines.append("for (t in 2:T) {")
        lines += ["  " + b for b in body]
        lines.append("}")
        for vec in ("v", "w"):
            r = rng.random()
            if r < 0.4:
                lines.append(f"target += phi * normal_lpdf({vec} | 0, 1 + s);")
            elif r < 0.6:
                lines.append(f"{vec} ~ normal(y, 1 + s);")
            elif r < 0.8:
                lines.append(f"target += -0.1 * dot_self({vec});")
            else:
                lines.append(f"target += -0.1 * square({vec}[T]);")
        if rng.random() < 0.3:
            lines.append("target += normal_lpdf(v | w, 2);")      # two model-block vectors: never fused
        return ("data { int T; vector[T] y; real phi; }\nparameters { real a; real b; real<lower=0> s; }\nmodel {\n  "
                + "\n  ".join(lines) + "\n}\n")


@pytest.mark.parametrize("seed", range(12))
def test_restructured_recurrence_programs_equal_their_plain_form(tmp_path, seed):
    """Differential test of the statement-level passes: every random recurrence program is generated with and without the
    restructuring (fusion of trailing vectorised densities, arrays reduced to rolling scalars); values and gradients of
    both builds must agree."""
    import re
    rng = np.random.default_rng(5000 + seed)
    text = LoopGen(rng).program()
    data = {"T": 6, "y": rng.normal(size=6).tolist()}
    x = rng.normal(size=(16, 3)) * 0.5
    out = {}
    for flag in (True, False):
        T.SC.RESTRUCTURE = flag
        try:
            src = T.SC.generate(text, data)
        finally:
            T.SC.RESTRUCTURE = True
        out[flag] = (T.HostModel(src, tmp_path / str(flag)).split(x, 0.7), len(re.findall(r"double v_[vw]\[\d+\]", src.text)))
    assert out[False][1] == 2, text
    for got, want in zip(out[True][0], out[False][0]):
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, err_msg=text)


def test_the_restructuring_fires_on_the_random_recurrence_programs():
    """The differential test above is only worth something if the passes do change programs: count them (generation
    only, nothing is compiled)."""
    import re
    fired = 0
    for seed in range(12):
        rng = np.random.default_rng(5000 + seed)
        text = LoopGen(rng).program()
        src = T.SC.generate(text, {"T": 6, "y": rng.normal(size=6).tolist()})
        fired += len(re.findall(r"double v_[vw]\[\d+\]", src.text)) < 2
    assert fired >= 3, fired
