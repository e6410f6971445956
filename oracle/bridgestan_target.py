"""TEST INFRASTRUCTURE -- optional BridgeStan-backed CPU target ("BridgeStan-backed CPU target as
optional oracle when the package exists").

When the `bridgestan` package (and with it stanc + Stan Math + a C++ toolchain) is importable, this evaluates a Stan
program exactly the way the reference does (smcnuts/model/bridgestan.py:18,46,78 of the reference:
`bs.StanModel.from_stan_file(model_path, data_path)`, `log_density`, `log_density_gradient`, Jacobian included), which
pins the model arithmetic of the oracle (oracle/models.py, oracle/smc_oracle.c) and of the generated device models
against Stan itself.  The package is NOT installable in the offline build image, so tests/test_bridgestan_pin.py skips
there and DESIGN.md keeps saying "parity unpinned against BridgeStan"; nothing in the product imports this module.

Tempering follows the reference: the data variable `phi` is rewritten in a private copy of the data file and the model
reloaded (bridgestan.py:122-146 of the reference rewrites the file in place); A = logp(phi = 0), B = logp(1) - logp(0).
"""
import json
import tempfile
from pathlib import Path

import numpy as np


def available():
    try:
        import bridgestan  # noqa: F401
        return True
    except Exception:
        return False


class BridgeStanTarget:
    def __init__(self, stan_path, data=None):
        import bridgestan as bs
        self._bs = bs
        self.stan_path = str(stan_path)
        self.data = dict(data or {})
        self._dir = Path(tempfile.mkdtemp(prefix="smcb_bs_"))
        self._models = {}
        self.dim = self._model(1.0).param_unc_num()

    def _model(self, phi):
        phi = float(phi)
        if phi not in self._models:
            data = dict(self.data)
            if "phi" in data or self._declares_phi():
                data["phi"] = phi
            path = self._dir / f"data_{len(self._models)}.json"
            path.write_text(json.dumps(data))
            make = getattr(self._bs.StanModel, "from_stan_file", None)
            self._models[phi] = make(self.stan_path, str(path)) if make else self._bs.StanModel(self.stan_path, str(path))
        return self._models[phi]

    def _declares_phi(self):
        import re
        text = Path(self.stan_path).read_text()
        m = re.search(r"data\s*\{(.*?)\}", text, re.S)
        return bool(m and re.search(r"\bphi\b", m.group(1)))

    def logpdf(self, x, phi=1.0):
        """log density on the unconstrained scale, Jacobian included; a Stan exception maps to -inf like the reference"""
        m = self._model(phi)
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        out = np.empty(len(x))
        for i, row in enumerate(x):
            try:
                out[i] = m.log_density(row)
            except Exception:
                out[i] = -np.inf
        return out

    def logpdfgrad(self, x, phi=1.0):
        m = self._model(phi)
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        out = np.empty_like(x)
        for i, row in enumerate(x):
            try:
                out[i] = m.log_density_gradient(row)[1]
            except Exception:
                out[i] = -np.inf
        return out

    def split(self, x):
        """(A, B) with logp(x, phi) = A + phi * B"""
        a = self.logpdf(x, 0.0)
        return a, self.logpdf(x, 1.0) - a

    def constrain(self, x):
        m = self._model(1.0)
        return np.array([m.param_constrain(row) for row in np.atleast_2d(np.asarray(x, dtype=np.float64))])
