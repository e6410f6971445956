"""`StanModel(model_name, model_path, data_path)` with the reference's constructor
(smcnuts/model/bridgestan.py:13 in the reference).  The two shipped programs (arma, PRMwCD) resolve to their
hand-tuned fused CUDA device functions -- when the file at model_path IS the shipped program (digest of its text) or no
file is given; any other Stan program of the supported subset, an edited arma.stan included, is translated and compiled
into a device model by smcnuts/model/generated.py (`use_builtin=False` forces that path for the shipped ones too).
BridgeStan itself is not used: the device path has no CPU fallback."""
import hashlib
import json
import re
from pathlib import Path

from .device_model import DeviceModel, arma_model, prmwcd_model

# The hand-tuned device functions implement exactly the two programs the reference ships (stan_models/arma/arma.stan,
# stan_models/PRMwCD/PRMwCD.stan).  A file of that NAME with other CONTENT (an edited prior, another likelihood) must not
# silently run the built-in density: the program text, comments and white space removed, is checked against these digests.
_BUILTIN_DIGESTS = {"arma": "076309e05d1f07f7", "PRMwCD": "029bf0615d2a8571"}


def program_digest(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"(//|#)[^\n]*", "", text)
    return hashlib.sha256(re.sub(r"\s+", "", text).encode()).hexdigest()[:16]


class StanModel(DeviceModel):
    def __init__(self, model_name, model_path=None, data_path=None, use_builtin=True):
        self.model_name, self.model_path, self.data_path = model_name, model_path, data_path
        data = None
        if data_path and str(data_path) != "None" and Path(data_path).exists():
            raw = Path(data_path).read_text()
            try:
                data = json.loads(raw)
            except json.JSONDecodeError:
                data = json.loads(raw + " 1.0}")  # the shipped PRMwCD.json is truncated after `"phi": `
        if use_builtin and model_name in _BUILTIN_DIGESTS and model_path and Path(model_path).exists():
            use_builtin = program_digest(Path(model_path).read_text()) == _BUILTIN_DIGESTS[model_name]
        self.resolved = "builtin" if (use_builtin and model_name in _BUILTIN_DIGESTS) else "generated"
        if model_name == "arma" and use_builtin:
            m = arma_model(None if data is None else data["y"])
        elif model_name == "PRMwCD" and use_builtin:
            m = prmwcd_model(data)
        else:
            if not model_path or not Path(model_path).exists():
                raise FileNotFoundError(f"Stan program {model_path!r} of model {model_name!r} not found")
            from .generated import GeneratedModel
            m = GeneratedModel(Path(model_path).read_text(), data or {}, model_name)
        self._generated = not isinstance(m, DeviceModel) or type(m) is not DeviceModel
        self.__dict__.update(m.__dict__)
        m._h = None  # ownership of the handle moved to self
        self.last_phi = 1.0

    def constrain(self, x, include_tparams=True, include_gqs=True):
        if self._generated:
            from .generated import GeneratedModel
            return GeneratedModel.constrain(self, x, include_tparams, include_gqs)
        return DeviceModel.constrain(self, x, include_tparams, include_gqs)

    def _update_phi(self, phi):
        """Kept for API compatibility (bridgestan.py:122-146): phi is a kernel argument here, nothing to reload."""
        self.last_phi = phi
