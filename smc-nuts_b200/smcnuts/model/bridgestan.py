"""`StanModel(model_name, model_path, data_path)` with the reference's constructor
(smcnuts/model/bridgestan.py:13 in the reference), resolved to the fused CUDA device function of the
same name.  BridgeStan itself is not used: the device path has no CPU fallback, and only the shipped
Stan programs (arma, PRMwCD) have device functions."""
import json
from pathlib import Path

from .device_model import DeviceModel, arma_model, prmwcd_model


class StanModel(DeviceModel):
    def __init__(self, model_name, model_path=None, data_path=None):
        self.model_name, self.model_path, self.data_path = model_name, model_path, data_path
        data = None
        if data_path and str(data_path) != "None" and Path(data_path).exists():
            raw = Path(data_path).read_text()
            try:
                data = json.loads(raw)
            except json.JSONDecodeError:
                data = json.loads(raw + " 1.0}")  # the shipped PRMwCD.json is truncated after `"phi": `
        if model_name == "arma":
            m = arma_model(None if data is None else data["y"])
        elif model_name == "PRMwCD":
            m = prmwcd_model(data)
        else:
            raise NotImplementedError(f"no device function for Stan model {model_name!r}; available: arma, PRMwCD")
        self.__dict__.update(m.__dict__)
        m._h = None  # ownership of the handle moved to self
        self.last_phi = 1.0

    def _update_phi(self, phi):
        """Kept for API compatibility (bridgestan.py:122-146): phi is a kernel argument here, nothing to reload."""
        self.last_phi = phi
