"""CPU: the C-ABI shared library loads and exports every symbol include/smcnuts_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import subprocess

import pytest

from smcnuts import _cabi


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(str(_cabi.LIB_PATH))
    declared = _cabi.declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/smcnuts_b200.h but not exported"


def test_python_binding_covers_the_header():
    assert sorted(_cabi._SIGS) == _cabi.declared_symbols()
    L = _cabi.lib()
    assert L.smcb_version() == 100
    assert L.smcb_reduce_workspace_bytes() > 0 and L.smcb_scan_workspace_bytes(1 << 20) >= (1 << 20) // 2048 * 8


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_argument_errors_are_reported_not_thrown():
    L = _cabi.lib()
    assert L.smcb_combine_logp(None, None, 1.0, 4, None, None) != 0
    assert b"bad argument" in L.smcb_last_error()
    with pytest.raises(_cabi.SmcbError):
        _cabi.call("smcb_lse_finalize", None, 0, None, None)


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from smcnuts.model.device_model import arma_model
    with pytest.raises(_cabi.SmcbError):
        arma_model()
