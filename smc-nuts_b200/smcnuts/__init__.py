"""smcnuts -- B200-native SMC-NUTS particle hot path behind the reference's plugin API.

Module paths and class names mirror the reference checkout (smcnuts/smc_sampler.py, samples/samples.py,
proposal/nuts.py, proposal/nuts_acc_rej.py, lkernel/*.py, tempering/adaptive_tempering.py,
estimate/*.py, model/bridgestan.py).  All arithmetic runs in hand-written sm_100a CUDA behind the C-ABI
of include/smcnuts_b200.h (loaded by ctypes in `_cabi`); there is no CPU fallback: importing the plugin
classes works anywhere, but the first compute call raises if the shared library or a CUDA device is
missing.
"""
__all__ = ["SMCSampler"]


def __getattr__(name):
    if name == "SMCSampler":
        from .smc_sampler import SMCSampler
        return SMCSampler
    raise AttributeError(name)
