"""N(0, I) proposal with the `rvs(N)` / `logpdf(x)` interface the reference expects of
`sample_proposal` and `momentum_proposal` (scipy frozen multivariate_normal in
experiments/run_experiments.py:110-111), served by the device Philox streams."""
import math

from . import _cabi, _device as dev


class StdNormal:
    def __init__(self, dim, seed=0, stream=_cabi.STREAM_MOMENTUM):
        self.dim, self.seed, self.stream, self.calls = int(dim), int(seed), stream, 0
        self.mean = [0.0] * self.dim

    def rvs(self, N, iteration=None, particle0=0):
        it = self.calls if iteration is None else iteration
        self.calls += 1
        out = dev.empty(N, self.dim)
        _cabi.call("smcb_normals", self.seed, it, self.stream, particle0, N, self.dim, dev.ptr(out), dev.stream_ptr())
        return out

    def logpdf(self, x):
        xd = dev.to_device(x).reshape(-1, self.dim)
        out = dev.empty(xd.shape[0])
        _cabi.call("smcb_row_half_sqnorm", dev.ptr(xd), xd.shape[0], self.dim, dev.ptr(out), dev.stream_ptr())
        out = -out - 0.5 * self.dim * math.log(2.0 * math.pi)
        return dev.like_input(out, x)
