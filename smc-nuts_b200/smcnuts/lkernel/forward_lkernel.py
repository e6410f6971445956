"""ForwardLKernel (reference: smcnuts/lkernel/forward_lkernel.py): L(r_new) = momentum_proposal.logpdf(-r_new)."""
from .. import _cabi, _device as dev


class ForwardLKernel:
    def __init__(self, target, momentum_proposal):
        self.target = target
        self.momentum_proposal = momentum_proposal

    def calculate_L(self, r_new, _):
        D = self.target.dim
        if dev.is_std_normal(self.momentum_proposal, D):
            rd = dev.to_device(r_new).reshape(-1, D)      # N(0, I) is even: logpdf(-r) = logpdf(r)
            out = dev.empty(rd.shape[0])
            _cabi.call("smcb_std_normal_logpdf", dev.ptr(rd), rd.shape[0], D, dev.ptr(out), dev.stream_ptr())
            return dev.like_input(out, r_new)
        # foreign momentum plugin (e.g. a scipy frozen distribution): it gets host NumPy, whatever container came in
        return self.momentum_proposal.logpdf(-1 * dev.to_numpy(r_new))
