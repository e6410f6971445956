"""NUTSProposalWithAccRej -- NUTS followed by the endpoint Metropolis-Hastings step, fused into the same
kernel (reference: smcnuts/proposal/nuts_acc_rej.py:27-52 + proposal/utils.py:22-34 re-evaluate the model
2N times in a second Python loop; here the log densities carried by the transition are reused)."""
from .nuts import MAX_TREE_DEPTH, NUTSProposal  # noqa: F401


class NUTSProposalWithAccRej(NUTSProposal):
    accept_reject = True

    def __init__(self, target, momentum_proposal, step_size, rng=None, max_tree_depth=MAX_TREE_DEPTH):
        super().__init__(target, momentum_proposal, step_size, rng, max_tree_depth)
