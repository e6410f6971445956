"""Particle sharding across GPUs: one process per GPU, contiguous global index ranges.

Rank p owns particles [p*N/P, (p+1)*N/P).  The reference is single-process; the
per-particle work (NUTS, model evaluations, reweighting, MH) needs no communication.  Coupling is only:
  * scalar / small-vector reductions (log-sum-exp triples, ESS candidates, moment sums)  -> all_gather / all_reduce
  * resampling: global exclusive scan of rank weight totals + all-to-all-v particle migration.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the transport; with no process group this
class degrades to a single shard and every collective is a no-op.
"""
import numpy as np
import torch
import torch.distributed as dist


class ShardContext:
    def __init__(self, group=None):
        self.enabled = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0

    # ---- partition helpers (pure integer logic, unit-tested on CPU)
    def local_count(self, N):
        if N % self.world:
            raise ValueError(f"N={N} must be divisible by the number of shards {self.world}")
        return N // self.world

    def offset(self, N):
        return self.rank * self.local_count(N)

    # ---- collectives
    def all_reduce_sum_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_sum_scalar(self, v):
        if self.world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device=self._dev())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def all_gather_vec(self, t):
        """[k] tensor per rank -> [P, k] tensor (rank order) on every rank."""
        if self.world == 1:
            return t.reshape(1, -1)
        out = torch.empty(self.world, t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
        return out

    def all_to_all_rows(self, send, send_counts, recv_counts):
        """all-to-all-v of row blocks: `send` [sum(send_counts), D] ordered by destination rank."""
        if self.world == 1:
            return send
        D = send.shape[1] if send.dim() > 1 else 1
        recv = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv.view(-1), send.contiguous().view(-1),
                               output_split_sizes=[int(c) * D for c in recv_counts],
                               input_split_sizes=[int(c) * D for c in send_counts], group=self.group)
        return recv

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)

    def _dev(self):
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def systematic_slot_bounds(boundaries, u0, n_total):
    """For cdf boundaries B_0 = 0 <= B_1 <= ... <= B_P (B_{q+1} = last cdf value held by rank q), return
    c[q] = number of systematic positions pos_j = (j + u0)/n_total, j = 0..n_total-1, with pos_j < B_q.
    Rank q then serves output slots [c[q], c[q+1]).  Uses the same fp64 expression for pos_j as the device
    kernel, so the split agrees bit for bit with the per-slot upper-bound search."""
    b = np.asarray(boundaries, dtype=np.float64)
    c = np.empty(len(b), dtype=np.int64)
    for q, B in enumerate(b):
        j = int(np.clip(np.ceil(B * n_total - u0), 0, n_total))
        while j > 0 and (np.float64(j - 1) + u0) / np.float64(n_total) >= B:
            j -= 1
        while j < n_total and (np.float64(j) + u0) / np.float64(n_total) < B:
            j += 1
        c[q] = j
    c[-1] = n_total  # the last boundary is the normalised total (== 1.0); every slot is served
    return c


def split_counts(lo, hi, m_per_rank, world):
    """Overlap of the slot range [lo, hi) with each destination rank's slots [d*m, (d+1)*m)."""
    return [max(0, min(hi, (d + 1) * m_per_rank) - max(lo, d * m_per_rank)) for d in range(world)]
