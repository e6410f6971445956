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
    def __init__(self, group=None, enabled=True):
        """enabled=False: a single-shard context even though a process group exists (a rank running an unsharded
        control case next to the sharded job, e.g. bench.py's sharded_check)."""
        self.enabled = bool(enabled) and dist.is_available() and dist.is_initialized()
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


class _RawDeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can view memory the C library allocated."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerBuffers:
    """Double-buffered, peer-visible destination buffers for the fused particle migration.

    Every rank allocates `nbuf` buffers of [rows, D] doubles (+ [rows] int64 ancestor indices) through the C-ABI
    (cudaMalloc + CUDA IPC handle), the handles are all-gathered, and every rank maps the other ranks' buffers (NVLink
    peer access).  The resampling kernels then store migrating rows directly into the destination's buffer
    (smcb_resample_systematic_push, smcb_resample_multinomial_push).

    LIFETIME: `views[b]` is a window into buffer b, which the peers overwrite again two resamples later (double
    buffering; the completion fence of the resample in between guarantees every rank has finished with it).  The
    sampler consumes a resampled set within the iteration that produced it; callers that keep one longer must clone
    it.  `close()` unmaps and frees the buffers (Resampler.close / SMCSampler.finish call it)."""

    def __init__(self, shard, rows, D, nbuf=2):
        import ctypes
        from . import _cabi
        self.shard, self.rows, self.D, self.nbuf, self.cur = shard, rows, D, nbuf, 0
        dev_ = torch.device("cuda", torch.cuda.current_device())
        nbytes = rows * D * 8 + rows * 8
        self.local, handles = [], torch.empty(nbuf * 64, dtype=torch.uint8)
        for b in range(nbuf):
            ptr, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
            _cabi.call("smcb_peer_alloc", nbytes, ctypes.byref(ptr), ctypes.cast(h, ctypes.c_void_p))
            self.local.append(ptr.value)
            handles[b * 64:(b + 1) * 64] = torch.frombuffer(bytearray(h), dtype=torch.uint8)
        allh = torch.empty(shard.world * nbuf * 64, dtype=torch.uint8, device=dev_)
        dist.all_gather_into_tensor(allh, handles.to(dev_), group=shard.group)
        allh = allh.cpu().numpy().reshape(shard.world, nbuf, 64)
        self.opened, tables, idx_tables = [], [], []
        for b in range(nbuf):
            ptrs = []
            for q in range(shard.world):
                if q == shard.rank:
                    ptrs.append(self.local[b])
                else:
                    p = ctypes.c_void_p()
                    hb = (ctypes.c_ubyte * 64).from_buffer_copy(allh[q, b].tobytes())
                    _cabi.call("smcb_peer_open", ctypes.cast(hb, ctypes.c_void_p), ctypes.byref(p))
                    self.opened.append(p.value)
                    ptrs.append(p.value)
            tables.append(torch.tensor(ptrs, dtype=torch.int64).to(dev_))
            idx_tables.append(torch.tensor([q + rows * D * 8 for q in ptrs], dtype=torch.int64).to(dev_))
        self.tables, self.idx_tables = tables, idx_tables
        self.views = [torch.as_tensor(_RawDeviceArray(p, (rows, D)), device=dev_) for p in self.local]
        self.idx_views = [torch.as_tensor(_RawDeviceArray(p + rows * D * 8, (rows,), "<i8"), device=dev_) for p in self.local]
        self._token = torch.zeros(1, dtype=torch.float64, device=dev_)

    def next(self):
        self.cur = (self.cur + 1) % self.nbuf
        return self.cur

    def fence(self):
        """Stream-ordered completion barrier: every rank's push kernel precedes its contribution."""
        dist.all_reduce(self._token, group=self.shard.group)

    def close(self):
        from . import _cabi
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.views, self.idx_views = [], []
        for p in self.opened:
            _cabi.lib().smcb_peer_close(p)
        for p in self.local:
            _cabi.lib().smcb_peer_free(p)
        self.opened, self.local = [], []


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
