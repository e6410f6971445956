"""Samples -- the device-resident particle set and its per-iteration operations.

Mirrors smcnuts/samples/samples.py of the reference (same constructor, same method names, same
attributes), with every array a float64 CUDA tensor in the reference's row-major [N, D] layout and every
operation a kernel behind include/smcnuts_b200.h.  When a torch.distributed process group exists the
particle set is sharded (rank p owns [p*N/P, (p+1)*N/P)); `N` is always the GLOBAL particle count.
"""

import numpy as np
import torch

from .. import _cabi, _device as dev
from ..lkernel.forward_lkernel import ForwardLKernel
from ..lkernel.gaussian_lkernel import GaussianApproxLKernel
from ..parallel import ShardContext, split_counts, systematic_slot_bounds
from ..tempering.adaptive_tempering import ESSTempering


class ScanState:
    """First half of the cdf scan, produced together with the normalised weights (smcb_normalise_tilesums): the scanned
    tile sums in `ws` and this rank's weight total.  Valid until the next normalisation."""

    def __init__(self, ws, total):
        self.ws, self.total = ws, total


def normalise(logw, shard, scan=False):
    """samples.py:91-113 -> (wn, stats, tris[, ScanState]) where stats is a device tensor [logZ, ESS] (global when
    sharded).  scan=True fuses the first pass of the resampling scan into the normalisation (the weights are read once
    for both)."""
    n = logw.shape[0]
    st = dev.stream_ptr()
    tri = dev.empty(3)
    _cabi.call("smcb_lse_partial", dev.ptr(logw), n, dev.ptr(tri), dev.ptr(dev.reduce_ws()), st)
    tris = shard.all_gather_vec(tri).contiguous()
    stats = dev.empty(2)
    _cabi.call("smcb_lse_finalize", dev.ptr(tris), shard.world, dev.ptr(stats), st)
    wn = dev.empty(n)
    if scan:
        ws = dev.workspace("scan", _cabi.lib().smcb_scan_workspace_bytes(n))
        total = dev.empty(1)
        _cabi.call("smcb_normalise_tilesums", dev.ptr(logw), n, dev.ptr(stats), dev.ptr(wn), dev.ptr(total), dev.ptr(ws), st)
        return wn, stats, tris, ScanState(ws, total)
    _cabi.call("smcb_normalise", dev.ptr(logw), n, dev.ptr(stats), dev.ptr(wn), st)
    return wn, stats, tris


class Resampler:
    """Ancestor selection + particle gather (samples.py:125-146).

    scheme "multinomial": numpy `rng.choice` semantics (cdf + upper-bound search of one uniform per slot).
    scheme "systematic" : positions (j + u0)/N (north_star; not in the reference).
    Sharded: global exclusive scan of the rank weight totals, then all-to-all-v migration of particle rows.
    """

    def __init__(self, N, seed, shard, stream=_cabi.STREAM_RESAMPLE, scheme="multinomial"):
        if scheme not in ("multinomial", "systematic"):
            raise ValueError("resampling scheme must be 'multinomial' or 'systematic'")
        self.N, self.seed, self.shard, self.stream, self.scheme = N, seed, shard, stream, scheme
        self.n_local = shard.local_count(N)
        self.offset = shard.offset(N)
        self.peers = None             # PeerBuffers for the fused NVLink migration (sharded systematic resampling)
        self.use_peer_push = True
        self.keep_idx = True          # materialise the ancestor indices (diagnostic; bench.py turns it off)
        self.last_idx = None          # ancestors of the last resample (global indices for this rank's slots)
        self._migrated = 0            # rows this rank received from other ranks in the last resample (diagnostic)

    def _cdf(self, wn, scan=None):
        n, st, sh = wn.shape[0], dev.stream_ptr(), self.shard
        if scan is not None:
            # the tile sums were produced with the normalised weights: only the second pass of the scan is left
            cdf = dev.empty(n)
            off = 0
            if sh.world > 1:
                totals = sh.all_gather_vec(scan.total).view(-1).contiguous()
                off_t = dev.empty(2)
                _cabi.call("smcb_rank_offsets", dev.ptr(totals), sh.world, sh.rank, dev.ptr(off_t), st)
                off = dev.ptr(off_t)
            _cabi.call("smcb_cdf_from_tilesums", dev.ptr(wn), n, off, dev.ptr(cdf), dev.ptr(scan.ws), st)
            return cdf
        ws = dev.workspace("scan", _cabi.lib().smcb_scan_workspace_bytes(n))
        cdf, total = dev.empty(n), dev.empty(1)
        if sh.world == 1:
            _cabi.call("smcb_cdf", dev.ptr(wn), n, 0, 0, dev.ptr(cdf), dev.ptr(total), dev.ptr(ws), st)
            return cdf
        # rank totals -> exclusive offsets (sequential fp64 sum in rank order, identical on every rank), all on the device
        tmp = dev.empty(n)
        _cabi.call("smcb_cdf", dev.ptr(wn), n, 0, 0, dev.ptr(tmp), dev.ptr(total), dev.ptr(ws), st)
        totals = sh.all_gather_vec(total).view(-1).contiguous()
        off_t = dev.empty(2)
        _cabi.call("smcb_rank_offsets", dev.ptr(totals), sh.world, sh.rank, dev.ptr(off_t), st)
        _cabi.call("smcb_cdf", dev.ptr(wn), n, off_t.data_ptr(), off_t.data_ptr() + 8, dev.ptr(cdf), dev.ptr(total),
                   dev.ptr(ws), st)
        return cdf

    def resample_rows(self, x, wn, iteration, scan=None):
        """Returns the resampled rows for this rank's output slots.  scan: the ScanState that came with `wn`."""
        return self.resample_from_cdf(x, self._cdf(wn, scan), iteration)

    def resample_from_cdf(self, x, cdf, iteration):
        sh, st = self.shard, dev.stream_ptr()
        n, D = x.shape
        if sh.world == 1:
            idx = dev.empty(n, dtype=torch.int64)
            if self.scheme == "multinomial":
                u = dev.empty(n)
                _cabi.call("smcb_uniforms", self.seed, iteration, self.stream, 0, n, 0, dev.ptr(u), st)
                _cabi.call("smcb_ancestors_multinomial", dev.ptr(cdf), n, dev.ptr(u), n, dev.ptr(idx), st)
            else:   # ancestors + gather in one kernel; idx is kept only as a diagnostic
                out = dev.empty(n, D)
                keep = idx if (self.keep_idx or D not in (2, 4, 8, 16, 32, 64)) else None
                ws = dev.workspace("resample", _cabi.lib().smcb_resample_workspace_bytes(n, D))
                fused = D in (2, 4, 8, 16, 32, 64)
                u0_dev = self._u0_dev(iteration) if fused else None       # stays on the device: no host round trip
                _cabi.call("smcb_resample_systematic", dev.ptr(cdf), n, 0.0 if fused else self._u0(iteration),
                           dev.ptr(u0_dev), 0, n, n, dev.ptr(x), D, dev.ptr(out), dev.ptr(keep), dev.ptr(ws), st)
                self.last_idx = keep
                return out
            out = dev.empty(n, D)
            _cabi.call("smcb_gather_rows", dev.ptr(x), dev.ptr(idx), n, D, dev.ptr(out), st)
            self.last_idx = idx
            return out
        if self.scheme == "systematic":
            return self._systematic_sharded(x, cdf, iteration)
        return self._multinomial_sharded(x, cdf, iteration)

    def _u0_dev(self, iteration):
        u = dev.empty(1)
        _cabi.call("smcb_uniforms", self.seed, iteration, self.stream, 0, 1, 0, dev.ptr(u), dev.stream_ptr())
        return u

    def _u0(self, iteration):
        u = dev.empty(1)
        _cabi.call("smcb_uniforms", self.seed, iteration, self.stream, 0, 1, 0, dev.ptr(u), dev.stream_ptr())
        return float(u.item())

    def _systematic_sharded(self, x, cdf, iteration):
        """Sorted positions -> every source rank serves one contiguous slot range; rows travel by all-to-all-v."""
        sh, st = self.shard, dev.stream_ptr()
        n, D = x.shape
        u0 = self._u0(iteration)
        lasts = sh.all_gather_vec(cdf[-1:].contiguous()).view(-1).cpu().numpy()
        bounds = systematic_slot_bounds(np.concatenate([[0.0], lasts]), u0, self.N)
        lo, hi = int(bounds[sh.rank]), int(bounds[sh.rank + 1])
        m = hi - lo
        mylo, myhi = sh.rank * self.n_local, (sh.rank + 1) * self.n_local
        recv_counts = [max(0, min(myhi, int(bounds[q + 1])) - max(mylo, int(bounds[q]))) for q in range(sh.world)]
        self._migrated = sum(c for q, c in enumerate(recv_counts) if q != sh.rank)
        if self.use_peer_push and D in (2, 4, 8, 16, 32, 64):
            # fused path: gather + migration in one kernel, rows stored straight into the destination GPU over NVLink
            self._peer_buffers(D)
            b = self.peers.next()
            keep = dev.empty(max(m, 1), dtype=torch.int64) if self.keep_idx else None
            if m:
                ws = dev.workspace("resample", _cabi.lib().smcb_resample_workspace_bytes(m, D))
                _cabi.call("smcb_resample_systematic_push", dev.ptr(cdf), n, u0, lo, self.N, m, dev.ptr(x), D,
                           dev.ptr(self.peers.tables[b]), self.n_local, dev.ptr(keep), dev.ptr(ws), st)
            self.peers.fence()
            self.last_idx = keep[:m] + self.offset if keep is not None else None
            return self.peers.views[b]
        idx = dev.empty(max(m, 1), dtype=torch.int64)
        send = dev.empty(max(m, 1), D)
        if m:
            ws = dev.workspace("resample", _cabi.lib().smcb_resample_workspace_bytes(m, D))
            _cabi.call("smcb_resample_systematic", dev.ptr(cdf), n, u0, 0, lo, self.N, m, dev.ptr(x), D, dev.ptr(send),
                       dev.ptr(idx), dev.ptr(ws), st)
        send_counts = split_counts(lo, hi, self.n_local, sh.world)
        self.last_idx = idx[:m] + self.offset
        return sh.all_to_all_rows(send[:m], send_counts, recv_counts)

    def _peer_buffers(self, D):
        if self.peers is None or self.peers.D != D:
            from ..parallel import PeerBuffers
            if self.peers is not None:
                self.peers.close()
            self.peers = PeerBuffers(self.shard, self.n_local, D)
        return self.peers

    def _multinomial_sharded(self, x, cdf, iteration):
        """Owner-push: every rank regenerates the N uniforms of the global slots from the Philox stream, resolves the ones
        that fall into its own cdf segment and stores the ancestors' rows straight into the slot owners' buffers over
        NVLink (smcb_resample_multinomial_push).  One tiny all-gather (the P segment ends) and the completion fence are
        the only collectives; nothing touches the host."""
        sh, st = self.shard, dev.stream_ptr()
        n, D = x.shape
        ends = sh.all_gather_vec(cdf[-1:].contiguous()).view(-1).contiguous()
        peers = self._peer_buffers(D)
        b = peers.next()
        _cabi.call("smcb_resample_multinomial_push", dev.ptr(cdf), n, dev.ptr(ends), sh.rank, sh.world, self.seed,
                   iteration, self.stream, self.N, self.offset, dev.ptr(x), D, dev.ptr(peers.tables[b]),
                   dev.ptr(peers.idx_tables[b]) if self.keep_idx else 0, self.n_local, st)
        peers.fence()
        self.last_idx = peers.idx_views[b] if self.keep_idx else None
        self._migrated = None         # counted lazily from last_idx (no host synchronisation on the hot path)
        return peers.views[b]

    @property
    def last_migrated_rows(self):
        if self._migrated is None:
            if self.last_idx is None:
                return -1
            i = self.last_idx
            self._migrated = int(((i < self.offset) | (i >= self.offset + self.n_local)).sum().item())
        return self._migrated

    def close(self):
        if self.peers is not None:
            self.peers.close()
            self.peers = None


class Samples:
    def __init__(self, N, D, sample_proposal, target, forward_kernel, lkernel, tempering, rng,
                 resampling="multinomial", shard: ShardContext = None) -> None:
        """
        N: GLOBAL number of samples; D: dimension; sample_proposal: q0; target: device model;
        forward_kernel: the NUTS proposal plugin; lkernel: "GaussianApproxLKernel" | "forwardsLKernel" |
        "asymptoticLKernel" (or an L-kernel plugin instance); tempering: bool; rng: seed source.
        """
        self.N = N
        self.D = D
        self.sample_proposal = sample_proposal
        self.forward_kernel = forward_kernel
        self.target = target
        self.rng = rng
        self.shard = shard or ShardContext()
        self.n_local = self.shard.local_count(N)
        self.offset = self.shard.offset(N)
        self.seed = forward_kernel.seed if hasattr(forward_kernel, "seed") else dev.seed_from_rng(rng)
        self.iteration = 0
        self.resampler = Resampler(N, self.seed, self.shard, scheme=resampling)
        self.resampled_last = False

        # L-kernel selection by string, as the reference (samples.py:39-48); plugin instances are accepted too
        if lkernel == "GaussianApproxLKernel":
            self.lkernel = GaussianApproxLKernel(target=self.target, N=self.N, shard=self.shard)
            self.reweight_strategy = self._non_asympototic_reweight
        elif lkernel == "forwardsLKernel":
            mp = self.forward_kernel.momentum_proposal if self.forward_kernel is not None else None
            self.lkernel = ForwardLKernel(target=self.target, momentum_proposal=mp)
            self.reweight_strategy = self._non_asympototic_reweight
        elif lkernel == "asymptoticLKernel":
            self.lkernel = None
            self.reweight_strategy = self._asymptotic_reweight
        elif hasattr(lkernel, "calculate_L"):
            self.lkernel = lkernel
            self.reweight_strategy = self._non_asympototic_reweight
        else:
            raise Exception("Unknown L-kernel supplied")

        if tempering:
            self.TemperingScheme = ESSTempering(self.N, self.target, alpha=0.5, shard=self.shard)
            self.update_temperature = self._tempering
            self.phi_old = 0.0
            self.phi_new = 0.0
        else:
            self.TemperingScheme = None
            self.update_temperature = lambda: 1.0
            self.phi_old = 1.0
            self.phi_new = 1.0
        self._stats = None
        self._scan = None
        self._carry = None       # (A, B, grad) at the current x, handed back to the NUTS kernel (skips its initial evaluation)
        self.carry_gradients = False
        self._split_new = None   # (A, B) at x_new from the last transition
        self._split_x = None     # (A, B) at x (pre-move), from the last transition
        self._ke = None

    # ------------------------------------------------------------------ helpers
    def _draw_std_normal(self, dist, stream, iteration):
        if dev.is_std_normal(dist, self.D):
            out = dev.empty(self.n_local, self.D)
            _cabi.call("smcb_normals", self.seed, iteration, stream, self.offset, self.n_local, self.D, dev.ptr(out),
                       dev.stream_ptr())
            return out
        z = np.asarray(dist.rvs(self.N)).reshape(self.N, self.D)   # generic plugin: host draw, then upload
        return dev.to_device(z[self.offset:self.offset + self.n_local])

    def _logpdf_q0(self, x):
        if dev.is_std_normal(self.sample_proposal, self.D):
            out = dev.empty(x.shape[0])
            _cabi.call("smcb_std_normal_logpdf", dev.ptr(x), x.shape[0], self.D, dev.ptr(out), dev.stream_ptr())
            return out
        return dev.to_device(self.sample_proposal.logpdf(x.cpu().numpy()))

    # ------------------------------------------------------------------ reference API
    def initialise_samples(self):
        """samples.py:63-88."""
        self.x = self._draw_std_normal(self.sample_proposal, _cabi.STREAM_INIT, 0)
        self.x_new = self.x
        self.ess = 0
        self.r = dev.zeros(self.n_local, self.D)
        self.r_new = dev.zeros(self.n_local, self.D)
        A, B = self.target.split(self.x)
        self._split_new = (A, B)
        self.phi_new = self.update_temperature()
        self.phi_old = self.phi_new
        lp = self.target.combine(A, B, self.phi_new)
        if dev.is_std_normal(self.sample_proposal, self.D):
            self.logw = dev.empty(self.n_local)
            _cabi.call("smcb_init_logw", dev.ptr(lp), dev.ptr(self.x), self.n_local, self.D, dev.ptr(self.logw),
                       dev.stream_ptr())
        else:
            self.logw = lp - self._logpdf_q0(self.x)
        self.logw_new = dev.zeros(self.n_local)
        self.wn = dev.zeros(self.n_local)

    def normalise_weights(self):
        """samples.py:91-105 (+ the sums calculate_ess needs, from the same pass)."""
        self.wn, self._stats, _, self._scan = normalise(self.logw, self.shard, scan=True)
        self._stats_host = None

    def _host_stats(self):
        if self._stats_host is None:
            self._stats_host = self._stats.cpu().numpy()
        return self._stats_host

    @property
    def log_likelihood(self):
        return float(self._host_stats()[0])

    def calculate_ess(self):
        """samples.py:108-113: 1 / sum(wn^2) = (sum e)^2 / sum e^2 of the same online pass."""
        self.ess = float(self._host_stats()[1])

    def resample_if_required(self):
        """samples.py:116-122 (threshold hard coded to 1/2)."""
        self.resampled_last = False
        if self.ess < self.N / 2:
            self._resample(self.x, self.wn, self.log_likelihood)
            self.resampled_last = True

    def _resample(self, x, wn, log_likelihood):
        """samples.py:125-146."""
        scan = self._scan if wn is self.wn else None
        self.x = self.resampler.resample_rows(x, wn, self.iteration, scan)
        self._carry = None   # the carried evaluation belongs to the pre-resampling rows
        self.logw = dev.empty(self.n_local)
        _cabi.call("smcb_uniform_logw", dev.ptr(self._stats), self.N, self.n_local, dev.ptr(self.logw), dev.stream_ptr())

    def prefetch_momentum(self):
        """Enqueue this iteration's momentum draw (samples.py:155) ahead of the host synchronisation on ESS: it depends
        only on (seed, iteration), and having it queued shortens the GPU idle gap between the resampling decision and
        the NUTS launch."""
        self._r_ready = (self.iteration, self._draw_std_normal(self.forward_kernel.momentum_proposal,
                                                               _cabi.STREAM_MOMENTUM, self.iteration))
        if hasattr(self.forward_kernel, "prepare"):      # the transition's output buffers and scratch, for the same reason
            can_carry = self.carry_gradients and self.TemperingScheme is None and \
                not getattr(self.forward_kernel, "accept_reject", False)
            self.forward_kernel.prepare(self.n_local, self.D, want_grad=can_carry)

    def propose_samples(self):
        """samples.py:149-158."""
        fk = self.forward_kernel
        ready = getattr(self, "_r_ready", None)
        if ready is not None and ready[0] == self.iteration:
            self.r, self._r_ready = ready[1], None
        else:
            self.r = self._draw_std_normal(fk.momentum_proposal, _cabi.STREAM_MOMENTUM, self.iteration)
        if hasattr(fk, "transition"):
            fk.particle0 = self.offset
            # constant temperature and no MH epilogue: the evaluation at x_new is the next iteration's evaluation at x.
            # MEASURED on B200 (arma, N = 2^20): skipping the 7 % initial evaluations this way costs more in candidate
            # gradient traffic than it saves (3.48 ms vs 3.17 ms per transition), so it is opt-in.
            can_carry = self.carry_gradients and self.TemperingScheme is None and not getattr(fk, "accept_reject", False)
            o = fk.transition(self.x, self.r, self.phi_new, iteration=self.iteration,
                              carry=self._carry if can_carry else None, want_grad=can_carry)
            self._carry = (o["A_new"], o["B_new"], o["g_new"]) if can_carry else None
            self.x_new, self.r_new = o["x_new"], o["r_new"]
            self._split_x, self._split_new = (o["A_old"], o["B_old"]), (o["A_new"], o["B_new"])
            self._ke = (o["ke_old"], o["ke_new"])
            self.n_leapfrog = o["n_leapfrog"]
        else:  # foreign proposal plugin: only the reference contract is available
            self.x_new, self.r_new = (dev.to_device(t) for t in fk.rvs(self.x, self.r, phi=self.phi_new))
            self._split_x, self._split_new, self._ke = self.target.split(self.x), self.target.split(self.x_new), None

    def reweight(self):
        self.logw_new = self.reweight_strategy()

    def _asymptotic_reweight(self):
        """samples.py:169-180: logw + logp(x, phi_new) - logp(x, phi_old) at the PRE-move x."""
        A, B = self._split_x
        out = dev.empty(self.n_local)
        _cabi.call("smcb_reweight_asymptotic", dev.ptr(self.logw), dev.ptr(A), dev.ptr(B), float(self.phi_new),
                   float(self.phi_old), self.n_local, dev.ptr(out), dev.stream_ptr())
        return out

    def _non_asympototic_reweight(self):
        """samples.py:183-196: logp at phi = 1 regardless of tempering."""
        st, n = dev.stream_ptr(), self.n_local
        out = dev.empty(n)
        fused = isinstance(self.lkernel, ForwardLKernel) and self._ke is not None and \
            dev.is_std_normal(self.forward_kernel.momentum_proposal, self.D)
        if fused:  # L(r_new) - q(r) = -ke_new + ke_old, both emitted by the NUTS kernel; logp(., 1) from its split densities
            _cabi.call("smcb_reweight_forward_split", dev.ptr(self.logw), dev.ptr(self._split_x[0]), dev.ptr(self._split_x[1]),
                       dev.ptr(self._split_new[0]), dev.ptr(self._split_new[1]), dev.ptr(self._ke[0]), dev.ptr(self._ke[1]),
                       1.0, n, dev.ptr(out), st)
            return out
        lp_x = self.target.combine(*self._split_x, 1.0)
        lp_xnew = self.target.combine(*self._split_new, 1.0)
        L = dev.to_device(self.lkernel.calculate_L(self.r_new, self.x_new))
        q = dev.to_device(self.forward_kernel.logpdf(self.r))
        _cabi.call("smcb_reweight_general", dev.ptr(self.logw), dev.ptr(lp_x), dev.ptr(lp_xnew), dev.ptr(L), dev.ptr(q),
                   n, dev.ptr(out), st)
        return out

    def _tempering(self):
        """samples.py:199-212 -> adaptive_tempering.py:18-63, on the split at x_new."""
        A, B = self._split_new
        self.phi_new = self.TemperingScheme.calculate_phi_from_split(A, B, float(self.phi_old))
        return self.phi_new

    def update_samples(self):
        """samples.py:215-222."""
        self.phi_old = self.phi_new
        self.x = self.x_new
        self.logw = self.logw_new
        self.iteration += 1
