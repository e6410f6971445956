"""Device plumbing: torch owns memory and streams, the C-ABI gets raw pointers.

Nothing here computes; every numeric operation is a call into libsmcnuts_b200.so.
"""
import numpy as np
import torch

from . import _cabi

F64 = torch.float64


_CUDA_OK = None
_DEVICES = {}


def device():
    """torch.device of the current CUDA device (cached objects: this is called for every allocation)."""
    global _CUDA_OK
    if _CUDA_OK is None:
        _CUDA_OK = torch.cuda.is_available()
    if not _CUDA_OK:
        raise _cabi.SmcbError("smcnuts device path needs a CUDA device (no CPU fallback exists)")
    i = torch.cuda.current_device()
    d = _DEVICES.get(i)
    if d is None:
        d = _DEVICES[i] = torch.device("cuda", i)
    return d


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def is_host(a):
    return not isinstance(a, torch.Tensor)


def to_device(a, dtype=F64):
    """numpy / torch(any device) -> contiguous CUDA tensor (no copy if already there)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device(), dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device())


def to_numpy(a):
    """Any container (CUDA / CPU tensor, numpy, list) -> host numpy array."""
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def like_input(t, ref):
    """Return `t` in the container type of `ref` (numpy in -> numpy out)."""
    if is_host(ref):
        return t.cpu().numpy()
    return t


_PINNED = {}


def pinned_buffer(tag, shape, dtype):
    """Reusable page-locked staging buffer (cudaHostAlloc costs milliseconds; never allocate one per call)."""
    key = (tag, tuple(shape), dtype)
    buf = _PINNED.get(key)
    if buf is None:
        buf = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        _PINNED[key] = buf
    return buf


def to_host_like(t, ref, tag="out"):
    """Device result -> the container type of `ref`.

    CUDA tensor in -> CUDA tensor out.  numpy in -> a fresh numpy array (copied out of a reusable pinned staging
    buffer).  Pinned CPU torch tensor in -> the pinned staging buffer itself is returned (zero extra copy; it is
    overwritten by the next call with the same tag and shape -- the fast path bench.py's e2e leg uses)."""
    if isinstance(ref, torch.Tensor) and ref.is_cuda:
        return t
    host = pinned_buffer(tag, t.shape, t.dtype)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    if isinstance(ref, torch.Tensor):
        return host if ref.is_pinned() else host.clone()
    return host.numpy().copy()


def host_view(a, n, d):
    """Host array / CPU tensor -> ([n, d] float64 CPU tensor sharing its memory when possible, is_pinned)."""
    if isinstance(a, torch.Tensor):
        t = a.to(dtype=F64).contiguous().reshape(n, d)
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).reshape(n, d)
    return t, t.is_pinned()


def host_result(pinned, ref):
    """Pinned staging buffer holding a result -> the container type of the host input `ref` (see to_host_like)."""
    if isinstance(ref, torch.Tensor):
        return pinned if ref.is_pinned() else pinned.clone()
    return pinned.numpy().copy()


_STREAMS = {}


def side_streams(n):
    """Reusable non-default streams of the current device (copy/compute pipelining)."""
    key = torch.cuda.current_device()
    pool = _STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream())
    return pool[:n]


def empty(*shape, dtype=F64):
    return torch.empty(*shape, dtype=dtype, device=device())


def zeros(*shape, dtype=F64):
    return torch.zeros(*shape, dtype=dtype, device=device())


_WS = {}


def workspace(tag, nbytes):
    """Grow-only scratch buffers keyed by purpose (the library never allocates)."""
    key = (tag, torch.cuda.current_device())
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device())
        _WS[key] = buf
    return buf


_REDUCE_WS_BYTES = None


def reduce_ws():
    global _REDUCE_WS_BYTES
    if _REDUCE_WS_BYTES is None:
        _REDUCE_WS_BYTES = _cabi.lib().smcb_reduce_workspace_bytes()
    return workspace("reduce", _REDUCE_WS_BYTES)


def seed_from_rng(rng):
    """Map the reference's `rng` argument to a 64-bit Philox seed.

    int -> used as is; numpy Generator / RandomState -> one draw from it (so runs stay reproducible
    from the caller's seed); anything with a `.seed` int attribute -> that; None -> 0.
    """
    if rng is None:
        return 0
    if isinstance(rng, (int, np.integer)):
        return int(rng) & (2 ** 64 - 1)
    if isinstance(rng, np.random.Generator):
        return int(rng.integers(0, 2 ** 63 - 1))
    if isinstance(rng, np.random.RandomState):
        return int(rng.randint(0, 2 ** 31 - 1)) | (int(rng.randint(0, 2 ** 31 - 1)) << 31)
    if hasattr(rng, "seed") and isinstance(getattr(rng, "seed"), (int, np.integer)):
        return int(rng.seed)
    raise TypeError(f"cannot derive a Philox seed from rng={rng!r}")


def is_std_normal(dist, dim):
    """True for N(0, I_dim): our StdNormal or a scipy frozen multivariate_normal with mean 0, cov I."""
    from .distributions import StdNormal
    if isinstance(dist, StdNormal):
        return dist.dim == dim
    mean, cov = getattr(dist, "mean", None), getattr(dist, "cov", None)
    if mean is None or cov is None or callable(mean):
        return False
    mean, cov = np.atleast_1d(np.asarray(mean, dtype=float)), np.atleast_2d(np.asarray(cov, dtype=float))
    return mean.shape == (dim,) and cov.shape == (dim, dim) and not mean.any() and np.array_equal(cov, np.eye(dim))
