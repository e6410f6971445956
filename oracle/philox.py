"""TEST INFRASTRUCTURE ONLY (oracle).  Philox4x32-10 in numpy + the replay RNG.

The reference draws every random number of the particle loop from ONE sequential
numpy stream shared by all particles (/root/reference/smcnuts/proposal/nuts.py:50-53,69,91,99,142;
nuts_acc_rej.py:46-47).  The B200 path replaces that with one counter-based Philox4x32-10
stream per (seed, iteration, stream-id, particle) so the result does not depend on
scheduling or GPU count.  This file is the normative CPU definition of those streams;
`ReplayRNG` exposes them through the numpy-style `.exponential()/.uniform()` methods the
unmodified reference calls, so reference and device consume identical draws.

Stream layout (shared with smc-nuts_b200/csrc/philox.cuh and oracle/smc_oracle.c):
    key  = (seed & 0xffffffff, seed >> 32)
    ctr  = (block, (iteration << 8) | stream_id, particle & 0xffffffff, particle >> 32)
    draw p of a stream -> block = p >> 1, words (2*(p&1), 2*(p&1)+1)
    u64  = (w[2h+1] << 32) | w[2h];   uniform = (u64 >> 11) * 2**-53  in [0,1)
    exponential(1) = -log1p(-uniform)
    normal pair j  = Box-Muller on draws (2j, 2j+1):  rad = sqrt(-2*log1p(-u1)),
                     z0 = rad*cos(2*pi*u2), z1 = rad*sin(2*pi*u2)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_NUTS = 0      # slice variable, directions, multinomial merges (nuts.py:69,91,99,142)
STREAM_MOMENTUM = 1  # r ~ N(0, I)                                   (samples.py:155)
STREAM_ACCREJ = 2    # endpoint MH uniform                            (utils.py:32)
STREAM_RESAMPLE = 3  # resampling uniforms, particle = output slot    (samples.py:139)
STREAM_INIT = 4      # x0 ~ N(0, I)                                   (samples.py:77)
STREAM_ESTIMATE = 5  # estimate_from_tempered resampling              (estimate_from_tempered.py:43)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    key = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = key[..., 0].copy(), key[..., 1].copy()
    for rnd in range(10):
        if rnd:
            k0 = (k0 + np.uint64(W0)) & MASK32
            k1 = (k1 + np.uint64(W1)) & MASK32
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _blocks(seed, iteration, stream, particle, block):
    particle = np.asarray(particle, dtype=np.uint64)
    block = np.asarray(block, dtype=np.uint64)
    particle, block = np.broadcast_arrays(particle, block)
    ctr = np.empty(particle.shape + (4,), dtype=np.uint64)
    ctr[..., 0] = block & MASK32
    ctr[..., 1] = np.uint64(((int(iteration) << 8) | int(stream)) & 0xFFFFFFFF)
    ctr[..., 2] = particle & MASK32
    ctr[..., 3] = particle >> np.uint64(32)
    key = np.empty(particle.shape + (2,), dtype=np.uint64)
    key[..., 0] = np.uint64(int(seed) & 0xFFFFFFFF)
    key[..., 1] = np.uint64((int(seed) >> 32) & 0xFFFFFFFF)
    return philox4x32_10(ctr, key)


def uniform(seed, iteration, stream, particle, draw):
    """Uniform [0,1) for draw index `draw` (array ok) of the given stream."""
    draw = np.asarray(draw, dtype=np.uint64)
    w = _blocks(seed, iteration, stream, particle, draw >> np.uint64(1)).astype(np.uint64)
    half = np.broadcast_to((draw & np.uint64(1)).astype(np.int64), w.shape[:-1])
    lo = np.where(half == 0, w[..., 0], w[..., 2])
    hi = np.where(half == 0, w[..., 1], w[..., 3])
    u64 = (hi << np.uint64(32)) | lo
    return (u64 >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def normals(seed, iteration, stream, particles, dim):
    """(len(particles), dim) standard normals, Box-Muller pairs (2j, 2j+1) per particle."""
    particles = np.asarray(particles, dtype=np.uint64)
    npair = (dim + 1) // 2
    j = np.arange(npair, dtype=np.uint64)
    u1 = uniform(seed, iteration, stream, particles[:, None], 2 * j[None, :])
    u2 = uniform(seed, iteration, stream, particles[:, None], 2 * j[None, :] + 1)
    rad = np.sqrt(-2.0 * np.log1p(-u1))
    ang = 2.0 * np.pi * u2
    z = np.empty((len(particles), 2 * npair))
    z[:, 0::2] = rad * np.cos(ang)
    z[:, 1::2] = rad * np.sin(ang)
    return z[:, :dim]


class ReplayRNG:
    """numpy-RNG look-alike that serves one particle's Philox stream sequentially.

    Hand it to the unmodified reference `NUTSProposal(..., rng=ReplayRNG(...))`; call
    `set_particle(i)` before each particle (the reference loops particles in order,
    nuts.py:50-53).  Every draw is logged for inspection.
    """

    def __init__(self, seed, iteration=0, stream=STREAM_NUTS, particle=0):
        self.seed, self.iteration, self.stream = seed, iteration, stream
        self.set_particle(particle)

    def set_particle(self, particle, stream=None):
        if stream is not None:
            self.stream = stream
        self.particle = int(particle)
        self.pos = 0
        self.log = []

    def _next(self):
        u = float(uniform(self.seed, self.iteration, self.stream, self.particle, self.pos))
        self.pos += 1
        return u

    def uniform(self, low=0.0, high=1.0, size=None):
        assert size is None
        u = low + (high - low) * self._next()
        self.log.append(("uniform", u))
        return u

    def exponential(self, scale=1.0, size=None):
        assert size is None
        e = -np.log1p(-self._next()) * scale
        self.log.append(("exponential", e))
        return e


# Random123 known-answer vectors for philox4x32-10 (kat_vectors): (ctr, key, expected)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]
