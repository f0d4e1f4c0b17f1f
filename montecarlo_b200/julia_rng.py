"""Host-side seeding of the device xoshiro256++ generator the way Julia seeds `Xoshiro(seed)`.

The reference builds one generator per chain, `rngs = [Xoshiro(seed + c - 1) for c in 1:M]` (src/metropolis.jl:262-263).
`Xoshiro(n::Integer)` is Julia's `Random` stdlib, not part of the reference tree [EXT]; Julia 1.7 - 1.10 define it as

    seed!(rng, n) = seed!(rng, make_seed(n))             make_seed: the 32-bit limbs of n, least significant first
    seed!(rng, v::Vector{UInt32}):  s0, s1, s2, s3 = reinterpret(UInt64, sha256(reinterpret(UInt8, v)))

i.e. the state is the SHA-256 digest of the seed's little-endian limbs, read as four little-endian 64-bit words
(Julia 1.11 changed the scheme).  No Julia toolchain exists in the build environment, but the recipe is **pinned by
the known answers printed in the Julia manual**: `rng = Xoshiro(1234); rand(rng, 2)` = [0.32597672886359486,
0.5490511363155669] (Xoshiro docstring) and `rng = Xoshiro(123); randn(rng, ComplexF64)` = -0.45660053706486897 -
1.0346749725929225im (randn docstring) are reproduced bit for bit from these states (tests/test_oracle.py::
test_julia_rng_known_answers on the host, tests/test_gpu_parity.py::test_xoshiro_device_draws_the_julia_manual_normals on
the device generator with the engine's default ziggurat tables = Julia's literal ki / wi / fi).

The hash is vectorised over chains with numpy (one 64-byte block per seed), so 2^24 chains seed in seconds.
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_seed", "sha256_blocks", "xoshiro_states", "xoshiro_state"]

_K = np.array([
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2], dtype=np.uint32)
_H0 = np.array([0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19],
               dtype=np.uint32)


def make_seed(n: int):
    """Random.make_seed(n::Integer) [EXT]: the 32-bit limbs of n >= 0, least significant first (at least one)."""
    n = int(n)
    if n < 0:
        raise ValueError("`n` must be non-negative.")       # Julia: DomainError
    limbs = []
    while True:
        limbs.append(n & 0xffffffff)
        n >>= 32
        if n == 0:
            return limbs


def _rotr(x, k):
    return (x >> np.uint32(k)) | (x << np.uint32(32 - k))


def sha256_blocks(w16: np.ndarray) -> np.ndarray:
    """SHA-256 compression of N independent single-block messages: w16 [N][16] big-endian message words (already
    padded) -> digests [N][8] as 32-bit words (FIPS 180-4 §6.2), vectorised over N."""
    w16 = np.ascontiguousarray(w16, dtype=np.uint32)
    n = w16.shape[0]
    w = [w16[:, i].copy() for i in range(16)]
    with np.errstate(over="ignore"):
        for i in range(16, 64):
            s0 = _rotr(w[i - 15], 7) ^ _rotr(w[i - 15], 18) ^ (w[i - 15] >> np.uint32(3))
            s1 = _rotr(w[i - 2], 17) ^ _rotr(w[i - 2], 19) ^ (w[i - 2] >> np.uint32(10))
            w.append(w[i - 16] + s0 + w[i - 7] + s1)
        a, b, c, d, e, f, g, h = (np.full(n, v, dtype=np.uint32) for v in _H0)
        for i in range(64):
            S1 = _rotr(e, 6) ^ _rotr(e, 11) ^ _rotr(e, 25)
            ch = (e & f) ^ (~e & g)
            t1 = h + S1 + ch + _K[i] + w[i]
            S0 = _rotr(a, 2) ^ _rotr(a, 13) ^ _rotr(a, 22)
            maj = (a & b) ^ (a & c) ^ (b & c)
            t2 = S0 + maj
            h, g, f, e, d, c, b, a = g, f, e, d + t1, c, b, a, t1 + t2
        out = np.stack([a, b, c, d, e, f, g, h], axis=1) + _H0
    return out


def _bswap32(x):
    return ((x & np.uint32(0xff)) << np.uint32(24)) | ((x & np.uint32(0xff00)) << np.uint32(8)) | \
           ((x >> np.uint32(8)) & np.uint32(0xff00)) | (x >> np.uint32(24))


def xoshiro_states(seeds) -> np.ndarray:
    """States [N][4] uint64 of `Xoshiro(seed)` for every non-negative integer in `seeds` (Julia 1.7 - 1.10 [EXT])."""
    seeds = np.ascontiguousarray(seeds)
    if seeds.size and (seeds.dtype.kind == "i") and seeds.min() < 0:
        raise ValueError("`n` must be non-negative.")
    s = seeds.astype(np.uint64)
    n = s.size
    lo = (s & np.uint64(0xffffffff)).astype(np.uint32)
    hi = (s >> np.uint64(32)).astype(np.uint32)
    two = hi != 0                                            # seeds >= 2^32 have two limbs
    w = np.zeros((n, 16), dtype=np.uint32)
    # message = limbs as little-endian bytes; SHA reads big-endian words -> byte-swap each limb
    w[:, 0] = _bswap32(lo)
    w[:, 1] = np.where(two, _bswap32(hi), np.uint32(0x80000000))      # 0x80 padding byte right after the message
    w[:, 2] = np.where(two, np.uint32(0x80000000), np.uint32(0))
    w[:, 15] = np.where(two, np.uint32(64), np.uint32(32))            # message length in bits
    h = sha256_blocks(w)
    # digest bytes = H0..H7 big-endian; reinterpret(UInt64, bytes) reads little-endian 8-byte groups
    hb = _bswap32(h).astype(np.uint64)
    return np.ascontiguousarray(hb[:, 0::2] | (hb[:, 1::2] << np.uint64(32)))


def xoshiro_state(seed: int) -> np.ndarray:
    """`Xoshiro(seed)` -> (s0, s1, s2, s3), any non-negative Python int (hashlib path for seeds of 3+ limbs)."""
    limbs = make_seed(seed)
    if len(limbs) <= 2:
        return xoshiro_states(np.array([seed], dtype=np.uint64))[0]
    import hashlib
    import struct
    d = hashlib.sha256(struct.pack("<%dI" % len(limbs), *limbs)).digest()
    return np.frombuffer(d, dtype="<u8").copy()
