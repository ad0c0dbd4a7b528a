"""TEST INFRASTRUCTURE ONLY -- numpy reference of the engine's counter-based RNG.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11): integer arithmetic, so the CUDA implementation must match
BIT-EXACTLY.  Engine convention (vi-hmc_b200/csrc/common.cuh): key = 64-bit seed (lo, hi); counter =
(global chain id, iteration, block j, stream) with stream 0 = momentum, 1 = Metropolis uniform, 2 = VI redraw;
block j yields the four values for coordinates 4j..4j+3.  u = ((x >> 8) + 0.5) / 2^24 (exact in fp32);
normals by Box-Muller: (x0,x1) -> r cos(2 pi u2), r sin(2 pi u2), r = sqrt(-2 ln u1); likewise (x2,x3).

Known-answer vectors: Random123's kat_vectors for philox4x32-10 (counter/key all zero, all ones, pi digits).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays; returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32).copy() for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u32_to_unit(x):
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def _stream_word(stream, chain):
    return np.uint32(stream) ^ ((np.asarray(chain, dtype=np.uint64) >> np.uint64(32)).astype(np.uint32) << np.uint32(8))


def normals(seed: int, chain0: int, chains: int, iteration: int, d: int, stream: int = 0) -> np.ndarray:
    """[chains, d] float64-accurate Box-Muller normals of the engine's stream (compare to fp32 with ~1e-6 tol)."""
    nb = (d + 3) // 4
    chain = (np.arange(chains, dtype=np.uint64) + np.uint64(chain0))[:, None]
    j = np.arange(nb, dtype=np.uint32)[None, :]
    x0, x1, x2, x3 = philox4x32_10((chain & MASK).astype(np.uint32), np.uint32(iteration), j, _stream_word(stream, chain),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((chains, nb, 4), dtype=np.float64)
    for a, b, o in ((x0, x1, 0), (x2, x3, 2)):
        u1 = u32_to_unit(a).astype(np.float64)
        u2 = u32_to_unit(b).astype(np.float64)
        r = np.sqrt(-2.0 * np.log(u1))
        out[:, :, o] = r * np.cos(2 * np.pi * u2)
        out[:, :, o + 1] = r * np.sin(2 * np.pi * u2)
    return out.reshape(chains, nb * 4)[:, :d]


def uniforms(seed: int, chain0: int, chains: int, iteration: int) -> np.ndarray:
    """[chains] float32, bit-exact with the engine's Metropolis uniforms."""
    chain = np.arange(chains, dtype=np.uint64) + np.uint64(chain0)
    x0, _, _, _ = philox4x32_10((chain & MASK).astype(np.uint32), np.uint32(iteration), np.uint32(0), _stream_word(1, chain),
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return u32_to_unit(x0)


# Random123 known-answer vectors for philox4x32-10: (counter, key, expected)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
