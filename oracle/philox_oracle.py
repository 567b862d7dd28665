"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the counter-based RNG of tgtc-style_b200/csrc/philox.cuh.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 reference
implementation).  The reference program draws its stratified jitter and sigma noise from torch's global generator
(utils.py:519-520, :372-374); those streams depend on torch's kernel launch geometry and are not reproduced -- the library's
seeded training step draws from THIS generator instead, and this file pins it: the published known-answer vectors for the
block function, and the element -> (counter, word) mapping of the two tensor streams.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr, key):
    """ctr: uint32 [...,4], key: uint32 [...,2] -> uint32 [...,4]"""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k = [np.asarray(key[..., i], dtype=np.uint32).copy() for i in range(2)]
    for _ in range(10):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        with np.errstate(over="ignore"):
            k = [k[0] + W0, k[1] + W1]
    return np.stack(c, axis=-1)


def _blocks(seed, stream, blk, tag):
    blk = np.asarray(blk, dtype=np.uint64)
    ctr = np.stack([blk.astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32), np.full(blk.shape, stream, np.uint32),
                    np.full(blk.shape, tag, np.uint32)], axis=-1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32), blk.shape + (2,))
    return philox4x32_10(ctr, key)


def uniform(seed, stream, n):
    """element e: counter (e/4, stream, 0), word e%4, top 24 bits -> [0,1) (philox.cuh: philox_uniform)"""
    e = np.arange(n, dtype=np.uint64)
    v = _blocks(seed, stream, e >> np.uint64(2), 0)
    w = v[np.arange(n), (e & np.uint64(3)).astype(np.int64)]
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def normal(seed, stream, n, std=1.0):
    """element e: counter (e/2, stream, 1), words 2(e%2), 2(e%2)+1 -> sqrt(-2 ln u1) cos(2 pi u2) (philox.cuh: philox_normal)"""
    e = np.arange(n, dtype=np.uint64)
    v = _blocks(seed, stream, e >> np.uint64(1), 1)
    w = ((e & np.uint64(1)) * np.uint64(2)).astype(np.int64)
    a, b = v[np.arange(n), w], v[np.arange(n), w + 1]
    u1 = ((a >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (b >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    return (np.float32(std) * (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)).astype(np.float32)


# Random123 known-answer vectors (kat_vectors: "philox4x32 10")
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
