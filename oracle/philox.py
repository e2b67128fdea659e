"""ORACLE — test infrastructure only (never imported by the product path).

Numpy restatement of the counter-based RNG the CUDA kernels use for
`eps ~ N(0,1)` (the reference draws eps with `torch.randn_like`,
BayTorch/modules/module.py:82-85; the stream itself is ours, so parity on eps is
"same eps injected on both sides", and this file pins OUR stream so that it can be
injected into the reference restatement).

Stream definition (must match mfvi_dip_mia_b200/csrc/philox.cuh bit for bit on the
integer part):

    Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi),
    counter = (block, stream_id, sample_id, step) with block = element_index // 4;
    the 4 output words of one block give 4 normals through two Box-Muller pairs:
        u1 = ((x0 >> 8) + 0.5) * 2^-24 , u2 = ((x1 >> 8) + 0.5) * 2^-24
        r  = sqrt(-2 ln u1) ; z0 = r cos(2 pi u2) ; z1 = r sin(2 pi u2)
        (x2, x3) -> (z2, z3) likewise; element_index % 4 selects z0..z3.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All args broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def _box_muller(xa, xb):
    u1 = ((xa >> np.uint32(8)).astype(np.float64) + 0.5) * 2.0 ** -24
    u2 = ((xb >> np.uint32(8)).astype(np.float64) + 0.5) * 2.0 ** -24
    r = np.sqrt(-2.0 * np.log(u1))
    th = 2.0 * np.pi * u2
    return r * np.cos(th), r * np.sin(th)


def philox_raw(n, seed, stream_id, sample_id, step):
    """uint32[n_blocks,4] raw words for elements 0..n-1 (n rounded up to 4)."""
    nb = (n + 3) // 4
    blk = np.arange(nb, dtype=np.uint32)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    x = philox4x32_10(blk, np.uint32(stream_id), np.uint32(sample_id), np.uint32(step), k0, k1)
    return np.stack(x, axis=1)


def philox_normal(n, seed, stream_id, sample_id, step):
    """float32[n] standard normals of stream (seed, stream_id, sample_id, step)."""
    raw = philox_raw(n, seed, stream_id, sample_id, step)
    z0, z1 = _box_muller(raw[:, 0], raw[:, 1])
    z2, z3 = _box_muller(raw[:, 2], raw[:, 3])
    z = np.stack([z0, z1, z2, z3], axis=1).reshape(-1)[:n]
    return z.astype(np.float32)


# Known-answer test from the Random123 distribution (kat_vectors, philox4x32 10 rounds):
#   counter = 0, key = 0                  -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
#   counter = ffffffff x4, key = ffffffff -> 408f276d 41c83b0e a20bc7c6 6d5451fd
#   counter = 243f6a88 85a308d3 13198a2e 03707344, key = a4093822 299f31d0
#                                         -> d16cfe09 94fdcceb 5001e420 24126ea1
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]
