"""TEST INFRASTRUCTURE: numpy restatement of csrc/rng.cuh (Philox4x32-10 dropout masks), so that the CPU oracle can run
BERT with exactly the masks the sm_100a kernels regenerate from (seed, offset, element index)."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4(seed: int, ctr: np.ndarray) -> np.ndarray:
    """ctr: uint64 array [n] -> uint32 [n, 4]"""
    ctr = ctr.astype(np.uint64)
    c = [(ctr & np.uint64(0xFFFFFFFF)).astype(np.uint64), (ctr >> np.uint64(32)).astype(np.uint64),
         np.zeros_like(ctr), np.zeros_like(ctr)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def keep_mask(seed: int, offset: int, n: int, p: float) -> np.ndarray:
    """bool [n]: element idx is kept iff philox(seed, offset + idx // 4)[idx % 4] >= floor(p * 2^32)"""
    groups = (n + 3) // 4
    w = philox4(seed, np.uint64(offset) + np.arange(groups, dtype=np.uint64)).reshape(-1)[:n]
    t = p * 4294967296.0
    thresh = 0xFFFFFFFF if t >= 4294967295.0 else int(t)
    return w >= np.uint32(thresh)
