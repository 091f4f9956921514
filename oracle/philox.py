"""CPU restatement (numpy) of the in-kernel latent generator of bcnf_flow_sample -- test infrastructure.

The reference draws z with ``sigma * torch.randn(...)`` on the CPU generator (src/bcnf/models/cnf.py:566, :578, :584);
a device kernel cannot reproduce that stream, so the B200 path defines its own: Philox4x32-10 (Salmon, Moraes, Dror,
Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the generator behind curand / torch CUDA) keyed by the
caller's 64-bit seed, one block per 4 consecutive elements of the row-major (n_rows, D) array, turned into normals by
Box-Muller.  `philox4x32_10` is pinned against the Random123 known-answer vectors in tests/test_philox_cpu.py; the
GPU kernel (csrc/common.cuh: philox_normal) is pinned against `normal_field`.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
CTR_HI = (np.uint32(0x6263), np.uint32(0x6e66))      # counter words 2, 3 of every block ("bc", "nf")


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = k0 + W0
            k1 = k1 + W1
    return np.stack(c, axis=-1)


def normal_field(seed: int, n_rows: int, d: int, sigma: float = 1.0) -> np.ndarray:
    """(n_rows, d) float32: element e = row * d + j is word e & 3 of the block at counter e >> 2 (words 2, 3 of the
    counter fixed), two Box-Muller pairs per block: (w0, w1) -> (r cos, r sin), (w2, w3) -> (r cos, r sin)."""
    n = n_rows * d
    e = np.arange(n, dtype=np.uint64)
    blk = e >> np.uint64(2)
    ctr = np.stack([blk.astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
                    np.full(n, CTR_HI[0], np.uint32), np.full(n, CTR_HI[1], np.uint32)], axis=-1)
    key = np.stack([np.full(n, np.uint32(seed & 0xFFFFFFFF)), np.full(n, np.uint32((seed >> 32) & 0xFFFFFFFF))], axis=-1)
    w = philox4x32_10(ctr, key)
    lane = (e & np.uint64(3)).astype(np.int64)
    pair = lane & 2
    a = np.take_along_axis(w, pair[:, None], axis=1)[:, 0]
    b = np.take_along_axis(w, (pair + 1)[:, None], axis=1)[:, 0]
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    u2 = (b >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    r = np.sqrt(np.float32(-2.0) * np.log(u1.astype(np.float64))).astype(np.float32)
    ang = 2.0 * np.pi * u2.astype(np.float64)
    z = r * np.where(lane & 1, np.sin(ang), np.cos(ang)).astype(np.float32)
    return (np.float32(sigma) * z).reshape(n_rows, d).astype(np.float32)
