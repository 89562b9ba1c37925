"""Seeded synthetic inputs shaped like BASELINE.json's configs, numpy only."""
from __future__ import annotations

import numpy as np

DEV = 1.015658
SNP_COUNTS = np.array([704, 392, 379, 268, 279, 431, 388, 194, 223, 94, 328, 331, 35, 82, 165, 218,
                       311, 81, 338, 150, 77, 119], dtype=np.float64)


def segments(G: int, nseg: int = 22) -> np.ndarray:
    """22 chromosome-like segments with lengths ∝ the bundled per-chromosome SNP counts."""
    w = SNP_COUNTS[:nseg] if nseg <= 22 else np.ones(nseg)
    off = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * G)]).astype(np.int64)
    off[-1] = G
    return off


def clone_of(N: int, rng) -> np.ndarray:
    return rng.choice(3, size=N, p=[0.5, 0.3, 0.2])


def expression(G: int, N: int, seed: int = 2):
    """Gene-major x[G, N]: clones A (+seg6, -seg9), B (-seg12, -first half of seg13), C none."""
    rng = np.random.default_rng(seed)
    seg = segments(G)
    clone = clone_of(N, rng)
    sig = rng.uniform(2.0, 3.1, size=N)
    state = np.zeros((G, 3))
    state[seg[6]:seg[7], 0] = 1
    state[seg[9]:seg[10], 0] = -1
    state[seg[12]:seg[13], 1] = -1
    state[seg[13]:(seg[13] + seg[14]) // 2, 1] = -1
    x = state[:, clone] * DEV + rng.standard_normal((G, N)) * sig
    return np.ascontiguousarray(x), seg, clone


def allele(G: int, N: int, seed: int = 3, deep: float = 0.0):
    """Gene-major (r_maf[G, N], n_sc[G, N]) int32; coverage prob 0.108, 1+NegBin(0.3, mean 3) <= 1245."""
    rng = np.random.default_rng(seed)
    seg = segments(G)
    clone = clone_of(N, rng)
    cov = rng.random((G, N)) < 0.108
    nb = 1 + rng.negative_binomial(0.3, 0.3 / (0.3 + 3.0), size=(G, N))
    n = np.where(cov, np.minimum(nb, 1245), 0).astype(np.int32)
    if deep > 0:
        big = rng.random((G, N)) < deep
        n[big] = rng.integers(300, 1246, size=int(big.sum()))
    p = np.full((G, 3), 0.45)
    p[seg[9]:seg[10], 0] = 0.1
    p[seg[12]:seg[13], 1] = 0.1
    p[seg[13]:(seg[13] + seg[14]) // 2, 1] = 0.1
    r = rng.binomial(n, p[:, clone]).astype(np.int32)
    return np.ascontiguousarray(r), np.ascontiguousarray(n), seg, clone
