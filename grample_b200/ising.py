"""Synthetic binary Ising torus (BASELINE.json config 5).

Variables row-major (id = r*W + c).  Factors mirror res/Grids_11.uai's file order: n unary
factors, then the horizontal pairs row by row (wrap-around pair last in each row), then the
vertical pairs column by column; each pair scope is (smaller id, larger id).  Tables:
unary [e^h, e^-h], h ~ U(-1, 1); pairwise [e^w, e^-w, e^-w, e^w], w ~ U(-wmax, wmax).
Generator: numpy PCG64 with a fixed seed (default 20260101).
"""
import numpy as np


def ising_torus(height, width, wmax=4.9, seed=20260101):
    """Returns the C-ABI model arrays (card, fixed, scope_off, scope_vars, tab_off, tables)."""
    H, W = int(height), int(width)
    n = H * W
    rng = np.random.Generator(np.random.PCG64(seed))
    h = rng.uniform(-1.0, 1.0, size=n)
    wh = rng.uniform(-wmax, wmax, size=n)  # horizontal couplings, indexed by (r, c) -> (r, c+1)
    wv = rng.uniform(-wmax, wmax, size=n)  # vertical couplings, indexed by (c, r) -> (r+1, c)

    ids = np.arange(n, dtype=np.int64).reshape(H, W)
    right = np.roll(ids, -1, axis=1)
    down = np.roll(ids, -1, axis=0)
    hp = np.stack([np.minimum(ids, right), np.maximum(ids, right)], axis=-1).reshape(n, 2)
    vp = np.stack([np.minimum(ids, down), np.maximum(ids, down)], axis=-1).transpose(1, 0, 2).reshape(n, 2)

    card = np.full(n, 2, dtype=np.int32)
    fixed = np.full(n, -1, dtype=np.int32)
    scope_vars = np.concatenate([np.arange(n, dtype=np.int64), hp.reshape(-1), vp.reshape(-1)]).astype(np.int32)
    scope_off = np.concatenate([np.arange(n + 1, dtype=np.int64), n + 2 * np.arange(1, 2 * n + 1, dtype=np.int64)]).astype(np.int32)
    tab_off = np.concatenate([2 * np.arange(n + 1, dtype=np.int64), 2 * n + 4 * np.arange(1, 2 * n + 1, dtype=np.int64)])
    unary = np.stack([np.exp(h), np.exp(-h)], axis=-1).reshape(-1)
    w = np.concatenate([wh, wv])
    pair = np.stack([np.exp(w), np.exp(-w), np.exp(-w), np.exp(w)], axis=-1).reshape(-1)
    tables = np.concatenate([unary, pair])
    return card, fixed, scope_off, scope_vars, tab_off, tables
