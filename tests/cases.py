"""Shared seeded cases for the parity tests (oracle and CUDA path see identical inputs)."""
from __future__ import annotations

import numpy as np

from merfish3d_analysis_b200 import synthetic
from oracle import decode_oracle as orc


def codebook16(n_blank: int = 10):
    m = synthetic.mhd4_codebook_matrix(16)
    df = synthetic.codebook_dataframe(m, n_blank=n_blank)
    return df, orc.load_codebook(df, 16)


def codebook22(n_words: int = 300, seed: int = 4004, n_blank: int = 10):
    m = synthetic.random_hw4_codebook_matrix(22, n_words, seed)
    df = synthetic.codebook_dataframe(m, n_blank=n_blank)
    return df, orc.load_codebook(df, 22)


def small_stack(matrix, shape=(12, 48, 64), seed=7, density=2.0e-3):
    """uint16 (bits, z, y, x) with enough spots to give tens of features."""
    return synthetic.make_stack(matrix, shape, seed, density=density)


def simple_vectors(n_bits, bkg=200.0, nrm=900.0, seed=0):
    rng = np.random.default_rng(seed)
    b = (bkg + rng.uniform(-5, 5, n_bits)).astype(np.float32)
    n = (nrm + rng.uniform(-50, 50, n_bits)).astype(np.float32)
    return b, n


def canonical_labels(labels: np.ndarray) -> np.ndarray:
    """Relabel surviving components 1..n in raster order of their first voxel."""
    flat = labels.ravel()
    fg = np.flatnonzero(flat)
    out = np.zeros_like(flat, dtype=np.int32)
    if fg.size:
        vals = flat[fg]
        _, first = np.unique(vals, return_index=True)
        order = np.argsort(first, kind="stable")
        uniq = np.unique(vals)
        rank = np.empty(uniq.size, dtype=np.int32)
        rank[order] = np.arange(1, uniq.size + 1, dtype=np.int32)
        out[fg] = rank[np.searchsorted(uniq, vals)]
    return out.reshape(labels.shape)
