"""Deterministic synthetic MERFISH stacks and codebooks.

Host (NumPy) generators are used by the parity tests and the CPU baseline; the device
(torch) generator builds the full-size bench stacks directly in HBM.  Value model:
background uint16 ~ Poisson(100)+100 per bit; spots are 3-D Gaussian blobs
(sigma (1.5, 1.2, 1.2) px, log-normal amplitude around 1500) at uniform random positions,
density 1 spot / 1e4 voxels, each lighting the 4 on-bits of a random codeword with
+-20 % per-bit gain.
"""

from __future__ import annotations

from itertools import combinations

import numpy as np
import pandas as pd

SPOT_SIGMA = (1.5, 1.2, 1.2)
SPOT_RADIUS = (5, 4, 4)
SPOT_DENSITY = 1.0e-4
SPOT_AMPLITUDE = 1500.0
BACKGROUND_LAMBDA = 100.0
BACKGROUND_OFFSET = 100


def mhd4_codebook_matrix(n_bits: int = 16) -> np.ndarray:
    """The 140-word Hamming-weight-4 / distance-4 code on 16 bits.

    Words are the planes of AG(4,2): 4-subsets {a,b,c,d} of 0..15 with a^b^c^d == 0
    (a Steiner system S(3,4,16)), in lexicographic order.
    """
    if n_bits != 16:
        raise ValueError("the closed-form MHD4 code is defined for 16 bits")
    rows = []
    for a, b, c, d in combinations(range(16), 4):
        if a ^ b ^ c ^ d == 0:
            r = np.zeros(16, dtype=np.int64)
            r[[a, b, c, d]] = 1
            rows.append(r)
    return np.stack(rows)


def random_hw4_codebook_matrix(n_bits: int, n_words: int, seed: int) -> np.ndarray:
    """Distinct Hamming-weight-4 words (MHD >= 2); used for the 22-bit configuration."""
    all_words = list(combinations(range(n_bits), 4))
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(all_words), size=min(n_words, len(all_words)), replace=False))
    m = np.zeros((pick.size, n_bits), dtype=np.int64)
    for i, p in enumerate(pick):
        m[i, list(all_words[p])] = 1
    return m


def codebook_dataframe(matrix: np.ndarray, n_blank: int = 0) -> pd.DataFrame:
    """Codebook in the datastore layout: ``gene_id`` then ``bit01..bitNN`` (DS:826-841)."""
    k, b = matrix.shape
    names = [f"gene{i:04d}" for i in range(k - n_blank)] + [f"Blank{i:02d}" for i in range(n_blank)]
    df = pd.DataFrame(matrix, columns=[f"bit{i + 1:02d}" for i in range(b)])
    df.insert(0, "gene_id", names)
    return df


def _blob_patch():
    rz, ry, rx = SPOT_RADIUS
    z = np.arange(-rz, rz + 1)[:, None, None] / SPOT_SIGMA[0]
    y = np.arange(-ry, ry + 1)[None, :, None] / SPOT_SIGMA[1]
    x = np.arange(-rx, rx + 1)[None, None, :] / SPOT_SIGMA[2]
    return np.exp(-0.5 * (z * z + y * y + x * x)).astype(np.float32)


def make_stack(
    matrix: np.ndarray,
    shape_zyx: tuple[int, int, int],
    seed: int,
    density: float = SPOT_DENSITY,
    amplitude: float = SPOT_AMPLITUDE,
    return_truth: bool = False,
):
    """Host generator: (bits, z, y, x) uint16 stack, deterministic in ``seed``."""
    rng = np.random.default_rng(seed)
    k, b = matrix.shape
    Z, Y, X = shape_zyx
    n_vox = Z * Y * X
    acc = rng.poisson(BACKGROUND_LAMBDA, size=(b, Z, Y, X)).astype(np.float32)
    acc += BACKGROUND_OFFSET
    n_spots = max(1, int(round(n_vox * density)))
    cz = rng.integers(0, Z, n_spots)
    cy = rng.integers(0, Y, n_spots)
    cx = rng.integers(0, X, n_spots)
    word = rng.integers(0, k, n_spots)
    amp = amplitude * np.exp(rng.normal(0.0, 0.25, n_spots)).astype(np.float32)
    gain = rng.uniform(0.8, 1.2, size=(n_spots, b)).astype(np.float32)
    patch = _blob_patch()
    rz, ry, rx = SPOT_RADIUS
    for s in range(n_spots):
        z0, z1 = max(cz[s] - rz, 0), min(cz[s] + rz + 1, Z)
        y0, y1 = max(cy[s] - ry, 0), min(cy[s] + ry + 1, Y)
        x0, x1 = max(cx[s] - rx, 0), min(cx[s] + rx + 1, X)
        p = patch[
            z0 - cz[s] + rz : z1 - cz[s] + rz,
            y0 - cy[s] + ry : y1 - cy[s] + ry,
            x0 - cx[s] + rx : x1 - cx[s] + rx,
        ]
        for bit in np.flatnonzero(matrix[word[s]]):
            acc[bit, z0:z1, y0:y1, x0:x1] += (amp[s] * gain[s, bit]) * p
    stack = np.clip(np.rint(acc), 0, 65535).astype(np.uint16)
    if return_truth:
        return stack, dict(z=cz, y=cy, x=cx, word=word)
    return stack


def make_stack_device(
    matrix: np.ndarray,
    shape_zyx: tuple[int, int, int],
    seed: int,
    device="cuda",
    density: float = SPOT_DENSITY,
    amplitude: float = SPOT_AMPLITUDE,
    out=None,
):
    """Device generator (torch ops; data synthesis only, never on the decode path).

    Same value model as :func:`make_stack`; not bit-identical to it.  Builds one bit
    volume at a time so the peak temporary is one float32 volume.
    """
    import torch

    k, b = matrix.shape
    Z, Y, X = shape_zyx
    n_vox = Z * Y * X
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    n_spots = max(1, int(round(n_vox * density)))
    cz = torch.randint(0, Z, (n_spots,), generator=g, device=device)
    cy = torch.randint(0, Y, (n_spots,), generator=g, device=device)
    cx = torch.randint(0, X, (n_spots,), generator=g, device=device)
    word = torch.randint(0, k, (n_spots,), generator=g, device=device)
    amp = amplitude * torch.exp(0.25 * torch.randn(n_spots, generator=g, device=device))
    gain = 0.8 + 0.4 * torch.rand((n_spots, b), generator=g, device=device)
    mat = torch.as_tensor(matrix, device=device, dtype=torch.bool)
    patch = torch.as_tensor(_blob_patch(), device=device)
    rz, ry, rx = SPOT_RADIUS
    oz, oy, ox = torch.meshgrid(
        torch.arange(-rz, rz + 1, device=device),
        torch.arange(-ry, ry + 1, device=device),
        torch.arange(-rx, rx + 1, device=device),
        indexing="ij",
    )
    oz, oy, ox, pf = oz.reshape(-1), oy.reshape(-1), ox.reshape(-1), patch.reshape(-1)
    if out is None:
        out = torch.empty((b, Z, Y, X), dtype=torch.uint16, device=device)
    lam = torch.full((Y, X), BACKGROUND_LAMBDA, device=device)
    for bit in range(b):
        vol = torch.empty((Z, Y, X), dtype=torch.float32, device=device)
        for z in range(Z):
            vol[z] = torch.poisson(lam, generator=g)
        vol += BACKGROUND_OFFSET
        sel = torch.nonzero(mat[word, bit]).reshape(-1)
        if sel.numel():
            for s0 in range(0, sel.numel(), 1 << 16):
                ss = sel[s0 : s0 + (1 << 16)]
                zz = cz[ss, None] + oz[None, :]
                yy = cy[ss, None] + oy[None, :]
                xx = cx[ss, None] + ox[None, :]
                ok = (zz >= 0) & (zz < Z) & (yy >= 0) & (yy < Y) & (xx >= 0) & (xx < X)
                val = (amp[ss] * gain[ss, bit])[:, None] * pf[None, :]
                lin = (zz * Y + yy) * X + xx
                vol.view(-1).index_put_((lin[ok],), val[ok], accumulate=True)
        out[bit] = vol.round_().clamp_(0, 65535).to(torch.uint16)
        del vol
    return out
