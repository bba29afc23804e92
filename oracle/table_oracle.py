"""CPU oracle for the post-decode table stage: blank-fraction filter,
tile-overlap de-duplication, within-tile (2-D mode) cluster de-duplication.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

NumPy/SciPy restatement of PD:3386-3846 (``_filter_all_barcodes_blank_fraction``),
PD:4137-4177 (``_remove_duplicates_in_tile_overlap``) and PD:4179-4363
(``_remove_duplicates_within_tile``).  PINNED: ``tests/golden/make_table_golden.py`` runs those
reference methods themselves (they only need NumPy / SciPy / pandas once the module imports
through ``reference_shims``) on a seeded transcript table and ``tests/test_cpu_table_stage.py``
requires these functions to reproduce their output.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.spatial import cKDTree


def is_blank(gene_ids) -> np.ndarray:
    s = pd.Series(list(gene_ids)).astype("string").str.lower().str.startswith("blank", na=False)
    return s.to_numpy(dtype=bool)


# ------------------------------------------------------------------------- histogram edges
def _fix_degenerate(edges: np.ndarray, values: np.ndarray) -> np.ndarray:
    """PD:3508-3528 (and twins): finite, >= 2 edges, non-collapsed, covering the data; the last
    edge is nudged up so the maximum falls inside the last bin."""
    edges = edges[np.isfinite(edges)]
    if edges.size < 2 or np.allclose(edges[0], edges[-1]):
        center = float(np.mean(values)) if edges.size < 2 else float(edges[0])
        edges = np.array([center - 0.5, center + 0.5], dtype=float)
    edges[0] = min(edges[0], float(np.min(values)))
    edges[-1] = max(edges[-1], float(np.max(values)))
    edges[-1] = np.nextafter(edges[-1], np.inf)
    return edges


def intensity_edges(values: np.ndarray) -> np.ndarray:
    """PD:3504-3528: deciles of magnitude_mean."""
    return _fix_degenerate(np.unique(np.quantile(values, np.linspace(0.0, 1.0, 11))), values)


def voxel_number_edges(values: np.ndarray) -> np.ndarray:
    """PD:3544-3602: unit bins when the area range is <= 10 wide, else floor(deciles) - 0.5."""
    lo = int(np.floor(np.min(values)))
    hi = int(np.ceil(np.max(values)))
    if hi - lo + 1 <= 10:
        edges = np.arange(lo - 0.5, hi + 1.5, 1.0)
    else:
        q = np.unique(np.floor(np.quantile(values, np.linspace(0.0, 1.0, 11))).astype(float))
        if q.size == 0:
            q = np.array([float(lo), float(hi + 1)])
        if q[0] > lo:
            q = np.insert(q, 0, float(lo))
        if q[-1] <= hi:
            q = np.append(q, float(hi + 1))
        edges = q - 0.5
    return _fix_degenerate(np.unique(np.asarray(edges, dtype=float)), values)


def vector_distance_edges(values: np.ndarray) -> np.ndarray:
    """PD:3618-3655: 10 equal-width bins over [min, max] of distance_min."""
    e = np.linspace(float(np.min(values)), float(np.max(values)), 11)
    return _fix_degenerate(np.unique(np.asarray(e, dtype=float)), values)


def explicit_edges(bins) -> np.ndarray:
    e = np.unique(np.asarray(bins, dtype=float))
    e = e[np.isfinite(e)]
    if e.size < 2:
        raise ValueError("Explicit histogram edges must contain at least two finite values.")
    e[-1] = np.nextafter(e[-1], np.inf)
    return e


# ------------------------------------------------------------------------- blank-fraction filter
def blank_fraction_filter(df: pd.DataFrame, blank_count: int, barcode_count: int,
                          target_gross_misid_rate: float = 0.05, intensity_bins=None, voxel_number_bins=None,
                          vector_distance_bins=None) -> tuple[np.ndarray, dict]:
    """PD:3386-3846.  Returns (keep mask over df rows, diagnostics)."""
    n = len(df)
    diag: dict = {"target_gross_misid_rate": float(target_gross_misid_rate), "chosen_threshold": np.nan,
                  "achieved_gross_misid_rate": np.inf, "target_reached": False}
    keep = np.zeros(n, dtype=bool)
    if n == 0:
        diag["reason"] = "no_transcripts"
        return keep, diag
    inten64 = df["magnitude_mean"].to_numpy(dtype=float)
    area64 = df["area"].to_numpy(dtype=float)
    dist64 = df["distance_min"].to_numpy(dtype=float)
    inten, area, dist = (v.astype(np.float32) for v in (inten64, area64, dist64))  # PD:3460-3468
    blank = is_blank(df["gene_id"])
    valid = np.isfinite(inten) & np.isfinite(area) & np.isfinite(dist)
    if not valid.any():
        diag["reason"] = "no_valid_features"
        return keep, diag
    if blank_count <= 0:
        diag["reason"] = "no_blank_barcodes"
        return valid.copy(), diag
    if not blank[valid].any():
        diag["reason"] = "no_blank_transcripts"
        return valid.copy(), diag
    e0 = explicit_edges(intensity_bins) if intensity_bins is not None else intensity_edges(inten64[valid])
    e1 = explicit_edges(voxel_number_bins) if voxel_number_bins is not None else voxel_number_edges(area64[valid])
    e2 = (explicit_edges(vector_distance_bins) if vector_distance_bins is not None
          else vector_distance_edges(dist64[valid]))
    diag.update(intensity_bins=e0, voxel_number_bins=e1, vector_distance_bins=e2)
    # binning happens on float32 copies of the edges and the values (PD:3661-3690)
    f0, f1, f2 = (e.astype(np.float32) for e in (e0, e1, e2))
    b0 = np.searchsorted(f0, inten, side="right") - 1
    b1 = np.searchsorted(f1, area, side="right") - 1
    b2 = np.searchsorted(f2, dist, side="right") - 1
    in_range = (valid & (b0 >= 0) & (b0 < f0.size - 1) & (b1 >= 0) & (b1 < f1.size - 1)
                & (b2 >= 0) & (b2 < f2.size - 1))
    if not in_range.any():
        diag["reason"] = "no_transcripts_in_histogram_range"
        return keep, diag
    shape = (f0.size - 1, f1.size - 1, f2.size - 1)
    flat = np.full(n, -1, dtype=np.int64)
    flat[in_range] = np.ravel_multi_index((b0[in_range], b1[in_range], b2[in_range]), dims=shape)
    all_hist = np.bincount(flat[in_range], minlength=int(np.prod(shape))).astype(np.int32)
    blank_hist = np.bincount(flat[in_range & blank], minlength=int(np.prod(shape))).astype(np.int32)
    frac_hist = np.full(all_hist.shape, np.nan, dtype=np.float32)
    nz = all_hist > 0
    frac_hist[nz] = (blank_hist[nz] / all_hist[nz]).astype(np.float32)  # int32/int32 -> float64 -> float32 store
    row_frac = np.full(n, np.nan)
    row_frac[in_range] = frac_hist[flat[in_range]]
    thresholds = np.unique(frac_hist[nz])
    sweep = []
    chosen = np.nan
    achieved = np.inf
    reached = False
    for t in thresholds:  # ascending: the LAST threshold meeting the target wins (PD:3774-3801)
        k = in_range & (row_frac <= float(t))
        if blank_count <= 0 or barcode_count <= 0 or not k.any():
            g = np.inf
        else:
            g = (np.count_nonzero(k & blank) / float(blank_count)) / (np.count_nonzero(k) / float(barcode_count))
        sweep.append((float(t), float(g), int(np.count_nonzero(k))))
        if g <= target_gross_misid_rate:
            chosen, achieved, keep, reached = float(t), float(g), k.copy(), True
    if not sweep:
        diag["reason"] = "no_nonempty_histogram_bins"
        return keep, diag
    if not reached:
        best = int(np.argmin([s[1] for s in sweep]))
        chosen, achieved = sweep[best][0], sweep[best][1]
        keep = in_range & (row_frac <= chosen)
    diag.update(chosen_threshold=chosen, achieved_gross_misid_rate=achieved, target_reached=reached,
                all_histogram=all_hist.reshape(shape), blank_histogram=blank_hist.reshape(shape),
                blank_fraction_histogram=frac_hist.reshape(shape), blank_fraction_bin=flat, blank_fraction=row_frac,
                threshold_sweep=pd.DataFrame(sweep, columns=["threshold", "gross_misid_rate", "kept_transcripts"]))
    return keep, diag


# ------------------------------------------------------------------------- de-duplication
def overlap_duplicates(coords_zyx: np.ndarray, tile_idx: np.ndarray, distance_min: np.ndarray,
                       radius: float = 0.75) -> np.ndarray:
    """PD:4137-4177: rows to DROP.  For every pair within ``radius`` (3-D, inclusive) that comes
    from two different tiles, the row with the larger (distance_min, row index) loses."""
    drop = np.zeros(len(coords_zyx), dtype=bool)
    if len(coords_zyx) < 2:
        return drop
    pairs = cKDTree(np.asarray(coords_zyx, dtype=float)).query_pairs(radius, output_type="ndarray")
    for i, j in pairs:
        if tile_idx[i] != tile_idx[j]:
            if (distance_min[i], i) <= (distance_min[j], j):
                drop[j] = True
            else:
                drop[i] = True
    return drop


def within_tile_duplicates(coords_zyx: np.ndarray, tile_idx: np.ndarray, gene_ids: np.ndarray,
                           distance_min: np.ndarray, radius_xy: float, radius_z: float) -> np.ndarray:
    """PD:4179-4363: rows to DROP.  Neighbours = same tile, same gene, XY distance <= radius_xy,
    0 < |dz| <= radius_z; per connected cluster keep the smallest (distance_min, row index)."""
    n = len(coords_zyx)
    drop = np.zeros(n, dtype=bool)
    if n < 2:
        return drop
    coords = np.asarray(coords_zyx, dtype=float)
    genes = np.asarray(gene_ids)
    for t in np.unique(tile_idx):
        loc = np.flatnonzero(tile_idx == t)
        if loc.size < 2:
            continue
        z = coords[loc, 0]
        pairs = cKDTree(coords[loc, 1:3]).query_pairs(r=radius_xy, output_type="ndarray")
        parent = np.arange(loc.size)

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        for i, j in pairs:
            if 0.0 < abs(z[i] - z[j]) <= radius_z and genes[loc[i]] == genes[loc[j]]:
                ri, rj = find(i), find(j)
                if ri != rj:
                    parent[max(ri, rj)] = min(ri, rj)
        roots = np.array([find(a) for a in range(loc.size)])
        for r in np.unique(roots):
            members = loc[roots == r]
            if members.size > 1:
                best = members[np.lexsort((members, distance_min[members]))][0]
                drop[members[members != best]] = True
    return drop


def assign_cells(points_yx: np.ndarray, polygons_yx: list[np.ndarray]) -> np.ndarray:
    """PD:4107-4135 restated without rtree / shapely (neither is installed, so this function cannot be pinned to the
    reference's own output; its containment predicate is cross-checked against OpenCV's pointPolygonTest, an independent
    implementation, in tests/test_cpu_table_stage.py):
    1 + index of the lowest-numbered polygon whose interior contains the point (even-odd rule, float64), else 0."""
    pts = np.asarray(points_yx, dtype=np.float64)
    out = np.zeros(len(pts), dtype=np.int64)
    for idx in range(len(polygons_yx) - 1, -1, -1):  # descending, so the lowest index is written last
        poly = np.asarray(polygons_yx[idx], dtype=np.float64)
        if len(poly) < 3:
            continue
        y0, x0, y1, x1 = poly[:, 0].min(), poly[:, 1].min(), poly[:, 0].max(), poly[:, 1].max()
        cand = np.flatnonzero((pts[:, 0] >= y0) & (pts[:, 0] <= y1) & (pts[:, 1] >= x0) & (pts[:, 1] <= x1))
        if cand.size == 0:
            continue
        py, px = pts[cand, 0], pts[cand, 1]
        inside = np.zeros(cand.size, dtype=bool)
        j = len(poly) - 1
        for k in range(len(poly)):
            yi, xi = poly[k]
            yj, xj = poly[j]
            cross = (yi > py) != (yj > py)
            with np.errstate(divide="ignore", invalid="ignore"):
                xc = (xj - xi) * (py - yi) / (yj - yi) + xi
            inside ^= cross & (px < xc)
            j = k
        out[cand[inside]] = idx + 1
    return out
