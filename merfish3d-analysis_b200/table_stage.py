"""Post-decode transcript-table stage behind the reference's method names.

Host side of ``_filter_all_barcodes_blank_fraction`` (PD:3386-3846), ``_filter_all_barcodes_LR``
(PD:3849-4058), ``_remove_duplicates_in_tile_overlap`` (PD:4137-4177) and
``_remove_duplicates_within_tile`` (PD:4179-4363).  Row-parallel work -- binning + histograms,
radius-neighbour searches, cluster resolution -- runs in ``csrc/table.cu`` through the C ABI
(``m3d_table_hist3d``, ``m3d_overlap_duplicates``, ``m3d_within_tile_duplicates``); what stays
here is O(bins) logic: histogram edges from quantiles, the threshold sweep, frame bookkeeping.
There is no CPU fallback for the device parts.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from ._capi import DecodeContext


def blank_mask(gene_ids) -> np.ndarray:
    """PD:3430-3436: case-insensitive ``blank*`` gene ids."""
    s = pd.Series(np.asarray(gene_ids, dtype=object)).astype("string")
    return s.str.lower().str.startswith("blank", na=False).to_numpy(dtype=bool)


# ------------------------------------------------------------------------- histogram edges
def _cover(edges: np.ndarray, values: np.ndarray) -> np.ndarray:
    """Common tail of the three edge recipes (PD:3508-3528, 3588-3616, 3632-3655)."""
    edges = edges[np.isfinite(edges)]
    if edges.size < 2:
        c = float(np.mean(values))
        edges = np.array([c - 0.5, c + 0.5], dtype=float)
    elif np.allclose(edges[0], edges[-1]):
        c = float(edges[0])
        edges = np.array([c - 0.5, c + 0.5], dtype=float)
    edges[0] = min(edges[0], float(np.min(values)))
    edges[-1] = max(edges[-1], float(np.max(values)))
    edges[-1] = np.nextafter(edges[-1], np.inf)
    return edges


def _explicit(bins) -> np.ndarray:
    e = np.unique(np.asarray(bins, dtype=float))
    e = e[np.isfinite(e)]
    if e.size < 2:
        raise ValueError("Explicit histogram edges must contain at least two finite values.")
    e[-1] = np.nextafter(e[-1], np.inf)
    return e


def _edges(intensity, area, distance, intensity_bins, voxel_number_bins, vector_distance_bins):
    deciles = np.linspace(0.0, 1.0, 11)
    if intensity_bins is not None:
        e0 = _explicit(intensity_bins)
    else:
        e0 = _cover(np.unique(np.quantile(intensity, deciles)), intensity)
    if voxel_number_bins is not None:
        e1 = _explicit(voxel_number_bins)
    else:
        lo, hi = int(np.floor(np.min(area))), int(np.ceil(np.max(area)))
        if hi - lo + 1 <= 10:
            e1 = np.arange(lo - 0.5, hi + 1.5, 1.0)
        else:
            q = np.unique(np.floor(np.quantile(area, deciles)).astype(float))
            if q.size == 0:
                q = np.array([float(lo), float(hi + 1)])
            if q[0] > lo:
                q = np.insert(q, 0, float(lo))
            if q[-1] <= hi:
                q = np.append(q, float(hi + 1))
            e1 = q - 0.5
        e1 = _cover(np.unique(np.asarray(e1, dtype=float)), area)
    if vector_distance_bins is not None:
        e2 = _explicit(vector_distance_bins)
    else:
        e2 = _cover(np.unique(np.linspace(float(np.min(distance)), float(np.max(distance)), 11)), distance)
    return e0, e1, e2


def _empty_diagnostics(target: float) -> dict:
    return {
        "target_gross_misid_rate": float(target), "chosen_threshold": np.nan,
        "achieved_gross_misid_rate": np.inf, "target_reached": False,
        "all_histogram": np.zeros((0, 0, 0), dtype=np.int64), "blank_histogram": np.zeros((0, 0, 0), dtype=np.int64),
        "blank_fraction_histogram": np.zeros((0, 0, 0), dtype=float),
        "intensity_bins": np.array([], dtype=float), "voxel_number_bins": np.array([], dtype=float),
        "vector_distance_bins": np.array([], dtype=float),
        "threshold_sweep": pd.DataFrame(columns=["threshold", "gross_misid_rate", "kept_transcripts"]),
    }


def filter_blank_fraction(ctx: DecodeContext, loaded: pd.DataFrame, blank_count: int, barcode_count: int,
                          target_gross_misid_rate: float = 0.05, intensity_bins=None, voxel_number_bins=None,
                          vector_distance_bins=None) -> tuple[pd.DataFrame, dict]:
    """PD:3386-3846.  Returns (filtered frame incl. the reference's annotation columns and
    ``cell_id = -1``, diagnostics dict with the reference's keys)."""
    import torch

    required = {"gene_id", "magnitude_mean", "area", "distance_min"}
    missing = sorted(required.difference(loaded.columns))
    if missing:
        raise ValueError("Blank-fraction filtering requires columns: " + ", ".join(missing)
                         + ". Re-decode transcripts with the exact caller.")
    ann = loaded.copy()
    n = len(ann)
    inten = ann["magnitude_mean"].to_numpy(dtype=float)
    area = ann["area"].to_numpy(dtype=float)
    dist = ann["distance_min"].to_numpy(dtype=float)
    blank = blank_mask(ann["gene_id"])
    ann["voxel_intensity"], ann["voxel_number"], ann["vector_distance"], ann["is_blank"] = inten, area, dist, blank
    flat_bins = np.full(n, -1, dtype=np.int64)
    row_frac = np.full(n, np.nan)
    keep = np.zeros(n, dtype=bool)
    diag = _empty_diagnostics(target_gross_misid_rate)

    def finish():
        ann["blank_fraction_bin"], ann["blank_fraction"], ann["blank_fraction_keep"] = flat_bins, row_frac, keep
        out = ann[keep].copy()
        out["cell_id"] = -1
        return out, diag

    if n == 0:
        diag["reason"] = "no_transcripts"
        return finish()
    f32 = [v.astype(np.float32) for v in (inten, area, dist)]  # the reference bins float32 copies (PD:3460-3468)
    valid = np.isfinite(f32[0]) & np.isfinite(f32[1]) & np.isfinite(f32[2])
    if not valid.any():
        diag["reason"] = "no_valid_features"
        return finish()
    if blank_count <= 0:
        keep = valid.copy()
        diag["reason"] = "no_blank_barcodes"
        return finish()
    if not blank[valid].any():
        keep = valid.copy()
        diag["reason"] = "no_blank_transcripts"
        return finish()
    e0, e1, e2 = _edges(inten[valid], area[valid], dist[valid], intensity_bins, voxel_number_bins, vector_distance_bins)
    diag.update(intensity_bins=e0, voxel_number_bins=e1, vector_distance_bins=e2)
    dev = ctx.device
    d_vals = [torch.from_numpy(v).to(dev) for v in f32]
    d_blank = torch.from_numpy(blank.astype(np.uint8)).to(dev)
    d_flat, d_all, d_blank_h = ctx.table_hist3d(d_vals[0], d_vals[1], d_vals[2], d_blank, e0, e1, e2)
    flat32 = d_flat.cpu().numpy()
    all_hist = d_all.cpu().numpy()
    blank_hist = d_blank_h.cpu().numpy()
    in_range = flat32 >= 0
    if not in_range.any():
        diag["reason"] = "no_transcripts_in_histogram_range"
        return finish()
    frac_hist = np.full(all_hist.shape, np.nan, dtype=np.float32)
    nz = all_hist > 0
    frac_hist[nz] = blank_hist[nz] / all_hist[nz]
    flat_bins[in_range] = flat32[in_range]
    row_frac[in_range] = frac_hist.ravel()[flat32[in_range]]
    # threshold sweep straight from the histograms (row masks are only needed for the winner):
    # kept(t) = sum of all_hist over bins with fraction <= t, likewise for the blank rows
    thresholds = np.unique(frac_hist[nz])
    sweep, chosen, achieved, reached = [], np.nan, np.inf, False
    for t in thresholds:
        sel = nz & (frac_hist <= t)
        total_kept, blank_kept = int(all_hist[sel].sum()), int(blank_hist[sel].sum())
        if blank_count <= 0 or barcode_count <= 0 or total_kept == 0:
            g = np.inf
        else:
            g = (blank_kept / float(blank_count)) / (total_kept / float(barcode_count))
        sweep.append({"threshold": float(t), "gross_misid_rate": float(g), "kept_transcripts": total_kept})
        if g <= target_gross_misid_rate:  # ascending sweep: the last threshold meeting the target wins
            chosen, achieved, reached = float(t), float(g), True
    if not sweep:
        diag["reason"] = "no_nonempty_histogram_bins"
        return finish()
    sweep_df = pd.DataFrame(sweep)
    if not reached:
        best = int(sweep_df["gross_misid_rate"].argmin())
        chosen = float(sweep_df.loc[best, "threshold"])
        achieved = float(sweep_df.loc[best, "gross_misid_rate"])
    keep = in_range & (row_frac <= chosen)
    diag.update(chosen_threshold=chosen, achieved_gross_misid_rate=achieved, target_reached=reached,
                all_histogram=all_hist, blank_histogram=blank_hist, blank_fraction_histogram=frac_hist,
                threshold_sweep=sweep_df)
    return finish()


# ------------------------------------------------------------------------- LR filter (host, scikit-learn)
def calculate_lr_fdr(df: pd.DataFrame, threshold: float, blank_count: int, barcode_count: int) -> float:
    """PD:3849-3905: (blank calls / blank codewords) / (coding calls / coding codewords)."""
    blank = blank_mask(df["gene_id"])
    if threshold >= 0:
        positive = df["predicted_probability"].to_numpy() > threshold
        df["prediction"] = positive  # the reference leaves its last prediction column in the frame (PD:3877)
        coding = int(np.count_nonzero(~blank & positive))
        noncoding = int(np.count_nonzero(blank & positive))
    else:
        coding, noncoding = int(np.count_nonzero(~blank)), int(np.count_nonzero(blank))
    if coding > 0:
        return (noncoding / blank_count) / (coding / (barcode_count - blank_count))
    return np.inf


def filter_lr(loaded: pd.DataFrame, is_3d: bool, blank_count: int, barcode_count: int, lr_fdr_target: float = 0.05,
              verbose: int = 0) -> pd.DataFrame:
    """PD:3907-4058: logistic-regression blank classifier, as in the reference a scikit-learn CPU
    step over the (small) table (liblinear, balanced subsample, seeds 42)."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import StandardScaler

    df = loaded.copy()
    df["X"] = ~blank_mask(df["gene_id"])
    columns = ["X", "area", "signal_mean", "s-b_mean", "distance_min", "magnitude_mean",
               "inertia_tensor_eigvals-0", "inertia_tensor_eigvals-1"]
    if is_3d:
        columns.append("inertia_tensor_eigvals-2")
    df_true, df_false = df[df["X"]][columns], df[~df["X"]][columns]
    if len(df_false) <= 1:
        if verbose >= 1:
            print("Insufficient Blank barcodes called for filtering.")
        out = loaded.copy()
        out["cell_id"] = -1
        return out
    combined = pd.concat([df_true.sample(n=len(df_false), random_state=42), df_false])
    x_train, _x_test, y_train, _y_test = train_test_split(combined.drop("X", axis=1), combined["X"], test_size=0.1,
                                                          random_state=42)
    scaler = StandardScaler()
    logistic = LogisticRegression(solver="liblinear", random_state=42)
    logistic.fit(scaler.fit_transform(x_train), y_train)
    df["predicted_probability"] = logistic.predict_proba(scaler.transform(df[columns[1:]]))[:, 1]
    coarse = 0
    for t in np.arange(0, 1, 0.1):
        if calculate_lr_fdr(df, t, blank_count, barcode_count) <= lr_fdr_target:
            coarse = t
            break
    fine = coarse
    for t in np.arange(coarse - 0.1, coarse + 0.1, 0.01):
        if calculate_lr_fdr(df, t, blank_count, barcode_count) <= lr_fdr_target:
            fine = t
            break
    out = df[df["predicted_probability"] > fine].copy()
    out["cell_id"] = -1
    return out


# ------------------------------------------------------------------------- de-duplication
def _coords_dev(ctx: DecodeContext, df: pd.DataFrame):
    import torch

    # pandas 3 hands out read-only views; torch wants writable memory
    zyx = np.array(df[["global_z", "global_y", "global_x"]].to_numpy(dtype=np.float64), order="C", copy=True)
    tile = np.array(df["tile_idx"].to_numpy(dtype=np.int32), copy=True)
    dmin = np.array(df["distance_min"].to_numpy(dtype=np.float64), copy=True)
    if np.isnan(dmin).any():
        raise ValueError("distance_min holds NaN; re-decode transcripts with the exact caller.")
    return (torch.from_numpy(zyx).to(ctx.device), torch.from_numpy(tile).to(ctx.device),
            torch.from_numpy(dmin).to(ctx.device))


def remove_duplicates_in_tile_overlap(ctx: DecodeContext, filtered: pd.DataFrame, radius: float = 0.75):
    """PD:4137-4177.  Returns (frame without the dropped rows, index reset; drop mask)."""
    df = filtered.reset_index(drop=True)
    if len(df) < 2:
        return df, np.zeros(len(df), dtype=bool)
    zyx, tile, dmin = _coords_dev(ctx, df)
    drop = ctx.overlap_duplicates(zyx, tile, dmin, radius).cpu().numpy().astype(bool)
    return df[~drop].reset_index(drop=True), drop


def remove_duplicates_within_tile(ctx: DecodeContext, table: pd.DataFrame, radius_xy: float = 0.1,
                                  radius_z: float = 0.5):
    """PD:4179-4363.  Returns (frame without the dropped rows, index reset; drop mask)."""
    import torch

    df = table.reset_index(drop=True)
    if len(df) < 2:
        return df, np.zeros(len(df), dtype=bool)
    zyx, tile, dmin = _coords_dev(ctx, df)
    gene = pd.factorize(df["gene_id"].to_numpy(), use_na_sentinel=False)[0].astype(np.int32)
    d_gene = torch.from_numpy(np.ascontiguousarray(gene)).to(ctx.device)
    drop = ctx.within_tile_duplicates(zyx, tile, d_gene, dmin, radius_xy, radius_z).cpu().numpy().astype(bool)
    return df[~drop].reset_index(drop=True), drop


# ------------------------------------------------------------------------- cell assignment
def polygon_grid(polygons_yx: list[np.ndarray | None], max_cells_per_axis: int = 2048):
    """Host index for ``m3d_assign_cells``: flattened vertices, offsets, bounding boxes and a uniform grid whose
    cells list (ascending) the polygons whose box touches them.  ``None`` / degenerate entries keep their index
    (cell ids are positions in the ROI archive, PD:4099-4105) but can never contain a point."""
    P = len(polygons_yx)
    counts = np.array([0 if p is None or len(p) < 3 else len(p) for p in polygons_yx], dtype=np.int64)
    offsets = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    verts = np.zeros((max(int(offsets[-1]), 1), 2), dtype=np.float64)
    bbox = np.zeros((max(P, 1), 4), dtype=np.float64)
    bbox[:, :2], bbox[:, 2:] = np.inf, -np.inf  # empty boxes never match
    for i, p in enumerate(polygons_yx):
        if counts[i]:
            q = np.asarray(p, dtype=np.float64)
            verts[offsets[i] : offsets[i + 1]] = q
            bbox[i] = (q[:, 0].min(), q[:, 1].min(), q[:, 0].max(), q[:, 1].max())
    valid = np.flatnonzero(counts > 0)
    if valid.size == 0:
        return dict(verts=verts, offsets=offsets, bbox=bbox, cell_start=np.zeros(2, dtype=np.int32),
                    cell_polys=np.zeros(1, dtype=np.int32), origin=(0.0, 0.0), cell=1.0, grid=(1, 1))
    lo = bbox[valid, :2].min(axis=0)
    hi = bbox[valid, 2:].max(axis=0)
    ext = bbox[valid, 2:] - bbox[valid, :2]
    cell = float(max(np.median(ext), (hi - lo).max() / max_cells_per_axis, 1e-9))
    gy = int(np.floor((hi[0] - lo[0]) / cell)) + 1
    gx = int(np.floor((hi[1] - lo[1]) / cell)) + 1
    c0 = np.floor((bbox[valid, :2] - lo) / cell).astype(np.int64)
    c1 = np.floor((bbox[valid, 2:] - lo) / cell).astype(np.int64)
    c1 = np.minimum(c1, [gy - 1, gx - 1])
    cells, polys = [], []
    for (y0, x0), (y1, x1), p in zip(c0, c1, valid):
        yy, xx = np.meshgrid(np.arange(y0, y1 + 1), np.arange(x0, x1 + 1), indexing="ij")
        c = (yy * gx + xx).ravel()
        cells.append(c)
        polys.append(np.full(c.size, p, dtype=np.int64))
    cells = np.concatenate(cells)
    polys = np.concatenate(polys)
    order = np.lexsort((polys, cells))  # by cell, then ascending polygon index
    cells, polys = cells[order], polys[order]
    cell_start = np.zeros(gy * gx + 1, dtype=np.int64)
    np.add.at(cell_start, cells + 1, 1)
    np.cumsum(cell_start, out=cell_start)
    return dict(verts=verts, offsets=offsets, bbox=bbox, cell_start=cell_start.astype(np.int32),
                cell_polys=polys.astype(np.int32), origin=(float(lo[0]), float(lo[1])), cell=cell, grid=(gy, gx))


def assign_cells(ctx: DecodeContext, filtered: pd.DataFrame, polygons_yx: list[np.ndarray | None]) -> np.ndarray:
    """PD:4107-4135: ``cell_id`` of every transcript = 1 + index of the (first) polygon containing
    (global_y, global_x), 0 when none does."""
    import torch

    if len(filtered) == 0:
        return np.zeros(0, dtype=np.int64)
    g = polygon_grid(polygons_yx)
    dev = ctx.device
    yx = np.array(filtered[["global_y", "global_x"]].to_numpy(dtype=np.float64), order="C", copy=True)
    ids = ctx.assign_cells(
        torch.from_numpy(yx).to(dev), torch.from_numpy(g["verts"]).to(dev), torch.from_numpy(g["offsets"]).to(dev),
        torch.from_numpy(g["bbox"]).to(dev), torch.from_numpy(g["cell_start"]).to(dev),
        torch.from_numpy(g["cell_polys"]).to(dev), g["origin"], g["cell"], g["grid"],
    )
    return ids.cpu().numpy().astype(np.int64)
