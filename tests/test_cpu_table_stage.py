"""Post-decode table stage: the oracle (``oracle/table_oracle.py``) against golden
vectors produced by the reference's own ``_filter_all_barcodes_blank_fraction`` /
``_remove_duplicates_within_tile`` / ``_remove_duplicates_in_tile_overlap`` (see
``tests/golden/make_table_golden.py``)."""
from pathlib import Path

import numpy as np
import pytest

import cases
from oracle import table_oracle as tor
from scenarios import synthetic_transcript_table

GOLDEN = Path(__file__).resolve().parent / "golden" / "reference_table_stage.npz"
VOXEL = {"3d": (0.315, 0.098, 0.098), "2d": (1.5, 0.1085, 0.1085)}
SEED = {"3d": 5150, "2d": 5151}


def table_for(mode):
    df_cb, cb = cases.codebook16()
    table = synthetic_transcript_table(df_cb, seed=SEED[mode], mode=mode)
    n_blank = sum(str(g).lower().startswith("blank") for g in cb["gene_ids"])
    return table, n_blank, len(cb["gene_ids"])


def oracle_chain(table, mode, blank_count, barcode_count):
    """filter -> (2-D: within-tile collapse) -> tile-overlap de-duplication, as decode_all_tiles does
    (PD:4849-4868).  Returns the surviving original row numbers after each step + diagnostics."""
    rows = np.arange(len(table))
    keep, diag = tor.blank_fraction_filter(table, blank_count, barcode_count, 0.05)
    cur = table[keep].reset_index(drop=True)
    rows = rows[keep]
    steps = {"filter": rows.copy()}
    if mode == "2d":
        vs = VOXEL[mode]
        drop = tor.within_tile_duplicates(cur[["global_z", "global_y", "global_x"]].to_numpy(float),
                                          cur["tile_idx"].to_numpy(), cur["gene_id"].to_numpy(),
                                          cur["distance_min"].to_numpy(float), vs[-1], vs[0])
        cur, rows = cur[~drop].reset_index(drop=True), rows[~drop]
        steps["within"] = rows.copy()
    drop = tor.overlap_duplicates(cur[["global_z", "global_y", "global_x"]].to_numpy(float),
                                  cur["tile_idx"].to_numpy(), cur["distance_min"].to_numpy(float), 0.75)
    steps["overlap"] = rows[~drop]
    return steps, diag


@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_table_oracle_equals_reference_golden(mode):
    g = np.load(GOLDEN)
    table, blank_count, barcode_count = table_for(mode)
    steps, diag = oracle_chain(table, mode, blank_count, barcode_count)
    np.testing.assert_array_equal(diag["all_histogram"], g[f"all_histogram_{mode}"])
    np.testing.assert_array_equal(diag["blank_histogram"], g[f"blank_histogram_{mode}"])
    np.testing.assert_array_equal(diag["blank_fraction_histogram"], g[f"blank_fraction_histogram_{mode}"])
    for k in ("intensity_bins", "voxel_number_bins", "vector_distance_bins"):
        np.testing.assert_array_equal(diag[k], g[f"{k}_{mode}"])
    np.testing.assert_array_equal(diag["threshold_sweep"].to_numpy(float), g[f"sweep_{mode}"])
    assert diag["chosen_threshold"] == float(g[f"chosen_threshold_{mode}"])
    assert diag["achieved_gross_misid_rate"] == float(g[f"achieved_rate_{mode}"])
    assert bool(diag["target_reached"]) == bool(g[f"target_reached_{mode}"])
    np.testing.assert_array_equal(steps["filter"], g[f"kept_filter_{mode}"])
    if mode == "2d":
        assert len(g["kept_within_2d"]) < len(g["kept_filter_2d"])
        np.testing.assert_array_equal(steps["within"], g["kept_within_2d"])
    assert len(g[f"kept_overlap_{mode}"]) < len(steps["filter"])
    np.testing.assert_array_equal(steps["overlap"], g[f"kept_overlap_{mode}"])


def test_blank_fraction_edge_cases():
    table, blank_count, barcode_count = table_for("3d")
    keep, diag = tor.blank_fraction_filter(table.iloc[:0], blank_count, barcode_count)
    assert diag["reason"] == "no_transcripts" and keep.size == 0
    keep, diag = tor.blank_fraction_filter(table, 0, barcode_count)
    assert diag["reason"] == "no_blank_barcodes" and keep.all()
    coding = table[~tor.is_blank(table["gene_id"])].reset_index(drop=True)
    keep, diag = tor.blank_fraction_filter(coding, blank_count, barcode_count)
    assert diag["reason"] == "no_blank_transcripts" and keep.all()
    bad = table.copy()
    bad["area"] = np.nan
    keep, diag = tor.blank_fraction_filter(bad, blank_count, barcode_count)
    assert diag["reason"] == "no_valid_features" and not keep.any()


@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_lr_filter_equals_reference_golden(mode):
    """``_filter_all_barcodes_LR`` is host-side scikit-learn in the reference as well: the product's mirror
    (``table_stage.filter_lr``) must keep the same rows with the same probabilities."""
    from merfish3d_analysis_b200 import table_stage

    g = np.load(GOLDEN)
    table, blank_count, barcode_count = table_for(mode)
    table["_row"] = np.arange(len(table))
    out = table_stage.filter_lr(table, mode == "3d", blank_count, barcode_count, 0.05)
    np.testing.assert_array_equal(out["_row"].to_numpy(), g[f"kept_lr_{mode}"])
    np.testing.assert_allclose(out["predicted_probability"].to_numpy(dtype=float), g[f"lr_probability_{mode}"],
                               rtol=1e-12)
    assert (out["cell_id"] == -1).all() and 0 < len(out) < len(table)


def test_histogram_edges_host_mirror_equals_oracle():
    """The product computes the blank-fraction histogram edges on the host (O(bins) work): same edges as the
    oracle's restatement of PD:3494-3655, including the degenerate single-value axes."""
    from merfish3d_analysis_b200 import table_stage

    table, _b, _k = table_for("3d")
    inten = table["magnitude_mean"].to_numpy(float)
    area = table["area"].to_numpy(float)
    dist = table["distance_min"].to_numpy(float)
    e0, e1, e2 = table_stage._edges(inten, area, dist, None, None, None)
    np.testing.assert_array_equal(e0, tor.intensity_edges(inten))
    np.testing.assert_array_equal(e1, tor.voxel_number_edges(area))
    np.testing.assert_array_equal(e2, tor.vector_distance_edges(dist))
    small = area[:50].clip(5, 9)  # <= 10 distinct integer areas -> unit bins
    flat = np.full(30, 0.25)
    e0, e1, e2 = table_stage._edges(flat, small, flat, None, None, None)
    np.testing.assert_array_equal(e0, tor.intensity_edges(flat))
    np.testing.assert_array_equal(e1, tor.voxel_number_edges(small))
    np.testing.assert_array_equal(e2, tor.vector_distance_edges(flat))
    with pytest.raises(ValueError, match="at least two finite"):
        table_stage._edges(inten, area, dist, [1.0], None, None)


def test_imagej_roi_round_trip_and_grid_index(tmp_path):
    """ROI archive reader (published ImageJ layout) and the host grid index behind m3d_assign_cells: every
    (point, containing polygon) pair of the oracle must be reachable through the point's grid cell."""
    import zipfile

    from merfish3d_analysis_b200 import roi, table_stage

    rng = np.random.default_rng(4)
    polys_xy = []
    for i in range(60):
        cy, cx = rng.uniform(0, 300, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, rng.integers(3, 30)))
        rad = rng.uniform(3, 25, ang.size)
        polys_xy.append(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1))
    path = tmp_path / "rois.zip"
    roi.write_roi_zip(path, polys_xy)
    with zipfile.ZipFile(path, "a") as zf:
        zf.writestr("junk.roi", b"not an roi")
        zf.writestr("readme.txt", b"ignored")
    back = roi.read_roi_zip(path)
    assert len(back) == 61 and back[-1] is None
    for a, b in zip(polys_xy, back):
        np.testing.assert_allclose(b, a.astype(np.float32).astype(np.float64), rtol=0, atol=0)
    # integer-only ROI (no sub-pixel block): vertices = int16 offsets + bounding-box corner
    raw = bytearray(roi.encode_polygon_roi(np.array([[10.0, 20.0], [40.0, 20.0], [40.0, 60.0]])))
    raw[50:52] = (0).to_bytes(2, "big")
    np.testing.assert_array_equal(roi.parse_roi(bytes(raw[: 64 + 4 * 3])), [[10, 20], [40, 20], [40, 60]])
    polys_yx = [p[:, ::-1] for p in back if p is not None]
    pts = np.round(rng.uniform(-20, 340, (4000, 2)), 2)
    want = tor.assign_cells(pts, polys_yx)
    g = table_stage.polygon_grid(polys_yx)
    gy, gx = g["grid"]
    cy = np.floor((pts[:, 0] - g["origin"][0]) / g["cell"]).astype(int)
    cx = np.floor((pts[:, 1] - g["origin"][1]) / g["cell"]).astype(int)
    hits = 0
    for i in np.flatnonzero(want > 0):
        assert 0 <= cy[i] < gy and 0 <= cx[i] < gx
        c = cy[i] * gx + cx[i]
        cand = g["cell_polys"][g["cell_start"][c] : g["cell_start"][c + 1]]
        assert want[i] - 1 in cand and np.all(np.diff(cand) > 0)  # reachable, candidates ascending
        hits += 1
    assert hits > 300


def test_assign_cells_containment_agrees_with_opencv():
    """shapely / rtree are not installed, so `assign_cells` (PD:4107-4135) cannot be pinned to them; OpenCV's
    pointPolygonTest is an independent point-in-polygon implementation that is: for simple polygons (cell outlines)
    interior membership is the same predicate as shapely's `contains` away from the boundary."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(17)
    polys = []
    for _ in range(40):  # star-shaped outlines: simple, non-convex, overlapping each other
        cy, cx = rng.uniform(100, 900, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, int(rng.integers(5, 40))))
        rad = rng.uniform(20, 120, ang.size)
        polys.append(np.stack([cy + rad * np.sin(ang), cx + rad * np.cos(ang)], axis=1))
    pts = rng.uniform(0, 1000, (20000, 2))
    contours = [np.ascontiguousarray(p[:, ::-1], dtype=np.float32).reshape(-1, 1, 2) for p in polys]  # (x, y)
    dist = np.array([[cv2.pointPolygonTest(c, (float(x), float(y)), True) for c in contours] for y, x in pts])
    clear = np.all(np.abs(dist) > 1e-2, axis=1)  # float32 contours: stay away from the outlines
    assert clear.sum() > 19000
    inside = dist > 0
    expected = np.where(inside.any(axis=1), inside.argmax(axis=1) + 1, 0)  # lowest-numbered containing polygon
    got = tor.assign_cells(pts, polys)
    np.testing.assert_array_equal(got[clear], expected[clear])
    assert (got[clear] > 0).sum() > 2000
