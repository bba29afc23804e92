"""Post-decode table stage on the device (csrc/table.cu through ``PixelDecoder``'s reference-named
methods) against golden vectors produced by the reference's own methods, and against the oracle
on larger random tables (ties at the inclusive radius, chains, dense clusters)."""
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

import cases
from oracle import table_oracle as tor
from scenarios import synthetic_transcript_table
from test_cpu_table_stage import GOLDEN, SEED, VOXEL, table_for

pytestmark = pytest.mark.gpu


def _decoder(tmp_path, mode, n_tiles=4):
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, _cb = cases.codebook16()
    ds = ArrayDataStore(tmp_path / f"store_{mode}", codebook=df_cb, voxel_size_zyx_um=VOXEL[mode],
                        microscope_type="3D" if mode == "3d" else "2D")
    for _ in range(n_tiles):
        ds.add_tile(np.zeros((16, 2, 4, 4), dtype=np.uint16))
    return PixelDecoder(ds, merfish_bits=16, verbose=0), ds


@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_table_stage_equals_reference_golden(tmp_path, mode):
    g = np.load(GOLDEN)
    table, _bc, _kc = table_for(mode)
    table["_row"] = np.arange(len(table))
    dec, ds = _decoder(tmp_path, mode)
    dec._df_barcodes_loaded = table.copy()
    dec._filter_all_barcodes_blank_fraction(target_gross_misid_rate=0.05)
    diag = dec._blank_fraction_filter_results
    np.testing.assert_array_equal(diag["all_histogram"], g[f"all_histogram_{mode}"])
    np.testing.assert_array_equal(diag["blank_histogram"], g[f"blank_histogram_{mode}"])
    np.testing.assert_array_equal(diag["blank_fraction_histogram"], g[f"blank_fraction_histogram_{mode}"])
    for k in ("intensity_bins", "voxel_number_bins", "vector_distance_bins"):
        np.testing.assert_array_equal(diag[k], g[f"{k}_{mode}"])
    np.testing.assert_array_equal(diag["threshold_sweep"].to_numpy(float), g[f"sweep_{mode}"])
    assert diag["chosen_threshold"] == float(g[f"chosen_threshold_{mode}"])
    assert diag["achieved_gross_misid_rate"] == float(g[f"achieved_rate_{mode}"])
    assert bool(diag["target_reached"]) == bool(g[f"target_reached_{mode}"])
    np.testing.assert_array_equal(dec._df_filtered_barcodes["_row"].to_numpy(), g[f"kept_filter_{mode}"])
    assert (dec._df_filtered_barcodes["cell_id"] == -1).all()
    if mode == "2d":
        vs = ds.voxel_size_zyx_um
        dec._remove_duplicates_within_tile(radius_xy=float(vs[-1]), radius_z=float(vs[0]))
        np.testing.assert_array_equal(dec._df_filtered_barcodes["_row"].to_numpy(), g["kept_within_2d"])
    dec._remove_duplicates_in_tile_overlap()
    np.testing.assert_array_equal(dec._df_filtered_barcodes["_row"].to_numpy(), g[f"kept_overlap_{mode}"])
    assert list(dec._df_filtered_barcodes.index) == list(range(len(dec._df_filtered_barcodes)))


def test_blank_fraction_edge_cases_on_device(tmp_path):
    table, _b, _k = table_for("3d")
    dec, _ds = _decoder(tmp_path, "3d")
    dec._df_barcodes_loaded = table.iloc[:0].copy()
    dec._filter_all_barcodes_blank_fraction()
    assert dec._blank_fraction_filter_results["reason"] == "no_transcripts" and len(dec._df_filtered_barcodes) == 0
    coding = table[~tor.is_blank(table["gene_id"])].reset_index(drop=True)
    dec._df_barcodes_loaded = coding
    dec._filter_all_barcodes_blank_fraction()
    assert dec._blank_fraction_filter_results["reason"] == "no_blank_transcripts"
    assert len(dec._df_filtered_barcodes) == len(coding)
    with pytest.raises(ValueError, match="requires columns"):
        dec._df_barcodes_loaded = table.drop(columns=["distance_min"])
        dec._filter_all_barcodes_blank_fraction()
    # explicit edges: rows outside them are dropped, NaN features are invalid
    t2 = table.copy()
    t2.loc[:9, "area"] = np.nan
    dec._df_barcodes_loaded = t2
    dec._filter_all_barcodes_blank_fraction(0.05, intensity_bins=[1.0, 2.0, 2.5, 4.0], voxel_number_bins=[0, 10, 30, 100],
                                            vector_distance_bins=[0.0, 0.2, 0.4, 0.7])
    keep, diag = tor.blank_fraction_filter(t2, dec._blank_count, dec._barcode_count, 0.05, [1.0, 2.0, 2.5, 4.0],
                                           [0, 10, 30, 100], [0.0, 0.2, 0.4, 0.7])
    np.testing.assert_array_equal(dec._blank_fraction_filter_results["all_histogram"], diag["all_histogram"])
    np.testing.assert_array_equal(dec._df_filtered_barcodes.index.to_numpy(), np.flatnonzero(keep))
    with pytest.raises(ValueError, match="filter_method"):
        dec._apply_filter_method("nope", 0.05, 0.05)


@pytest.mark.parametrize("seed", [1, 2])
def test_overlap_duplicates_random_with_radius_ties(seed):
    """Coordinates on a 0.25 um lattice: many pairs sit EXACTLY at the 0.75 um radius (inclusive) and
    many share distance_min, so the (distance_min, row) tie-break and the float64 d^2 <= r^2 rule are
    both exercised against the cKDTree oracle."""
    import torch

    from merfish3d_analysis_b200._capi import DecodeContext

    rng = np.random.default_rng(seed)
    n = 40000
    zyx = np.round(rng.integers(0, 60, (n, 3)) * 0.25 + 1000.0 * rng.integers(0, 2, (n, 1)), 2)
    tile = rng.integers(0, 5, n).astype(np.int32)
    dmin = rng.integers(0, 8, n) / 16.0
    ref = tor.overlap_duplicates(zyx, tile, dmin, 0.75)
    assert 0.05 < ref.mean() < 0.999
    ctx = DecodeContext(np.eye(4, dtype=np.float32)[:2], (), device=0)
    got = ctx.overlap_duplicates(torch.from_numpy(zyx).cuda(), torch.from_numpy(tile).cuda(),
                                 torch.from_numpy(dmin).cuda(), 0.75).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(got, ref)
    ctx.close()


@pytest.mark.parametrize("seed", [3, 4])
def test_within_tile_duplicates_random_chains(seed):
    import torch

    from merfish3d_analysis_b200._capi import DecodeContext

    rng = np.random.default_rng(seed)
    n = 30000
    zyx = np.empty((n, 3))
    zyx[:, 0] = rng.integers(0, 6, n) * 1.5
    zyx[:, 1:] = np.round(rng.integers(0, 400, (n, 2)) * 0.05, 2)  # 0.05 um lattice vs radius 0.1: exact ties
    tile = rng.integers(0, 3, n).astype(np.int32)
    gene = rng.integers(0, 4, n).astype(np.int32)
    dmin = rng.integers(0, 6, n) / 8.0
    ref = tor.within_tile_duplicates(zyx, tile, gene, dmin, 0.1, 1.5)
    assert 0.02 < ref.mean() < 0.9
    ctx = DecodeContext(np.eye(4, dtype=np.float32)[:2], (), device=0)
    got = ctx.within_tile_duplicates(torch.from_numpy(zyx).cuda(), torch.from_numpy(tile).cuda(),
                                     torch.from_numpy(gene).cuda(), torch.from_numpy(dmin).cuda(), 0.1, 1.5)
    np.testing.assert_array_equal(got.cpu().numpy().astype(bool), ref)
    ctx.close()


def test_decode_all_tiles_runs_the_table_stage(tmp_path):
    """decode_all_tiles end to end on overlapping tiles: per-tile parquet files, then the pooled
    filter + overlap de-duplication, written to all_tiles_filtered_decoded_features/."""
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    base = cases.small_stack(cb["matrix"], shape=(10, 48, 64), seed=61, density=5e-3)
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb)
    # two tiles imaging the same field (second shifted by 3 um in x): every molecule in the overlap is seen twice
    ds.add_tile(base, stage_origin_zyx_um=(0.0, 0.0, 0.0))
    ds.add_tile(base, stage_origin_zyx_um=(0.0, 0.0, 0.2))
    bkg, nrm = cases.simple_vectors(16)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec.decode_all_tiles(assign_to_cells=True, lowpass_sigma=None, minimum_pixels=4, normalization_method="global")
    per_tile = [ds.load_local_decoded_spots(t) for t in range(2)]
    assert all(p is not None and len(p) > 10 for p in per_tile)
    out = ds.load_global_filtered_decoded_spots()
    pooled = pd.concat(per_tile, ignore_index=True)
    keep, _diag = tor.blank_fraction_filter(pooled, dec._blank_count, dec._barcode_count, 0.05)
    cur = pooled[keep].reset_index(drop=True)
    drop = tor.overlap_duplicates(cur[["global_z", "global_y", "global_x"]].to_numpy(float), cur["tile_idx"].to_numpy(),
                                  cur["distance_min"].to_numpy(float), 0.75)
    want = cur[~drop].reset_index(drop=True)
    assert drop.sum() > 5 and len(out) == len(want)
    np.testing.assert_array_equal(out["global_x"].to_numpy(), want["global_x"].to_numpy())
    assert out["gene_id"].tolist() == want["gene_id"].tolist()
    assert (out["cell_id"] == -1).all()
    # re-filtering from the saved per-tile tables gives the same result
    dec2 = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec2.optimize_filtering(assign_to_cells=False)
    out2 = ds.load_global_filtered_decoded_spots()
    assert len(out2) == len(out)


def test_assign_cells_from_imagej_rois(tmp_path):
    """decode_all_tiles' last step: ImageJ ROI archive -> polygons -> device point-in-polygon, against the
    oracle's even-odd restatement (shapely / rtree are not installed: unpinned), incl. concave and overlapping
    cells (lowest index wins), an unusable ROI in the archive and points far outside every cell."""
    import zipfile

    from merfish3d_analysis_b200 import roi

    rng = np.random.default_rng(8)
    polys_xy = []
    for i in range(400):  # star-shaped (concave) cells on a jittered lattice, some overlapping
        cy, cx = (i // 20) * 30.0 + rng.uniform(0, 12), (i % 20) * 30.0 + rng.uniform(0, 12)
        ang = np.sort(rng.uniform(0, 2 * np.pi, rng.integers(8, 40)))
        rad = rng.uniform(6, 20, ang.size)
        polys_xy.append(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1))
    dec, ds = _decoder(tmp_path, "3d")
    roi_dir = Path(ds._datastore_path) / "segmentation" / "cellpose" / "imagej_rois"
    roi_dir.mkdir(parents=True)
    roi.write_roi_zip(roi_dir / "global_coords_rois.zip", polys_xy)
    with zipfile.ZipFile(roi_dir / "global_coords_rois.zip", "a") as zf:
        zf.writestr("zzz_line.roi", b"Iout" + bytes(60))  # an ROI without an outline: skipped, ids unchanged
    n = 50000
    df = pd.DataFrame({"global_y": np.round(rng.uniform(-50, 650, n), 2), "global_x": np.round(rng.uniform(-50, 650, n), 2),
                       "gene_id": "g", "cell_id": -1})
    dec._df_filtered_barcodes = df.copy()
    dec._assign_cells()
    got = dec._df_filtered_barcodes["cell_id"].to_numpy()
    # the archive stores float32 vertices: the oracle sees what the reader returns
    polys_yx = [p[:, ::-1] for p in roi.read_roi_zip(roi_dir / "global_coords_rois.zip") if p is not None]
    assert len(polys_yx) == 400
    want = tor.assign_cells(df[["global_y", "global_x"]].to_numpy(float), polys_yx)
    np.testing.assert_array_equal(got, want)
    assert 0.2 < (got > 0).mean() < 0.9 and got.max() <= 400
    # no archive: reported and skipped like the reference
    dec2, _ds2 = _decoder(tmp_path / "b", "3d")
    dec2._df_filtered_barcodes = df.copy()
    dec2._assign_cells()
    assert (dec2._df_filtered_barcodes["cell_id"] == -1).all()
