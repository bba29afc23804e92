"""GPU parity tests through the reference-facing API: ``PixelDecoder`` (C ABI underneath)
against the oracle's ``decode_tile`` / ``optimize_normalization`` on the same inputs."""
import numpy as np
import pandas as pd
import pytest

import cases
from oracle import decode_oracle as orc

pytestmark = pytest.mark.gpu

REL = 1e-5


def _store(tmp_path, df_cb, stacks, predictors=None, microscope_type="3D", **tile_kw):
    """array-fed datastore; extra keywords go to add_tile (stage / round transform metadata)."""
    from merfish3d_analysis_b200.datastore import ArrayDataStore

    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb, voxel_size_zyx_um=(0.315, 0.098, 0.098),
                        microscope_type=microscope_type)
    for i, st in enumerate(stacks):
        ds.add_tile(st, None if predictors is None else predictors[i], **tile_kw)
    return ds


def _compare_tables(got: pd.DataFrame, ref: pd.DataFrame, n_bits=16):
    assert list(got.columns) == list(ref.columns)
    assert len(got) == len(ref)
    exact = ["area", "barcode_id", "gene_id", "tile_idx", "on_bit_1", "on_bit_2", "on_bit_3", "on_bit_4",
             "tile_z", "tile_y", "tile_x"]
    for c in exact:
        assert got[c].tolist() == ref[c].tolist(), c
    close = ["z", "y", "x", "distance_min", "magnitude_mean", "global_z", "global_y", "global_x",
             "signal_mean", "bkd_mean", "s-b_mean"] + [f"bit{i:02d}_mean_intensity" for i in range(1, n_bits + 1)]
    for c in close:
        np.testing.assert_allclose(
            got[c].to_numpy(dtype=np.float64), ref[c].to_numpy(dtype=np.float64), rtol=REL, atol=1e-7, err_msg=c
        )
    for k in range(3):
        c = f"inertia_tensor_eigvals-{k}"
        np.testing.assert_allclose(got[c].to_numpy(dtype=np.float64), ref[c].to_numpy(dtype=np.float64),
                                   rtol=1e-7, atol=1e-9, err_msg=c)


@pytest.mark.parametrize("lowpass", [None, (3.0, 1.0, 1.0)])
@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_decode_one_tile_matches_oracle(tmp_path, lowpass, mode):
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(14, 48, 64), seed=29, density=3e-3)
    bkg, nrm = cases.simple_vectors(16, nrm=300.0 if lowpass else 900.0)
    origin = np.array([100.0, 20.0, -3.0], dtype=np.float32)
    cam = np.array([[1, 0, 0, 0], [0, -0.07, -1, 0], [0, -1, 0.07, 0], [0, 0, 0, 1]], dtype=np.float32)
    ds = _store(tmp_path, df_cb, [stack], stage_origin_zyx_um=origin, camera_to_stage_affine=cam)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0, decode_mode=mode)
    out = dec.decode_one_tile(0, lowpass_sigma=lowpass, minimum_pixels=4, normalization_method="global",
                              return_results=True)
    got = dec.decoded_barcodes
    ref, imgs = orc.decode_tile(
        stack, None, cb, bkg, nrm, is_3d=(mode == "3d"), lowpass_sigma=lowpass, minimum_pixels=4,
        spacing=ds.voxel_size_zyx_um, origin=origin, camera_to_stage=cam,
    )
    assert len(ref) > 10
    image, scaled, magnitude, distance, decoded = out
    np.testing.assert_array_equal(decoded, imgs["decoded"])
    np.testing.assert_array_equal(image, imgs["image"].astype(np.float32))
    np.testing.assert_array_equal(scaled, imgs["scaled"])
    np.testing.assert_array_equal(magnitude, imgs["magnitude"])
    np.testing.assert_array_equal(distance, imgs["distance"])
    _compare_tables(got, ref)
    # production path (no result images) must give the same table
    dec2 = PixelDecoder(ds, merfish_bits=16, verbose=0, decode_mode=mode)
    assert dec2.decode_one_tile(0, lowpass_sigma=lowpass, minimum_pixels=4, normalization_method="global") is None
    pd.testing.assert_frame_equal(dec2.decoded_barcodes, got)
    np.testing.assert_array_equal(dec2.decoded_image, decoded)


def test_decode_with_predictor_zcrop_and_exclusions(tmp_path):
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(16, 40, 56), seed=31, density=3e-3)
    rng = np.random.default_rng(5)
    pred = (rng.uniform(size=stack.shape) > 0.05).astype(np.float32) * rng.uniform(0.9, 1.0, size=stack.shape).astype(np.float32)
    bkg, nrm = cases.simple_vectors(16)
    ds = _store(tmp_path, df_cb, [stack], [pred])
    ds.save_decode_normalization_vectors(None, "iterative", nrm, bkg)
    excluded_gene = "gene0003"
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0, z_range=(2, 14), excluded_gene_ids=[excluded_gene])
    dec.decode_one_tile(0, lowpass_sigma=None, minimum_pixels=4)
    got = dec.decoded_barcodes
    ref, _ = orc.decode_tile(
        stack[:, 2:14], pred[:, 2:14], cb, bkg, nrm, lowpass_sigma=None, minimum_pixels=4,
        excluded=(cb["gene_ids"].index(excluded_gene),), spacing=ds.voxel_size_zyx_um, z_offset=2.0,
    )
    assert len(ref) > 5 and excluded_gene not in set(ref["gene_id"])
    _compare_tables(got, ref)


def test_decode_api_errors(tmp_path):
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(1, 16, 16), seed=1)
    ds = _store(tmp_path, df_cb, [stack])
    with pytest.raises(ValueError, match="decode_mode"):
        PixelDecoder(ds, decode_mode="4d")
    dec = PixelDecoder(ds, verbose=0, decode_mode="3d")
    with pytest.raises(ValueError, match="two z planes"):
        dec.decode_one_tile(0, normalization_method="none")
    auto = PixelDecoder(ds, verbose=0)  # 'auto' does not enforce the two-plane rule (PD:1895)
    with pytest.raises(ValueError, match="three values"):
        auto.decode_one_tile(0, normalization_method="none", lowpass_sigma=(1, 1))
    with pytest.raises(ValueError, match="normalization_method"):
        auto.decode_one_tile(0, normalization_method="bogus")
    assert auto.decode_one_tile(0, normalization_method="none", lowpass_sigma=None) is None
    assert len(auto.decoded_barcodes) == 0 and list(auto.decoded_barcodes.columns)[:4] == ["area", "z", "y", "x"]


def test_optimize_normalization_matches_oracle(tmp_path):
    """global percentile seed + 3 iterations of decode -> per-bit medians (sim-like config)."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stacks = [cases.small_stack(cb["matrix"], shape=(12, 48, 64), seed=37 + i, density=4e-3) for i in range(2)]
    ds = _store(tmp_path, df_cb, stacks)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    sigma = (1.0, 0.5, 0.5)
    dec.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=sigma,
                                           magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1])
    ref = orc.optimize_normalization([(s, None) for s in stacks], cb, 3, True, sigma, (0.9, 10.0), 4)
    g_n, g_b = ds.load_decode_normalization_vectors(None, "global")
    np.testing.assert_array_equal(g_n, ref["global_"][0])
    np.testing.assert_array_equal(g_b, ref["global_"][1])
    i_n, i_b = ds.load_decode_normalization_vectors(None, "iterative")
    assert ref["history"][-1][2] > 20
    np.testing.assert_array_equal(i_n, ref["iterative"][0])
    np.testing.assert_array_equal(i_b, ref["iterative"][1])
    md = ds.load_decode_normalization_metadata(None, "iterative")
    assert md["codebook_sha256"] == orc.codebook_fingerprint(cb["matrix"], cb["gene_ids"])
    # a decoder with a different codebook must refuse the cached vectors (PD:1220-1230)
    df2 = df_cb.iloc[:-1].reset_index(drop=True)
    from merfish3d_analysis_b200.datastore import ArrayDataStore

    ds2 = ArrayDataStore(tmp_path / "qi2labdatastore")
    ds2._codebook = df2
    dec2 = PixelDecoder(ds2, merfish_bits=16, verbose=0)
    ds2.add_tile(stacks[0])
    with pytest.raises(ValueError, match="different active codebook"):
        dec2.decode_one_tile(0, normalization_method="iterative")


def test_decode_all_tiles_writes_reference_layout(tmp_path):
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stacks = [cases.small_stack(cb["matrix"], shape=(8, 32, 48), seed=51 + i, density=4e-3) for i in range(3)]
    ds = _store(tmp_path, df_cb, stacks)
    bkg, nrm = cases.simple_vectors(16)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    ds.save_decode_normalization_vectors(None, "iterative", nrm, bkg)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec.decode_all_tiles(lowpass_sigma=None, minimum_pixels=4)
    for i, tile_id in enumerate(ds.tile_ids):
        p = tmp_path / "qi2labdatastore" / "decoded" / f"{tile_id}_decoded_features.parquet"
        assert p.exists()
        got = pd.read_parquet(p)
        ref, _ = orc.decode_tile(stacks[i], None, cb, bkg, nrm, lowpass_sigma=None, minimum_pixels=4,
                                 tile_idx=i, spacing=ds.voxel_size_zyx_um)
        _compare_tables(got, ref)
    assert len(dec._df_barcodes_loaded) == sum(len(pd.read_parquet(p)) for p in (tmp_path / "qi2labdatastore" / "decoded").glob("*.parquet"))


@pytest.mark.parametrize("lowpass", [None, (3.0, 1.0, 1.0)])
@pytest.mark.parametrize("n_slabs", [2, 3, 5])
def test_z_slab_sharding_equals_unsharded(tmp_path, lowpass, n_slabs):
    """z-slab sharding (here: sequential slabs on one GPU) == the unsharded volume,
    including components cut by an interface, merged size filters and oversized components."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(30, 40, 56), seed=83, density=4e-3)
    # a long uniform column crossing every interface (> 500 voxels -> dropped everywhere) and a
    # thin one that is only large enough once its pieces are merged
    on = np.flatnonzero(cb["matrix"][7])
    stack[:, :, 2:8, 2:8] = 210
    stack[np.ix_(on, np.arange(30), np.arange(2, 8), np.arange(2, 8))] = 2000
    on2 = np.flatnonzero(cb["matrix"][21])
    stack[:, 3:27, 30, 40] = 205
    stack[np.ix_(on2, np.arange(3, 27), [30], [40])] = 1900
    bkg, nrm = cases.simple_vectors(16, nrm=300.0 if lowpass else 900.0)
    ds = _store(tmp_path, df_cb, [stack])
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    ref_dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    ref_dec.decode_one_tile(0, lowpass_sigma=lowpass, minimum_pixels=12, normalization_method="global")
    ref = ref_dec.decoded_barcodes
    ref_img = ref_dec.decoded_image
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec.decode_one_tile_sharded(0, n_slabs=n_slabs, lowpass_sigma=lowpass, minimum_pixels=12,
                                normalization_method="global")
    got = dec.decoded_barcodes
    np.testing.assert_array_equal(dec.decoded_image, ref_img)
    assert len(ref) > 10
    pd.testing.assert_frame_equal(got, ref)  # bit-identical table, same row order
    if lowpass is None:
        assert (ref["area"] == 24).any()  # the thin column survives only as a merged component


@pytest.mark.parametrize("lowpass", [None, (3.0, 1.0, 1.0)])
def test_decode_with_round_transforms_matches_oracle(tmp_path, lowpass):
    """bits of later rounds are resampled into the round-1 frame on load (affine, order 1)."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(14, 48, 64), seed=91, density=3e-3)
    rng = np.random.default_rng(7)
    bit_round = [1 + (b // 2) % 4 for b in range(16)]  # rounds 1..4, two bits each, repeating
    xfs = {}
    for r in (2, 3, 4):
        xf = np.eye(4, dtype=np.float32)
        xf[:3, :3] += rng.normal(0, 0.002, (3, 3)).astype(np.float32)
        xf[:3, 3] = (rng.uniform(-1.0, 1.0, 3) * np.array([0.315, 0.098, 0.098])).astype(np.float32)
        xfs[r] = xf
    pred = rng.uniform(0.9, 1.0, size=stack.shape).astype(np.float32)
    ds = _store(tmp_path, df_cb, [stack], [pred], bit_round=bit_round, round_transforms_zyx_um=xfs)
    bkg, nrm = cases.simple_vectors(16, nrm=300.0 if lowpass else 800.0)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0, z_range=(1, 13))
    out = dec.decode_one_tile(0, lowpass_sigma=lowpass, minimum_pixels=4, normalization_method="global",
                              return_results=True)
    per_bit = [None if r == 1 else xfs[r] for r in bit_round]
    # the reference warps the full volume, then crops z (PD:1890)
    w = orc.weight_readout(stack, pred)
    w = np.stack([orc.warp_to_reference(w[b], per_bit[b], ds.voxel_size_zyx_um) for b in range(16)])[:, 1:13]
    ref, imgs = orc.decode_tile(w, None, cb, bkg, nrm, lowpass_sigma=lowpass, minimum_pixels=4,
                                spacing=ds.voxel_size_zyx_um, z_offset=1.0)
    np.testing.assert_array_equal(out[0], imgs["image"])
    np.testing.assert_array_equal(out[4], imgs["decoded"])
    assert len(ref) > 5
    _compare_tables(dec.decoded_barcodes, ref)
    # sharded decode of warped data gives the same table
    dec2 = PixelDecoder(ds, merfish_bits=16, verbose=0, z_range=(1, 13))
    dec2.decode_one_tile_sharded(0, n_slabs=3, lowpass_sigma=lowpass, minimum_pixels=4, normalization_method="global")
    dec3 = PixelDecoder(ds, merfish_bits=16, verbose=0, z_range=(1, 13))
    dec3.decode_one_tile(0, lowpass_sigma=lowpass, minimum_pixels=4, normalization_method="global")
    pd.testing.assert_frame_equal(dec2.decoded_barcodes, dec3.decoded_barcodes)


@pytest.mark.parametrize("lowpass", [None, (3.0, 1.0, 1.0)])
def test_multi_tile_pipeline_equals_tile_by_tile(tmp_path, lowpass):
    """decode_all_tiles stages tile t+1 on a side stream (prefetch thread, persistent double buffers, the
    per-bit low-pass behind the upload) while tile t is finished: every per-tile table must equal a plain
    decode_one_tile of that tile, also when the whole run is repeated on the same decoder."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    rng = np.random.default_rng(5)
    stacks = [cases.small_stack(cb["matrix"], shape=(9, 40, 56), seed=400 + i, density=5e-3) for i in range(4)]
    preds = [None, rng.uniform(0.5, 1.0, stacks[1].shape).astype(np.float32), None, None]
    ds = _store(tmp_path, df_cb, [stacks[0]])
    for st, pr in zip(stacks[1:], preds[1:]):
        ds.add_tile(st, pr)
    bkg, nrm = cases.simple_vectors(16, nrm=300.0 if lowpass else 900.0)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    kw = dict(lowpass_sigma=lowpass, minimum_pixels=4, normalization_method="global")
    single = []
    for t in range(4):
        one = PixelDecoder(ds, merfish_bits=16, verbose=0)
        one.decode_one_tile(t, **kw)
        single.append(one.decoded_barcodes)
    assert sum(len(s) for s in single) > 40
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    for _rep in range(2):
        dec.decode_all_tiles(assign_to_cells=False, **kw)
        for t in range(4):
            got = ds.load_local_decoded_spots(t)
            pd.testing.assert_frame_equal(got.reset_index(drop=True), single[t].reset_index(drop=True),
                                          check_dtype=False)


@pytest.mark.parametrize("lowpass", [None, (3.0, 1.0, 1.0)])
def test_multi_tile_pipeline_on_an_unregistered_store(tmp_path, lowpass):
    """Tiles whose later-round bits need the decode-time warp (the usual case) are prefetched too: tile t+1 is
    uploaded, warped and low-passed on the side stream while tile t is finished.  Per-tile tables must equal plain
    decode_one_tile calls; the optimiser (tile cache on and off) must give the same vectors either way."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    rng = np.random.default_rng(17)
    bit_round = [1 + (b // 2) % 4 for b in range(16)]
    xfs = {}
    for r in (2, 3, 4):
        xf = np.eye(4, dtype=np.float32)
        xf[:3, :3] += rng.normal(0, 0.002, (3, 3)).astype(np.float32)
        xf[:3, 3] = (rng.uniform(-1.0, 1.0, 3) * np.array([0.315, 0.098, 0.098])).astype(np.float32)
        xfs[r] = xf
    stacks = [cases.small_stack(cb["matrix"], shape=(12, 40, 56), seed=600 + i, density=5e-3) for i in range(4)]
    preds = [None, rng.uniform(0.8, 1.0, stacks[1].shape).astype(np.float32), None, None]
    ds = _store(tmp_path, df_cb, stacks, preds, bit_round=bit_round, round_transforms_zyx_um=xfs)
    bkg, nrm = cases.simple_vectors(16, nrm=300.0 if lowpass else 800.0)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    kw = dict(lowpass_sigma=lowpass, minimum_pixels=4, normalization_method="global")
    single = []
    for t in range(4):
        one = PixelDecoder(ds, merfish_bits=16, verbose=0)
        one.decode_one_tile(t, **kw)
        single.append(one.decoded_barcodes)
    assert sum(len(s) for s in single) > 20
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    scheduled = []
    orig = dec._schedule_prefetch

    def spy(tile_idx, *a, **k):
        scheduled.append(tile_idx)
        return orig(tile_idx, *a, **k)

    dec._schedule_prefetch = spy
    dec.decode_all_tiles(assign_to_cells=False, **kw)
    assert scheduled == [1, 2, 3]  # every following tile was staged ahead
    for t in range(4):
        got = ds.load_local_decoded_spots(t)
        pd.testing.assert_frame_equal(got.reset_index(drop=True), single[t].reset_index(drop=True), check_dtype=False)
    vectors = []
    for budget in (None, 0):
        d = PixelDecoder(ds, merfish_bits=16, verbose=0)
        d.tile_cache_budget_bytes = budget
        d.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=lowpass,
                                             magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2, 3])
        vectors.append(ds.load_decode_normalization_vectors(None, "iterative"))
        assert d._optimizer_timing["cache"]["resident_tiles"] == (4 if budget is None else 0)
    np.testing.assert_array_equal(vectors[0][0], vectors[1][0])
    np.testing.assert_array_equal(vectors[0][1], vectors[1][1])


def test_multi_tile_pipeline_with_staged_pageable_uploads(tmp_path):
    """Same, with bit volumes large enough (10 MB each, pageable NumPy memory) to go through the pinned
    staging ring: tile t+1 is staged by the prefetch thread's upload workers while tile t runs on the main
    stream, with the low-pass chained per bit behind the copies."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stacks = [cases.small_stack(cb["matrix"], shape=(20, 512, 512), seed=500 + i, density=2e-4) for i in range(3)]
    ds = _store(tmp_path, df_cb, stacks)
    bkg, nrm = cases.simple_vectors(16, nrm=300.0)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    kw = dict(lowpass_sigma=(3.0, 1.0, 1.0), minimum_pixels=6, normalization_method="global")
    single = []
    for t in range(3):
        one = PixelDecoder(ds, merfish_bits=16, verbose=0)
        one.decode_one_tile(t, **kw)
        single.append(one.decoded_barcodes)
        one._cleanup()
    assert sum(len(s) for s in single) > 100
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec.decode_all_tiles(assign_to_cells=False, **kw)
    for t in range(3):
        got = ds.load_local_decoded_spots(t)
        pd.testing.assert_frame_equal(got.reset_index(drop=True), single[t].reset_index(drop=True), check_dtype=False)


def test_decode_all_tiles_without_any_transcript(tmp_path):
    """Background-only tiles: every stage must pass empty tables through (per-tile parquet, filter, overlap
    de-duplication, filtered table) without raising, and the optimiser keeps the rounded global vectors."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, _cb = cases.codebook16()
    rng = np.random.default_rng(2)
    stacks = [(rng.poisson(100, (16, 6, 32, 40)) + 100).astype(np.uint16) for _ in range(2)]
    ds = _store(tmp_path, df_cb, stacks)
    bkg, nrm = cases.simple_vectors(16)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec.decode_all_tiles(assign_to_cells=True, lowpass_sigma=None, normalization_method="global")
    for t in range(2):
        got = ds.load_local_decoded_spots(t)
        assert got is not None and len(got) == 0 and "gene_id" in got.columns
    out = ds.load_global_filtered_decoded_spots()
    assert out is not None and len(out) == 0
    dec.optimize_normalization_by_decoding(n_iterations=2, lowpass_sigma=None, tile_indices=[0, 1])
    i_n, i_b = ds.load_decode_normalization_vectors(None, "iterative")
    g_n, g_b = ds.load_decode_normalization_vectors(None, "global")
    assert np.all(np.isfinite(i_n)) and np.all(np.isfinite(i_b)) and i_n.shape == (16,)


def test_local_multi_gpu_threads_equal_single_gpu(tmp_path):
    """``num_gpus=2`` without torch.distributed: one host thread and one decoder clone per local GPU,
    contiguous tile chunks like PD:4811-4838.  Same per-tile tables and filtered table as one GPU."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    stacks = [cases.small_stack(cb["matrix"], shape=(9, 40, 56), seed=600 + i, density=5e-3) for i in range(5)]
    bkg, nrm = cases.simple_vectors(16)
    outs = []
    for n_gpus in (1, 2):
        ds = ArrayDataStore(tmp_path / f"store{n_gpus}", codebook=df_cb)
        for st in stacks:
            ds.add_tile(st)
        ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
        dec = PixelDecoder(ds, merfish_bits=16, num_gpus=n_gpus, verbose=0)
        dec.decode_all_tiles(assign_to_cells=False, lowpass_sigma=(3.0, 1.0, 1.0), minimum_pixels=4,
                             normalization_method="global")
        outs.append(([ds.load_local_decoded_spots(t) for t in range(5)], ds.load_global_filtered_decoded_spots()))
    for a, b in zip(outs[0][0], outs[1][0]):
        pd.testing.assert_frame_equal(a, b)
    pd.testing.assert_frame_equal(outs[0][1], outs[1][1])
    assert sum(len(a) for a in outs[0][0]) > 30
