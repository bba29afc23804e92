"""The CUDA path (``PixelDecoder`` over the C ABI) against golden vectors produced by the
REFERENCE's own code (``tests/golden/reference_*.npz``, see ``make_reference_golden.py``):
result images bit-exact, codeword assignments / areas / ids bit-exact, table floats within the
north-star tolerance (1e-5 relative)."""
from pathlib import Path

import numpy as np
import pytest

import cases
from scenarios import SCENARIOS, scenario_inputs, stack_digest, warp_tile_kwargs
from test_cpu_reference_golden import compare_with_reference_table, golden_table

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
REL = 1e-5


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_cuda_path_equals_reference_golden(tmp_path, name):
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    sc = SCENARIOS[name]
    g = np.load(GOLDEN / f"reference_{name}.npz")
    df_cb, _cb, stack, pred, bkg, nrm, excluded = scenario_inputs(sc)
    slim = bool(sc.get("slim"))
    if slim:  # the seeded inputs are the fixture's inputs
        assert stack_digest(stack) == str(g["stack_sha256"])
    else:
        np.testing.assert_array_equal(stack, g["stack"])
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb, microscope_type=sc.get("microscope", "3D"))
    extra = warp_tile_kwargs(sc)[0] if sc.get("warp") else {}
    ds.add_tile(stack, predictors=pred, stage_origin_zyx_um=sc.get("origin"), **extra)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    ref = golden_table(g)
    kw = dict(lowpass_sigma=sc["lowpass"], minimum_pixels=sc["min_px"], magnitude_threshold=sc.get("mag"),
              normalization_method=sc["norm"])
    dec = PixelDecoder(ds, merfish_bits=int(sc.get("bits", 16)), verbose=0, z_range=sc.get("z_range"), excluded_gene_ids=excluded)
    dec._optimize_normalization_weights = dec._collect_chromatic_centroids = bool(sc.get("chroma"))
    image, scaled, magnitude, distance, decoded = dec.decode_one_tile(0, return_results=True, **kw)
    if not slim:
        np.testing.assert_array_equal(image, g["image"])
        np.testing.assert_array_equal(scaled, g["scaled"])
    np.testing.assert_array_equal(decoded, g["decoded"])
    np.testing.assert_array_equal(magnitude, g["magnitude"])
    np.testing.assert_array_equal(distance, g["distance"])
    compare_with_reference_table(dec.decoded_barcodes, ref, rel=REL)
    # production path: gate + search + fused labelling, no result images
    dec2 = PixelDecoder(ds, merfish_bits=int(sc.get("bits", 16)), verbose=0, z_range=sc.get("z_range"), excluded_gene_ids=excluded)
    dec2._optimize_normalization_weights = dec2._collect_chromatic_centroids = bool(sc.get("chroma"))
    assert dec2.decode_one_tile(0, **kw) is None
    np.testing.assert_array_equal(dec2.decoded_image, g["decoded"])
    compare_with_reference_table(dec2.decoded_barcodes, ref, rel=REL)


def test_cuda_optimizer_equals_reference_golden(tmp_path):
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    g = np.load(GOLDEN / "reference_optimizer.npz")
    df_cb, _cb = cases.codebook16()
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb)
    for st in g["stacks"]:
        ds.add_tile(st)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=None,
                                           magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2])
    g_n, g_b = ds.load_decode_normalization_vectors(None, "global")
    i_n, i_b = ds.load_decode_normalization_vectors(None, "iterative")
    np.testing.assert_array_equal(g_n, g["global_normalization"])
    np.testing.assert_array_equal(g_b, g["global_background"])
    np.testing.assert_array_equal(i_n, g["iterative_normalization"])
    np.testing.assert_array_equal(i_b, g["iterative_background"])


@pytest.mark.parametrize("budget,lowpass", [(None, None), (0, None), ("one", None), (None, (3.0, 1.0, 1.0)), ("one", (3.0, 1.0, 1.0))])
def test_optimizer_tile_cache_does_not_change_the_result(tmp_path, budget, lowpass):
    """The optimiser keeps its tiles' decode inputs in HBM after the first pass (no re-upload, no re-filtering).
    Whatever the budget -- everything resident, nothing resident, or room for one tile so that the others keep
    streaming through the prefetch slots -- the vectors and the last iteration's tables are the same; with the
    low-pass off they are the reference's golden vectors."""
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    g = np.load(GOLDEN / "reference_optimizer.npz")
    df_cb, _cb = cases.codebook16()
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb)
    for st in g["stacks"]:
        ds.add_tile(st)
    tile_bytes = g["stacks"][0].nbytes * (2 if lowpass else 1)  # low-passed stacks are float32
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    if budget is not None:
        dec.tile_cache_budget_bytes = int(tile_bytes * 1.5) if budget == "one" else int(budget)
    dec._keep_temp_tables = True
    kept = {}
    orig = dec._save_barcodes

    def spy():
        kept[dec._tile_idx] = dec._df_barcodes.copy()
        orig()

    dec._save_barcodes = spy
    dec.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=lowpass,
                                           magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2])
    stats = dec._optimizer_timing["cache"]
    want_resident = {None: 3, 0: 0, "one": 1}[budget]
    assert stats["resident_tiles"] == want_resident, stats
    # with the low-pass on, the percentile seed's filtered volumes are the decode inputs (no value above the hot-pixel
    # threshold here): resident tiles are already there in iteration 0
    first = want_resident if lowpass is not None else 0
    assert stats.get("seeded_tiles", 0) == first, stats
    assert stats["hits"] == first + 2 * want_resident and stats["misses"] == 9 - first - 2 * want_resident, stats
    assert [it["cache_hits"] for it in dec._optimizer_timing["iterations"]] == [first, want_resident, want_resident]
    i_n, i_b = ds.load_decode_normalization_vectors(None, "iterative")
    if lowpass is None:
        np.testing.assert_array_equal(i_n, g["iterative_normalization"])
        np.testing.assert_array_equal(i_b, g["iterative_background"])
    # against an uncached run of the same thing
    ds2 = ArrayDataStore(tmp_path / "plain" / "qi2labdatastore", codebook=df_cb)
    for st in g["stacks"]:
        ds2.add_tile(st)
    dec2 = PixelDecoder(ds2, merfish_bits=16, verbose=0)
    dec2.tile_cache_budget_bytes = 0
    dec2._keep_temp_tables = True  # data frames per tile (the default 3-D path keeps the tables on the device)
    kept2 = {}
    orig2 = dec2._save_barcodes

    def spy2():
        kept2[dec2._tile_idx] = dec2._df_barcodes.copy()
        orig2()

    dec2._save_barcodes = spy2
    dec2.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=lowpass,
                                            magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2])
    p_n, p_b = ds2.load_decode_normalization_vectors(None, "iterative")
    np.testing.assert_array_equal(i_n, p_n)
    np.testing.assert_array_equal(i_b, p_b)
    import pandas as pd

    assert sorted(kept) == sorted(kept2) == [0, 1, 2]
    for t in kept:
        pd.testing.assert_frame_equal(kept[t], kept2[t])
    assert dec._tile_cache is None and not dec._buffers  # everything released at the end
    # the default path: feature tables stay on the device, the medians are taken from them there
    ds3 = ArrayDataStore(tmp_path / "device" / "qi2labdatastore", codebook=df_cb)
    for st in g["stacks"]:
        ds3.add_tile(st)
    dec3 = PixelDecoder(ds3, merfish_bits=16, verbose=0)
    if budget is not None:
        dec3.tile_cache_budget_bytes = dec.tile_cache_budget_bytes
    dec3.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=lowpass,
                                            magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2])
    d_n, d_b = ds3.load_decode_normalization_vectors(None, "iterative")
    np.testing.assert_array_equal(d_n, i_n)
    np.testing.assert_array_equal(d_b, i_b)
    assert [it["transcripts_pooled"] for it in dec3._optimizer_timing["iterations"]] == \
        [it["transcripts_pooled"] for it in dec._optimizer_timing["iterations"]]


def test_cuda_path_on_the_simulation_cli_sequence(tmp_path):
    """configs[0] through the public API exactly as the reference's simulation CLI drives it
    (cli/statphysbio_simulation/pixeldecode.py:259-292): optimize_normalization_by_decoding(n_random_tiles=1,
    n_iterations=3, magnitude (0.9, 10), minimum 28 px) then decode_all_tiles(assign_to_cells=False, blank-fraction
    filter).  Vectors, the per-tile table and the filtered table equal the reference's own run."""
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder
    from scenarios import SIM_CFG0, simcfg0_stack
    from test_cpu_reference_golden import simcfg0_tables

    g = np.load(GOLDEN / "reference_simcfg0.npz")
    df_cb, _cb, stack = simcfg0_stack()
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb)
    ds.add_tile(stack)
    dec = PixelDecoder(datastore=ds, use_mask=False, merfish_bits=16, verbose=0)
    dec.optimize_normalization_by_decoding(
        n_random_tiles=1, n_iterations=SIM_CFG0["iterations"], lowpass_sigma=SIM_CFG0["lowpass"],
        magnitude_threshold=SIM_CFG0["magnitude"], minimum_pixels=SIM_CFG0["min_px"], feature_predictor_threshold=0.5,
        estimate_chromatic_affines=False)
    g_n, g_b = ds.load_decode_normalization_vectors(None, "global")
    i_n, i_b = ds.load_decode_normalization_vectors(None, "iterative")
    np.testing.assert_array_equal(g_n, g["global_normalization"])
    np.testing.assert_array_equal(g_b, g["global_background"])
    np.testing.assert_array_equal(i_n, g["iterative_normalization"])
    np.testing.assert_array_equal(i_b, g["iterative_background"])
    dec.decode_all_tiles(
        assign_to_cells=False, lowpass_sigma=SIM_CFG0["lowpass"], magnitude_threshold=SIM_CFG0["magnitude"],
        minimum_pixels=SIM_CFG0["min_px"], feature_predictor_threshold=0.5, duplicate_radius_xy=None,
        duplicate_radius_z=None, filter_method="blank_fraction", target_gross_misid_rate=0.05, lr_fdr_target=0.05)
    tile, filt = simcfg0_tables(g)
    compare_with_reference_table(ds.load_local_decoded_spots(0), tile, rel=REL)
    got = ds.load_global_filtered_decoded_spots()
    compare_with_reference_table(got[[c for c in filt.columns]], filt, rel=REL)


def test_optimizer_seed_does_not_cache_tiles_with_hot_pixels(tmp_path):
    """The percentile seed replaces values above 50 000 (PD:1072-1074); the decode does not.  A tile that holds such a
    value must not take its decode input from the seed's volumes: it is staged by the first iteration instead, and the
    vectors equal those of a run without any cache."""
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    g = np.load(GOLDEN / "reference_optimizer.npz")
    df_cb, _cb = cases.codebook16()
    stacks = [st.copy() for st in g["stacks"]]
    stacks[1][3, 2, 5, 5] = 60000
    out = []
    for i, budget in enumerate((None, 0)):
        ds = ArrayDataStore(tmp_path / f"s{i}" / "qi2labdatastore", codebook=df_cb)
        for st in stacks:
            ds.add_tile(st)
        dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
        dec.tile_cache_budget_bytes = budget
        dec.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=(3.0, 1.0, 1.0),
                                               magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2])
        out.append((ds.load_decode_normalization_vectors(None, "global"), ds.load_decode_normalization_vectors(None, "iterative"),
                    dec._optimizer_timing["cache"]))
    assert out[0][2].get("seeded_tiles", 0) == 2 and out[0][2]["resident_tiles"] == 3
    assert [out[0][2]["hits"], out[0][2]["misses"]] == [2 + 3 + 3, 1]
    for a, b in zip(out[0][:2], out[1][:2]):
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
