"""The oracle against golden vectors produced by the REFERENCE's own code.

``tests/golden/reference_*.npz`` were written by ``tests/golden/make_reference_golden.py``, which
executes ``/root/reference/src/merfish3danalysis/PixelDecoder.py`` unmodified with its GPU wheels
answered by NumPy/SciPy (``tests/golden/reference_shims.py``).  These tests pin the oracle --
the checker of every GPU parity test -- to what the reference computes: decoded / magnitude /
distance / scaled images bit-exact, transcript table identical.  When ``/root/reference`` is present
(build container) one more case runs the reference live on a fresh seed."""
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

import cases
from oracle import decode_oracle as orc
from scenarios import SCENARIOS, scenario_inputs, stack_digest, warp_tile_kwargs

GOLDEN = Path(__file__).resolve().parent / "golden"


def golden_table(g) -> pd.DataFrame:
    df = pd.DataFrame(g["table"], columns=[str(c) for c in g["table_columns"]])
    df["gene_id"] = [str(x) for x in g["gene_id"]]
    return df


def compare_with_reference_table(got: pd.DataFrame, ref: pd.DataFrame, rel=0.0):
    """``rel=0``: identical values (oracle); GPU tests pass the north-star tolerance."""
    assert len(got) == len(ref)
    assert [str(g) for g in got["gene_id"]] == ref["gene_id"].tolist()
    for c in ref.columns:
        if c == "gene_id":
            continue
        a = got[c].to_numpy(dtype=np.float64)
        b = ref[c].to_numpy(dtype=np.float64)
        if c.startswith("inertia_tensor_eigvals"):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-9, err_msg=c)  # eigensolver round-off
        elif c.endswith(("_center_z", "_center_y", "_center_x", "_intensity_sum")):
            # float64 sums of float32 weights: plane-wise vs full-volume vs atomic order differ by round-off
            assert np.array_equal(np.isnan(a), np.isnan(b)), c
            np.testing.assert_allclose(a, b, rtol=1e-9 if rel == 0.0 else rel, atol=1e-9, err_msg=c)
        elif rel == 0.0 or c in ("area", "barcode_id", "tile_idx", "on_bit_1", "on_bit_2", "on_bit_3", "on_bit_4",
                                 "tile_z", "tile_y", "tile_x"):
            np.testing.assert_array_equal(a, b, err_msg=c)
        else:
            np.testing.assert_allclose(a, b, rtol=rel, atol=1e-7, err_msg=c)


def oracle_on_scenario(sc):
    df_cb, cb, stack, pred, bkg, nrm, excluded = scenario_inputs(sc)
    is_3d = sc.get("microscope", "3D") != "2D"
    zr = sc.get("z_range")
    readout = stack if zr is None else stack[:, zr[0]:zr[1]]
    predictor = pred if (pred is None or zr is None) else pred[:, zr[0]:zr[1]]
    if sc.get("warp"):  # the warp samples the whole native volume; the z crop follows it (PD:1882-1890)
        _kw, bit_xf, bit_flows = warp_tile_kwargs(sc)
        full = orc.weight_readout(stack, pred)
        vols = [orc.warp_to_reference(full[b], bit_xf[b], (0.315, 0.098, 0.098)) if bit_flows[b] is None
                else orc.warp_to_reference_with_flow(full[b], bit_xf[b], (0.315, 0.098, 0.098), *bit_flows[b])
                for b in range(16)]
        readout = np.stack(vols)
        readout = readout if zr is None else readout[:, zr[0]:zr[1]]
        predictor = None
    excl_idx = [] if excluded is None else [cb["gene_ids"].index(g) for g in excluded]
    coords = dict(spacing=(0.315, 0.098, 0.098))
    if sc.get("origin") is not None:
        coords["origin"] = np.asarray(sc["origin"], dtype=np.float32)
    if zr is not None:
        coords["z_offset"] = float(zr[0])
    kw = {}
    if sc.get("mag") is not None:
        kw["magnitude_threshold"] = sc["mag"]
    if sc.get("chroma"):
        kw.update(optimize_mode=True, collect_centroids=(7, 1e-6))
    return orc.decode_tile(readout, predictor, cb, bkg, nrm, is_3d=is_3d, lowpass_sigma=sc["lowpass"],
                           minimum_pixels=sc["min_px"], excluded=excl_idx, **kw, **coords)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_equals_reference_golden(name):
    sc = SCENARIOS[name]
    g = np.load(GOLDEN / f"reference_{name}.npz")
    df, imgs = oracle_on_scenario(sc)
    if sc.get("slim"):
        assert stack_digest(scenario_inputs(sc)[2]) == str(g["stack_sha256"])  # the seeded input is the fixture's input
    else:
        np.testing.assert_array_equal(np.asarray(imgs["image"], dtype=np.float32), g["image"])
        np.testing.assert_array_equal(imgs["scaled"], g["scaled"])
    np.testing.assert_array_equal(imgs["decoded"], g["decoded"])
    np.testing.assert_array_equal(imgs["magnitude"], g["magnitude"])
    np.testing.assert_array_equal(imgs["distance"], g["distance"])
    ref = golden_table(g)
    assert len(ref) > 20
    compare_with_reference_table(df, ref)


def test_oracle_optimizer_equals_reference_golden():
    g = np.load(GOLDEN / "reference_optimizer.npz")
    _df, cb = cases.codebook16()
    res = orc.optimize_normalization([(s, None) for s in g["stacks"]], cb, 3, True, None, (0.9, 10.0), 4)
    np.testing.assert_array_equal(res["global_"][0], g["global_normalization"])
    np.testing.assert_array_equal(res["global_"][1], g["global_background"])
    np.testing.assert_array_equal(res["iterative"][0], g["iterative_normalization"])
    np.testing.assert_array_equal(res["iterative"][1], g["iterative_background"])


def simcfg0_tables(g):
    tile = pd.DataFrame(g["tile_table"], columns=[str(c) for c in g["tile_table_columns"]])
    tile["gene_id"] = [str(x) for x in g["tile_gene_id"]]
    filt = pd.DataFrame(g["filtered_table"], columns=[str(c) for c in g["filtered_table_columns"]])
    filt["gene_id"] = [str(x) for x in g["filtered_gene_id"]]
    return tile, filt


def test_oracle_equals_reference_on_the_simulation_cli_sequence():
    """configs[0]: optimiser (1 tile x 3 iterations, magnitude (0.9, 10), minimum 28 px, default low-pass) then
    decode_all_tiles with the blank-fraction filter, as cli/statphysbio_simulation/pixeldecode.py:259-292 calls them
    (fixture: the reference's own run, ``make_reference_golden.py simcfg0``)."""
    from oracle import table_oracle as tor
    from scenarios import SIM_CFG0, simcfg0_stack

    g = np.load(GOLDEN / "reference_simcfg0.npz")
    _df_cb, cb, stack = simcfg0_stack()
    np.testing.assert_array_equal(stack, g["stack"])
    res = orc.optimize_normalization([(stack, None)], cb, SIM_CFG0["iterations"], True, SIM_CFG0["lowpass"],
                                     SIM_CFG0["magnitude"], SIM_CFG0["min_px"])
    np.testing.assert_array_equal(res["global_"][0], g["global_normalization"])
    np.testing.assert_array_equal(res["global_"][1], g["global_background"])
    np.testing.assert_array_equal(res["iterative"][0], g["iterative_normalization"])
    np.testing.assert_array_equal(res["iterative"][1], g["iterative_background"])
    df, _imgs = orc.decode_tile(stack, None, cb, res["iterative"][1], res["iterative"][0], True, SIM_CFG0["lowpass"],
                                SIM_CFG0["magnitude"], SIM_CFG0["min_px"], spacing=(0.315, 0.098, 0.098))
    tile, filt = simcfg0_tables(g)
    assert len(tile) > 100 and tile["area"].min() >= SIM_CFG0["min_px"]
    compare_with_reference_table(df, tile)
    n_blank = sum(str(x).lower().startswith("blank") for x in cb["gene_ids"])
    keep, _diag = tor.blank_fraction_filter(df, n_blank, len(cb["gene_ids"]), 0.05)
    assert 0 < int(keep.sum()) < len(df)
    # the reference's filter adds its own annotation columns (voxel_intensity ... cell_id): compare what the decode made
    compare_with_reference_table(df[keep].reset_index(drop=True), filt[[c for c in filt.columns if c in tile.columns]])
    assert (filt["cell_id"] == -1).all() and filt["blank_fraction_keep"].all()


def test_reference_live_fresh_seed(tmp_path):
    """Build container only: run the reference itself on a seed no fixture holds."""
    import reference_shims as rs

    if not rs.reference_available():
        pytest.skip("/root/reference is not present on this machine")
    from merfish3d_analysis_b200.datastore import ArrayDataStore

    RefPD = rs.load_reference_pixeldecoder()
    sc = dict(shape=(8, 32, 40), seed=977, density=6e-3, lowpass=(1.0, 0.5, 0.5), norm="global", min_px=5,
              bkg=150.0, nrm=500.0, mag=(0.9, 10.0))
    df_cb, cb, stack, pred, bkg, nrm, _ex = scenario_inputs(sc)
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb)
    ds.add_tile(stack)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = RefPD(ds, merfish_bits=16, verbose=0)
    with rs.pandas2_semantics():
        res = dec.decode_one_tile(0, return_results=True, lowpass_sigma=sc["lowpass"], minimum_pixels=sc["min_px"],
                                  magnitude_threshold=sc["mag"], normalization_method="global")
    df, imgs = oracle_on_scenario(sc)
    np.testing.assert_array_equal(imgs["decoded"], res[4])
    np.testing.assert_array_equal(imgs["distance"], res[3])
    np.testing.assert_array_equal(imgs["magnitude"], res[2])
    np.testing.assert_array_equal(imgs["scaled"], res[1])
    ref = dec._df_barcodes
    assert len(ref) > 5
    compare_with_reference_table(df, ref.assign(gene_id=[str(x) for x in ref["gene_id"]]))
