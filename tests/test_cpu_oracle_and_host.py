"""CPU tests: the oracle against the reference's own known-answer tests and SciPy, the host
logic (codebook, exclusions, coordinates, normalisation medians, datastore layout) and the
C-ABI library's exported symbols.  No GPU compute is called here."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
from scipy import ndimage as ndi

import cases
from oracle import decode_oracle as orc

ROOT = Path(__file__).resolve().parent.parent


# ------------------------------------------------------------------ reference known answers
def test_thresholds_for_hw4_codebook():
    """PD:778-791 closed forms for on=4."""
    _df, cb = cases.codebook16()
    assert cb["on_bit_count"] == 4
    assert cb["pixel_assignment_threshold"] == pytest.approx(0.7653668647, abs=1e-9)
    assert cb["transcript_distance_threshold"] == pytest.approx(0.6058108931, abs=1e-9)
    assert cb["matrix"].shape == (140, 16) and cb["blank_count"] == 10
    unit = orc.normalize_codebook(cb["matrix"])
    assert set(np.unique(unit)) == {0.0, 0.5}


def test_one_on_bit_rows_are_dropped():
    df, _ = cases.codebook16()
    extra = pd.DataFrame([["single"] + [1] + [0] * 15], columns=df.columns)
    cb = orc.load_codebook(pd.concat([df, extra], ignore_index=True), 16)
    assert cb["matrix"].shape[0] == 140 and "single" not in cb["gene_ids"]


def test_exclusion_golden_vector_oracle_and_host():
    """tests/test_optimization_codeword_exclusions.py:114-120 of the reference."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    for fn in (orc.suppress_excluded, PixelDecoder._suppress_excluded_codeword_assignments):
        decoded = np.asarray([0, 1, 2, 1, -1], dtype=np.int16)
        nearest = np.asarray([0, 1, 2, 1, 1], dtype=np.int16)
        fn(decoded, nearest, (1,))
        np.testing.assert_array_equal(decoded, np.asarray([0, -1, 2, -1, -1]))


def test_warp_pixel_golden_oracle_and_host():
    """tests/test_pixeldecoder_coordinates.py:6-41 of the reference."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    pixel = np.array([10.0, 20.0, 30.0], dtype=np.float32)
    spacing = np.array([0.32, 0.098, 0.098], dtype=np.float32)
    origin = np.array([2761.3, 107.81, 0.0], dtype=np.float32)
    cam = np.array([[1, 0, 0, 0], [0, -0.07, -1, 0], [0, -1, 0.07, 0], [0, 0, 0, 1]], dtype=np.float32)
    aff = np.array([[1, 0, 0, 0.5], [0, 1, 0, 1.25], [0, 0, 1, -2.0], [0, 0, 0, 1]], dtype=np.float32)
    physical = pixel * spacing + origin
    camera_space = (cam @ np.array([*physical, 1.0]))[:3]
    expected = (aff @ np.array([*camera_space, 1.0]))[:3]
    for fn in (orc.warp_pixel, PixelDecoder._warp_pixel):
        np.testing.assert_allclose(fn(pixel, spacing, origin, aff, cam), expected, rtol=0, atol=1e-5)


def _host_decoder():
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    d = PixelDecoder.__new__(PixelDecoder)
    d._gene_ids = ["GeneA", "GeneB", "GeneC"]
    d._codebook_matrix = np.asarray([[1, 1, 0, 0], [1, 0, 1, 0], [1, 0, 0, 1]], dtype=np.int8)
    return d


def test_resolve_exclusions_reference_cases():
    """tests/test_optimization_codeword_exclusions.py:96-111 of the reference."""
    d = _host_decoder()
    assert d._resolve_excluded_gene_ids([" GeneB ", "GeneB"]) == (("GeneB",), (1,))
    with pytest.raises(ValueError, match="case-sensitive"):
        d._resolve_excluded_gene_ids(["geneb"])
    with pytest.raises(ValueError, match="every codeword"):
        d._resolve_excluded_gene_ids(["GeneA", "GeneB", "GeneC"])
    assert d._resolve_excluded_gene_ids(None) == ((), ())


# ------------------------------------------------------------------ oracle self-consistency
@pytest.mark.parametrize("sigma,axis", [(3.0, 0), (1.0, 1), (1.0, 2), (2.0, 1)])
def test_correlate1d_restatement_is_scipy_bit_exact(sigma, axis):
    rng = np.random.default_rng(3)
    a = rng.gamma(2.0, 300.0, size=(9, 21, 17)).astype(np.float32)
    ref = ndi.gaussian_filter1d(a, sigma, axis=axis)
    got = orc.correlate1d_restated(a, sigma, axis)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_decode_pixels_golden_fixture():
    """committed fixture (tests/golden/make_golden.py): oracle output must not drift."""
    g = np.load(ROOT / "tests" / "golden" / "decode_small.npz", allow_pickle=False)
    _df, cb = cases.codebook16()
    unit = orc.normalize_codebook(cb["matrix"])
    out = orc.decode_pixels(g["stack"].astype(np.float32), unit, g["bkg"], g["nrm"],
                            cb["pixel_assignment_threshold"], (1.5, 10.0))
    np.testing.assert_array_equal(out["decoded"], g["decoded"])
    np.testing.assert_array_equal(out["magnitude"].view(np.uint16), g["magnitude"].view(np.uint16))
    np.testing.assert_array_equal(out["distance"].view(np.uint16), g["distance"].view(np.uint16))
    labels = orc.filter_label_sizes(orc.label_decoded(out["decoded"], True), 4)
    np.testing.assert_array_equal(cases.canonical_labels(labels), g["labels"])


def test_zero_norm_and_tie_semantics():
    """SURVEY 8a notes: zero vector -> magnitude -1, distance 1, index 0; ties -> lowest index."""
    _df, cb = cases.codebook16()
    unit = orc.normalize_codebook(cb["matrix"])
    stack = np.zeros((16, 1, 1, 2), dtype=np.float32)
    stack[:, 0, 0, 1] = 1.0  # all bits equal: equidistant to every codeword
    out = orc.decode_pixels(stack, unit, None, None, cb["pixel_assignment_threshold"], (-2.0, 10.0))
    assert float(out["magnitude"][0, 0, 0]) == -1.0 and float(out["distance"][0, 0, 0]) == 1.0
    assert out["decoded"][0, 0, 0] == -1 and out["decoded"][0, 0, 1] == -1
    d, idx = orc.nearest_codeword(np.full((16, 1), 0.25, dtype=np.float32), unit)
    assert idx[0] == 0


def test_label_equal_value_26_connectivity_and_filters():
    dec = np.full((2, 4, 6), -1, dtype=np.int16)
    dec[0, 0, 0] = 5
    dec[1, 1, 1] = 5  # diagonal neighbour, same value -> joined in 3-D, separate in 2-D
    dec[0, 0, 2] = 6  # adjacent but different value -> separate
    dec[0, 0, 1] = 6
    l3 = orc.label_decoded(dec, True)
    assert l3[0, 0, 0] == l3[1, 1, 1] != 0 and l3[0, 0, 1] == l3[0, 0, 2] != l3[0, 0, 0]
    l2 = orc.label_decoded(dec, False)
    assert l2[0, 0, 0] != l2[1, 1, 1]
    kept = orc.filter_label_sizes(l3, 2.9)  # int(2.9)-1 = 1 -> keep area >= 2
    assert set(np.unique(kept)) == {0, l3[0, 0, 0], l3[0, 0, 1]}
    assert orc.filter_label_sizes(l3, 3).max() == 0


def test_fp16_transcript_gate_value():
    """largest float16 distance passing the on=4 transcript gate is 1240/2048."""
    _df, cb = cases.codebook16()
    t = cb["transcript_distance_threshold"]
    assert np.float32(np.float16(1240 / 2048)) <= t < np.float32(np.float16(1241 / 2048))


def test_iterative_vectors_match_reference_loop():
    """vectorised medians (product) == the reference's iterrows loop restated in the oracle."""
    from merfish3d_analysis_b200 import normalization as nz

    rng = np.random.default_rng(11)
    _df, cb = cases.codebook16()
    n = 300
    words = rng.integers(0, 140, n)
    on = np.argsort(~cb["matrix"].astype(bool), axis=1)[:, :4] + 1
    data = {f"bit{i:02d}_mean_intensity": rng.gamma(2.0, 200.0, n).astype(np.float32) for i in range(1, 17)}
    df = pd.DataFrame(data)
    df["gene_id"] = [cb["gene_ids"][w] for w in words]
    for k in range(4):
        df[f"on_bit_{k + 1}"] = on[words, k]
    got = nz.iterative_normalization_vectors(df, 16)
    ref = orc.iterative_normalization_vectors(df, 16)
    np.testing.assert_array_equal(got[0], ref[0])
    np.testing.assert_array_equal(got[1], ref[1])
    blanks_only = df[df["gene_id"].str.lower().str.startswith("blank")]
    assert nz.iterative_normalization_vectors(blanks_only, 16) is None
    assert orc.iterative_normalization_vectors(blanks_only, 16) is None


# ------------------------------------------------------------------ datastore layout
def test_datastore_layout_and_vector_round_trip(tmp_path):
    from merfish3d_analysis_b200.datastore import ArrayDataStore

    df_cb, cb = cases.codebook16()
    root = tmp_path / "qi2labdatastore"
    ds = ArrayDataStore(root, codebook=df_cb)
    stack = cases.small_stack(cb["matrix"], shape=(3, 8, 8), seed=1)
    tid = ds.add_tile(stack, persist=True)
    assert tid == "tile0000" and (root / "readouts" / "tile0000" / "bit001" / "corrected_data.npy").exists()
    nrm = np.linspace(1, 2, 16, dtype=np.float32) / 3
    bkg = np.linspace(3, 4, 16, dtype=np.float32) / 7
    ds.save_decode_normalization_vectors(None, "iterative", nrm, bkg, metadata={"codebook_sha256": "x"})
    ds.save_decode_normalization_vectors("runA", "global", nrm, bkg, decode_mode="3d")
    reopened = ArrayDataStore(root)  # workers re-open by path
    n2, b2 = reopened.load_decode_normalization_vectors(None, "iterative")
    np.testing.assert_array_equal(n2, nrm)  # float32 -> JSON -> float32 is lossless (SURVEY 8a note)
    np.testing.assert_array_equal(b2, bkg)
    assert reopened.load_decode_normalization_metadata(None, "iterative") == {"codebook_sha256": "x"}
    assert reopened.load_decode_normalization_vectors("runA", "global")[0] is not None
    assert reopened.load_decode_normalization_vectors("runA", "iterative") == (None, None)
    assert reopened.decoded_temporary_dir(None, 2) == root / "decoded" / "temporary" / "iteration_002"
    np.testing.assert_array_equal(reopened.load_local_readout_image(0, "bit003").result(), stack[2])
    with pytest.raises(ValueError):
        reopened.load_decode_normalization_vectors("bad key!", "global")
    t = pd.DataFrame({"gene_id": ["a"], "distance_min": [0.1]})
    reopened.save_local_decoded_spots(t, 0)
    assert (root / "decoded" / "tile0000_decoded_features.parquet").exists()
    pd.testing.assert_frame_equal(reopened.load_local_decoded_spots("tile0000"), t)
    reopened.save_global_filtered_decoded_spots(t)
    assert (root / "all_tiles_filtered_decoded_features" / "decoded_features.csv.gz").exists()


def test_contiguous_tile_chunks_like_reference():
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    chunks = PixelDecoder._contiguous_chunks(list(range(10)), 4)  # PD:4811-4818: ceil(10/4)=3
    assert chunks == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9]]
    assert PixelDecoder._contiguous_chunks([0, 1], 4) == [[0], [1], [], []]


def test_sharded_resolve_merges_filters_and_poisons():
    """pure host logic of the z-slab merge (merfish3d-analysis_b200/sharded.py)."""
    from merfish3d_analysis_b200 import sharded as sh

    assert sh.split_z(10, 3) == [(0, 3), (3, 7), (7, 10)] and sh.split_z(2, 4) == [(0, 1), (1, 2)]
    assert sh.lowpass_z_radius((3.0, 1.0, 1.0)) == 12
    # slab0 ids: 0 (area 10, alone) 1 (area 9, crosses) 2 (area 300, crosses) 3 (area 3, alone, too small)
    # slab1 ids: 0 (area 8, joins slab0:1 and slab2:0) 1 (area 300, joins slab0:2 -> 600 > 500) 2 (area 20, poisoned)
    # slab2 ids: 0 (area 2, joins slab1:0)
    areas = [np.array([10.0, 9.0, 300.0, 3.0]), np.array([8.0, 300.0, 20.0]), np.array([2.0])]
    pairs = [np.zeros((0, 2), int), np.array([[0, 1], [1, 2]]), np.array([[0, 0]])]
    poisoned = [np.zeros(0, int), np.array([2]), np.zeros(0, int)]
    keep, groups = sh.resolve(areas, pairs, poisoned, minimum_pixels=16.9, maximum_pixels=500)
    assert [k.tolist() for k in keep] == [[False, False, False, False], [False, False, False], [False]]
    assert groups == [[(0, 1), (1, 0), (2, 0)]]  # 9 + 8 + 2 = 19 >= 16; the 600-voxel and poisoned ones are gone
    keep, groups = sh.resolve(areas, pairs, poisoned, minimum_pixels=10, maximum_pixels=500)
    assert keep[0].tolist() == [True, False, False, False]


# ------------------------------------------------------------------ C ABI
def test_shared_library_exports_every_declared_symbol():
    from merfish3d_analysis_b200 import _capi

    header = (ROOT / "include" / "m3d_b200.h").read_text()
    declared = set(re.findall(r"\b(m3d_[a-z0-9_]+)\s*\(", header))
    declared.discard("m3d_ctx")
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    lib = ctypes.CDLL(str(_capi.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    lib.m3d_abi_version.restype = ctypes.c_int
    assert lib.m3d_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch

    from merfish3d_analysis_b200 import _capi

    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    unit = orc.normalize_codebook(cases.codebook16()[1]["matrix"]).astype(np.float32)
    with pytest.raises(_capi.M3dError, match="no CPU fallback"):
        _capi.DecodeContext(unit)


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "merfish3d-analysis_b200"
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|import_module\(.oracle|oracle[/.]_ref|decode_oracle\s+import", re.M)
    for p in pkg.rglob("*.py"):
        assert not pat.search(p.read_text()), p
    for p in pkg.glob("csrc/*"):
        assert "#include" not in "".join(l for l in p.read_text().splitlines() if "oracle" in l), p


def _upstream_centroid_case():
    """tests/test_optimization_codeword_exclusions.py:153-205 of the reference (known-answer case)."""
    labels = np.asarray(
        [[[0, 1, 0], [0, 0, 2]], [[0, 0, 0], [3, 0, 0]], [[0, 1, 0], [0, 0, 2]]], dtype=np.int32)
    intensity = np.arange(1, labels.size + 1, dtype=np.float32).reshape(labels.shape)
    intensity[0, 0, 0] = -5.0
    from scipy.ndimage import grey_dilation

    dil = grey_dilation(labels, size=(3, 1, 1))
    w = np.maximum(intensity, np.float32(0))
    coords = [np.arange(n, dtype=np.float32).reshape(s) for n, s in
              zip(labels.shape, ((-1, 1, 1), (1, -1, 1), (1, 1, -1)))]
    ml = int(labels.max()) + 1
    expected = [np.bincount(dil.ravel(), weights=w.ravel(), minlength=ml)]
    expected += [np.bincount(dil.ravel(), weights=(w * c).ravel(), minlength=ml) for c in coords]
    peak = np.zeros(ml, dtype=np.float32)
    np.maximum.at(peak, labels.ravel(), w.ravel())
    return labels, intensity, expected + [peak]


def test_plane_wise_centroid_statistics_upstream_known_answer():
    labels, intensity, expected = _upstream_centroid_case()
    got = orc.plane_wise_centroid_statistics(labels, intensity, z_support=3, minlength=int(labels.max()) + 1)
    for g, e in zip(got, expected, strict=True):
        np.testing.assert_allclose(g, e)


# ---------------------------------------------------------------------------------------------- pooled medians (optimiser exchange)
def _host_hist_backend():
    import torch

    from merfish3d_analysis_b200 import normalization as nm

    return nm._numpy_hist_backend, (lambda rows: torch.zeros((rows, 2048), dtype=torch.int64))


def test_pooled_medians_equal_numpy_median():
    """The radix-select walk behind the optimiser's exchange (three rounds of digit histograms for all multisets together)
    returns exactly np.median of the float64 view of float32 values: odd / even counts, duplicates, negatives, zeros of
    both signs, a single element, an empty multiset, denormals and infinities."""
    import torch

    from merfish3d_analysis_b200 import normalization as nm

    hist_fn, new_hist = _host_hist_backend()
    rng = np.random.default_rng(5)
    sets = [
        rng.normal(0, 100, 1001).astype(np.float32),
        rng.normal(0, 100, 1000).astype(np.float32),
        np.repeat(np.float32([3.5, 3.5, 7.25]), 40),
        np.float32([0.0, -0.0, 0.0, -0.0]),
        np.float32([42.0]),
        np.float32([]),
        np.float32([1e-42, 2e-42, np.inf, -np.inf, 5.0, 6.0]),
        rng.integers(0, 5, 999).astype(np.float32),
        (rng.gamma(2.0, 300.0, 20001)).astype(np.float32),
    ]
    got = nm.pooled_medians([torch.from_numpy(a.copy()) for a in sets], hist_fn, new_hist)
    for a, g in zip(sets, got):
        if a.size == 0:
            assert np.isnan(g)
        else:
            want = float(np.median(a.astype(np.float64)))
            assert g == want or (np.isnan(g) and np.isnan(want)), (a[:5], g, want)
    # the batched backend (one call per round: DecodeContext.select_hist_batch on the device) walks the same digits
    calls = []

    def batch_fn(rows, hist, shift):
        calls.append(len(rows))
        for r, (data, pm, pv) in enumerate(rows):
            if data.numel():
                hist_fn(data, hist[r], pm, pv, shift)

    got_b = nm.pooled_medians([torch.from_numpy(a.copy()) for a in sets], None, new_hist, hist_batch_fn=batch_fn)
    assert len(calls) == 3 and calls[0] == len(sets)
    assert all((g == h) or (np.isnan(g) and np.isnan(h)) for g, h in zip(got, got_b))


def test_pooled_medians_over_ranks_equal_the_pooled_median():
    """Splitting every multiset over 'ranks' and summing the histograms (what all_reduce does) gives the median of the union."""
    import torch

    from merfish3d_analysis_b200 import normalization as nm

    hist_fn, new_hist = _host_hist_backend()
    rng = np.random.default_rng(6)
    full = [rng.gamma(2.0, 200.0, n).astype(np.float32) for n in (0, 1, 2, 777, 4000)]
    parts = [[a[r::3] for a in full] for r in range(3)]  # three ranks; some get nothing of the small sets

    # emulate the collective: run the walk once, with a histogram function that sums over all ranks' parts
    def hist_all(data, row, pm, pv, sh):
        q = int(data[0].item()) if data.numel() else -1  # the query index is smuggled in as the first element
        for r in range(3):
            t = torch.from_numpy(parts[r][q].copy())
            if t.numel():
                hist_fn(t, row, pm, pv, sh)

    tags = [torch.tensor([float(i)] + [0.0] * 0, dtype=torch.float32) for i in range(len(full))]
    got = nm.pooled_medians(tags, hist_all, new_hist)
    for a, g in zip(full, got):
        if a.size == 0:
            assert np.isnan(g)
        else:
            assert g == float(np.median(a.astype(np.float64)))


def test_iterative_vector_queries_match_the_dataframe_statistic():
    """queries -> pooled medians -> rounding == the pandas restatement of PD:1263-1368, incl. blank genes in any case,
    missing gene ids, NaN means, rows with on-bits outside the columns, and bits nobody has on."""
    import torch

    from merfish3d_analysis_b200 import normalization as nm

    hist_fn, new_hist = _host_hist_backend()
    rng = np.random.default_rng(7)
    n = 3000
    df = pd.DataFrame({f"bit{i:02d}_mean_intensity": rng.gamma(2, 100, n).astype(np.float32).astype(np.float64) for i in range(1, 17)})
    genes = np.array([f"g{i}" for i in range(30)] + ["Blank-1", "blank2", "BLANK_3"], dtype=object)
    df["gene_id"] = genes[rng.integers(0, len(genes), n)]
    df.loc[7, "gene_id"] = None
    on = np.stack([rng.permutation(15)[:4] + 1 for _ in range(n)])  # bit 16 is never on
    for k in range(4):
        df[f"on_bit_{k + 1}"] = on[:, k]
    df.loc[11, "bit03_mean_intensity"] = np.nan
    want = nm.iterative_normalization_vectors(df, 16)
    q, kept = nm.iterative_vector_queries(df, 16, torch.device("cpu"))
    assert kept == int((~df["gene_id"].astype("string").str.lower().str.startswith("blank", na=False)).sum())
    med = nm.pooled_medians(q, hist_fn, new_hist)
    got = nm.finish_iterative_vectors(med[:16], med[16:])
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
    assert got[0][15] == 1.0  # no transcript has bit 16 on: NaN -> 1
    # values that are not float32 numbers are refused (the caller then gathers the tables)
    df2 = df.copy()
    df2.loc[0, "bit01_mean_intensity"] = 0.1  # not representable in float32
    assert nm.iterative_vector_queries(df2, 16, torch.device("cpu"))[0] is None


def test_device_row_queries_nan_padded_equal_the_compacted_multisets(tmp_path):
    """The optimiser's per-iteration multisets built from the device tables (per-bit means + codeword index): the dense
    NaN-padded columns handed to m3d_select_hist_batch hold exactly the values of the compacted selections, and the
    medians taken from them (NaN = no entry) are the pandas statistic of the reference's loop (PD:1290-1368)."""
    import torch

    import cases
    from merfish3d_analysis_b200 import normalization as nm
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    df_cb, cb = cases.codebook16()
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=df_cb)
    ds.add_tile(np.zeros((16, 2, 8, 8), dtype=np.uint16))
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    rng = np.random.default_rng(17)
    rows = []
    for n in (700, 0, 1300):
        vals = rng.gamma(2.0, 120.0, (n, 16)).astype(np.float32)
        if n:
            vals[rng.integers(0, n, 9), rng.integers(0, 16, 9)] = np.nan  # a mean that is not a number: no entry
        rows.append((torch.from_numpy(vals), torch.from_numpy(rng.integers(0, len(dec._gene_ids), n).astype(np.int64))))
    rows.append(None)
    cpu = torch.device("cpu")
    q_c, kept_c, n_c = dec._device_row_queries(rows, cpu)
    q_p, kept_p, n_p = dec._device_row_queries(rows, cpu, nan_padded=True)
    assert (kept_c, n_c) == (kept_p, n_p) and len(q_c) == len(q_p) == 32 and n_c == 2000
    for a, b in zip(q_c, q_p):
        assert b.is_contiguous() and b.numel() == kept_p
        np.testing.assert_array_equal(np.sort(a.numpy()), np.sort(b.numpy()[~np.isnan(b.numpy())]))
    # medians: batch restatement that skips NaN like the kernel, against the DataFrame statistic
    hist_fn, new_hist = _host_hist_backend()

    def batch_fn(items, hist, shift):
        for r, (data, pm, pv) in enumerate(items):
            d = data[~torch.isnan(data)]
            if d.numel():
                hist_fn(d, hist[r], pm, pv, shift)

    med = nm.pooled_medians(q_p, None, new_hist, hist_batch_fn=batch_fn)
    got = nm.finish_iterative_vectors(med[:16], med[16:])
    vals = torch.cat([r[0] for r in rows if r is not None]).numpy()
    words = torch.cat([r[1] for r in rows if r is not None]).numpy()
    on = np.argsort(~dec._codebook_matrix.astype(bool), axis=1)[:, :4] + 1
    df = pd.DataFrame({f"bit{i + 1:02d}_mean_intensity": vals[:, i].astype(np.float64) for i in range(16)})
    df["gene_id"] = [dec._gene_ids[w] for w in words]
    for k in range(4):
        df[f"on_bit_{k + 1}"] = on[words, k]
    want = nm.iterative_normalization_vectors(df, 16)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
