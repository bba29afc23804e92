"""GPU parity tests: every C-ABI kernel against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): codeword image and canonically relabelled feature ids
bit-exact; distances / intensities within 1e-5 relative.  Because the kernels reproduce
the oracle's float32 operation order, most comparisons below are in fact exact.
"""
import numpy as np
import pytest
from scipy import ndimage as ndi

import cases

GOLDEN = __import__("pathlib").Path(__file__).resolve().parent / "golden"
from oracle import decode_oracle as orc

pytestmark = pytest.mark.gpu

REL = 1e-5


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def _same_f16(a, b, name=""):
    """bit-identical float16 images (NaN payloads aside: NaN must sit in the same voxels)."""
    assert a.dtype == np.float16 and b.dtype == np.float16
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    np.testing.assert_array_equal(nan_a, nan_b, err_msg=name + " NaN mask")
    np.testing.assert_array_equal(
        np.where(nan_a, 0, a.view(np.uint16)), np.where(nan_b, 0, b.view(np.uint16)), err_msg=name
    )


def _ctx(cb, excluded=()):
    from merfish3d_analysis_b200._capi import DecodeContext

    unit = orc.normalize_codebook(cb["matrix"]).astype(np.float32)
    return DecodeContext(unit, excluded, device=0), unit


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _decode_both(torch, cb, stack, bkg, nrm, mag=(1.5, 10.0), excluded=(), dense=True):
    ctx, unit = _ctx(cb, excluded)
    ctx.set_normalization(bkg, nrm)
    ctx.set_thresholds(cb["pixel_assignment_threshold"], mag[0], mag[1])
    d_stack = _dev(torch, stack)
    shape = stack.shape[1:]
    dec = torch.empty(shape, dtype=torch.int16, device="cuda")
    imgs = {}
    if dense:
        m = torch.empty(shape, dtype=torch.float16, device="cuda")
        d = torch.empty(shape, dtype=torch.float16, device="cuda")
        s = torch.empty(stack.shape, dtype=torch.float16, device="cuda")
        ctx.decode(d_stack, dec, m, d, s)
        imgs = dict(magnitude=m.cpu().numpy(), distance=d.cpu().numpy(), scaled=s.cpu().numpy())
    else:
        ctx.decode(d_stack, dec)
    torch.cuda.synchronize()
    imgs["decoded"] = dec.cpu().numpy()
    ref = orc.decode_pixels(
        np.asarray(stack, dtype=np.float32), unit, bkg, nrm, cb["pixel_assignment_threshold"], mag, excluded
    )
    return ctx, d_stack, dec, imgs, ref


@pytest.mark.parametrize("dense", [True, False])
@pytest.mark.parametrize("use_norm", [True, False])
def test_decode_u16_bit_exact(torch, dense, use_norm):
    _df, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(10, 40, 64), seed=11)
    bkg, nrm = cases.simple_vectors(16) if use_norm else (None, None)
    if not use_norm:  # unnormalised decode of a float stack that already lives in [0, 1]-ish units
        stack = ((stack.astype(np.float32) - 200.0) / 900.0).astype(np.float32)
    _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, dense=dense)
    np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    assert (ref["decoded"] >= 0).sum() > 100
    if dense:
        for k in ("magnitude", "distance", "scaled"):
            _same_f16(got[k], ref[k], k)


def test_decode_f32_input_with_specials(torch):
    """float32 stack (the low-passed case) including zeros, negatives, NaN and inf voxels."""
    _df, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(6, 32, 48), seed=5).astype(np.float32)
    rng = np.random.default_rng(3)
    stack += rng.normal(0, 0.37, stack.shape).astype(np.float32)
    stack[:, 0, 0, :8] = 0.0
    stack[3, 1, 1, 1] = np.nan
    stack[4, 1, 2, 2] = np.inf
    stack[5, 1, 3, 3] = -np.inf
    bkg, nrm = cases.simple_vectors(16, seed=2)
    with np.errstate(all="ignore"):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm)
    np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    for k in ("magnitude", "distance", "scaled"):
        _same_f16(got[k], ref[k], k)


def test_decode_degenerate_vectors_use_safe_division(torch):
    """normalisation entries of 0 / tiny / huge force the IEEE-division instantiation."""
    _df, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(4, 32, 32), seed=9)
    bkg, nrm = cases.simple_vectors(16, seed=4)
    nrm[2] = 0.0
    nrm[5] = 1e-30
    bkg[7] = 0.0
    with np.errstate(all="ignore"):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm)
        _c2, _s2, _d2, got2, _ = _decode_both(torch, cb, stack, bkg, nrm, dense=False)
    np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    np.testing.assert_array_equal(got2["decoded"], ref["decoded"])
    for k in ("magnitude", "distance", "scaled"):
        _same_f16(got[k], ref[k], k)


@pytest.mark.parametrize("dense", [True, False])
def test_decode_division_guards_f32(torch, dense):
    """The scale and unit-vector divisions run as a multiply by a correctly rounded reciprocal plus one FMA correction
    (voxel_math.cuh, div_by_rcp) when every operand is of ordinary size, and as the IEEE division otherwise.  A float32
    stack with values on both sides of every guard -- exact background hits (zero numerators), differences around
    2^-40, denormals, huge values, a norm with an all-ones significand, values that make the whole trace tiny -- must
    still reproduce NumPy's correctly rounded divisions bit for bit."""
    _df, cb = cases.codebook16()
    rng = np.random.default_rng(77)
    stack = cases.small_stack(cb["matrix"], shape=(6, 32, 48), seed=31).astype(np.float32)
    stack += rng.normal(0, 0.21, stack.shape).astype(np.float32)
    bkg, nrm = cases.simple_vectors(16, seed=8)
    flat = stack.reshape(16, -1)
    n = flat.shape[1]
    for b in range(16):
        idx = rng.choice(n, 400, replace=False)
        flat[b, idx[:50]] = bkg[b]                                        # zero numerator
        flat[b, idx[50:100]] = np.nextafter(bkg[b], np.float32(np.inf))   # one ulp above the background
        flat[b, idx[100:150]] = bkg[b] + np.float32(2.0 ** -41)           # rounds to the background or just above it
        flat[b, idx[150:200]] = np.float32(1e-42)                         # denormal sample
        flat[b, idx[200:250]] = np.float32(3e38)                          # huge
        flat[b, idx[250:300]] = bkg[b] + nrm[b] * np.float32(1e-13)       # quotient below 2^-40
        flat[b, idx[300:350]] = bkg[b] + nrm[b]                           # quotient exactly 1
        flat[b, idx[350:400]] = bkg[b] + nrm[b] * np.float32(0.99999994)  # quotient just below 1
    # whole traces that are tiny (norm below 2^-40) or have a norm with an all-ones significand
    tiny_vox = rng.choice(n, 64, replace=False)
    flat[:, tiny_vox] = (bkg + nrm * np.float32(3e-14))[:, None]
    ones_vox = rng.choice(n, 64, replace=False)
    flat[:, ones_vox] = bkg[:, None]
    flat[0, ones_vox] = bkg[0] + nrm[0] * np.float32(np.float32(0.99999994))  # single non-zero entry -> norm = that entry
    with np.errstate(all="ignore"):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, mag=(0.05, 10.0), dense=dense)
    np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    if dense:
        for k in ("magnitude", "distance", "scaled"):
            _same_f16(got[k], ref[k], k)


def test_decode_division_guards_vectors(torch):
    """Vectors on both sides of the host's reciprocal-division conditions (|nrm| in [2^-40, 2^40], significand not all
    ones, |bkg| in {0} or [2^-20, 2^30]); uint16 input, every result image against the oracle."""
    _df, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(4, 32, 32), seed=19)
    bkg, nrm = cases.simple_vectors(16, seed=5)
    nrm[0] = np.float32(2.0 ** 41)
    nrm[1] = np.float32(2.0 ** -41)
    nrm[2] = np.float32(np.nextafter(np.float32(1024.0), np.float32(0)))  # all-ones significand
    nrm[3] = np.float32(-700.0)
    bkg[4] = np.float32(1e-9)
    bkg[5] = np.float32(0.0)
    bkg[6] = np.float32(3e9)
    bkg[7] = np.float32(187.0)  # integer background: exact zero numerators
    nrm[8] = np.float32(2.0 ** 40)
    nrm[9] = np.float32(2.0 ** -40)
    with np.errstate(all="ignore"):
        for dense in (True, False):
            _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, mag=(0.05, 10.0), dense=dense)
            np.testing.assert_array_equal(got["decoded"], ref["decoded"])
            if dense:
                for k in ("magnitude", "distance", "scaled"):
                    _same_f16(got[k], ref[k], k)


@pytest.mark.parametrize("layout", ["packed", "isolated"])
def test_equal_distance_from_smaller_sum(torch, layout):
    """The dense-regime evaluation compares SUMS and takes a square root only when the sum improves.  Two different
    sums can round to the same float32 distance; NumPy's arg-min then keeps the FIRST index even if the later codeword
    has the smaller sum.  tests/golden/sqrt_tie_traces.npy holds 200 traces found by search
    (make_sqrt_tie_traces.py) on which exactly that happens; ``packed`` puts them side by side (whole warps unresolved:
    tensor-core marking path), ``isolated`` between background voxels (warp-cooperative path)."""
    _df, cb = cases.codebook16()
    traces = np.load(GOLDEN / "sqrt_tie_traces.npy")
    unit = orc.normalize_codebook(cb["matrix"][:, :16]).astype(np.float32)
    if layout == "packed":
        stack = np.ascontiguousarray(traces.T.reshape(16, 1, 8, 25))
    else:
        stack = np.zeros((16, 1, 40, 200), dtype=np.float32)
        stack[:, 0, np.arange(200) % 40, np.arange(200)] = traces.T
    for dense in (True, False):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, None, None, mag=(1e-3, 10.0), dense=dense)
        np.testing.assert_array_equal(got["decoded"], ref["decoded"])
        if dense:
            _same_f16(got["distance"], ref["distance"], "distance")
    # the fixture really holds the case: the winner (first arg-min of the distance) is not the codeword of least sum
    xs = orc.scale_traces(traces.T.copy(), np.zeros(16, np.float32), np.ones(16, np.float32))
    xh, _mag = orc.normalize_traces(np.clip(xs, 0, 1).astype(np.float32))
    xh = xh.T.astype(np.float32)
    acc = None
    for b in range(16):
        t = (xh[:, b:b + 1] - unit[None, :, b]).astype(np.float32)
        term = (t * t).astype(np.float32)
        acc = term if acc is None else (acc + term).astype(np.float32)
    d = np.sqrt(acc).astype(np.float32)
    first = d.argmin(1)
    least = acc.argmin(1)
    assert ((least > first) & (acc[np.arange(len(first)), first] > acc.min(1))).all()


def test_decode_all_foreground_worst_case(torch):
    """every voxel passes the magnitude gate: the search cannot hide behind sparsity."""
    _df, cb = cases.codebook16()
    rng = np.random.default_rng(21)
    stack = rng.integers(200, 1400, size=(16, 4, 32, 64)).astype(np.uint16)
    bkg, nrm = cases.simple_vectors(16, seed=6)
    for dense in (True, False):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, mag=(0.5, 10.0), dense=dense)
        np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    assert (ref["magnitude"].astype(np.float32) >= 0.5).mean() > 0.99


@pytest.mark.parametrize("code", ["mhd4_16", "hw4_22_k1000"])
def test_decode_saturated_traces_every_codeword_ties(torch, code):
    """Traces whose bits all clip to 1 are equidistant (in exact arithmetic) from EVERY codeword: the tensor-core
    marking pass hands the whole codebook to the exact evaluation for those voxels, so a warp queues 32 x K
    (voxel, codeword) pairs -- many times the queue's capacity (chunked evaluation, running minimum carried across
    chunks) -- and the float32 sequential sums decide among K near-equal distances (first arg-min).  Mixed with
    partly saturated and ordinary voxels so that queue segments of very different length share rounds."""
    if code == "mhd4_16":
        _df, cb = cases.codebook16()
        nb = 16
    else:
        _df, cb = cases.codebook22(n_words=1000, seed=4100)
        nb = 22
    rng = np.random.default_rng(77)
    shape = (2, 24, 64)
    stack = rng.integers(200, 1400, size=(nb,) + shape).astype(np.uint16)
    sat = rng.random(shape) < 0.4
    stack[:, sat] = 60000  # every bit clips to 1
    part = (~sat) & (rng.random(shape) < 0.3)  # a random subset of bits clips: ties among the codewords inside it
    hot = rng.random((nb,) + shape) < 0.5
    stack[hot & part[None]] = 60000
    bkg, nrm = cases.simple_vectors(nb, seed=9)
    for dense in (True, False):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, mag=(0.5, 10.0), dense=dense)
        np.testing.assert_array_equal(got["decoded"], ref["decoded"])
        if dense:
            _same_f16(got["distance"], ref["distance"], "distance")


@pytest.mark.parametrize("n_words", [3, 15, 16, 17, 33, 49])
def test_decode_dense_regime_small_and_ragged_codebooks(torch, n_words):
    """Dense-candidate regime with codebooks whose size is not a multiple of the 16-codeword MMA tile (a last tile of one
    row, exactly one tile, fewer rows than a tile): the padding rows of the fragment table have a zero on-bit sum and
    must never be handed to the exact evaluation, the per-tile prefix counts / rank select must address the right
    codeword.  Noise-level vectors: every voxel is a candidate, exact ties are common."""
    from merfish3d_analysis_b200 import synthetic

    m = synthetic.mhd4_codebook_matrix(16)[:n_words]
    df = synthetic.codebook_dataframe(m, n_blank=min(2, n_words - 1))
    cb = orc.load_codebook(df, 16)
    rng = np.random.default_rng(100 + n_words)
    stack = rng.integers(150, 260, size=(16, 3, 40, 64)).astype(np.uint16)
    stack[:, rng.random((3, 40, 64)) < 0.1] = 0  # all-zero traces: zero norm, every on-bit sum 0
    bkg = (187.0 + rng.uniform(-3, 3, 16)).astype(np.float32)
    nrm = (17.0 + rng.uniform(-2, 2, 16)).astype(np.float32)
    for dense in (True, False):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, mag=(0.05, 10.0), dense=dense)
        np.testing.assert_array_equal(got["decoded"], ref["decoded"])
        if dense:
            _same_f16(got["distance"], ref["distance"], "distance")
    assert (ref["decoded"] >= 0).mean() > 0.005


def test_decode_exclusions_do_not_fall_through(torch):
    _df, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(6, 40, 40), seed=13)
    bkg, nrm = cases.simple_vectors(16)
    _c, _s, _d, base, _ = _decode_both(torch, cb, stack, bkg, nrm, dense=False)
    present = np.unique(base["decoded"][base["decoded"] >= 0])
    excluded = tuple(int(v) for v in present[:3])
    _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, excluded=excluded, dense=False)
    np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    assert not np.isin(got["decoded"], excluded).any()
    changed = base["decoded"] != got["decoded"]
    assert np.all(got["decoded"][changed] == -1) and changed.any()


def test_decode_22bit_codebook(torch):
    _df, cb = cases.codebook22(n_words=300)
    stack = cases.small_stack(cb["matrix"], shape=(4, 40, 64), seed=17)
    bkg, nrm = cases.simple_vectors(22, seed=8)
    for dense in (True, False):
        _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm, dense=dense)
        np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    assert (ref["decoded"] >= 0).sum() > 50


def test_decode_non_binary_codebook_generic_path(torch):
    """mixed on-bit counts (3/4/5) -> rows have different non-zero values -> generic search."""
    import pandas as pd

    rng = np.random.default_rng(31)
    rows = []
    for k in range(40):
        r = np.zeros(16, dtype=np.int64)
        r[rng.choice(16, size=int(rng.integers(3, 6)), replace=False)] = 1
        rows.append(r)
    m = np.unique(np.stack(rows), axis=0)
    df = pd.DataFrame(m, columns=[f"bit{i + 1:02d}" for i in range(16)])
    df.insert(0, "gene_id", [f"g{i}" for i in range(len(m))])
    cb = orc.load_codebook(df, 16)
    stack = cases.small_stack(m[(m.sum(1) == 4)], shape=(4, 32, 48), seed=19)
    bkg, nrm = cases.simple_vectors(16, seed=10)
    _c, _s, _d, got, ref = _decode_both(torch, cb, stack, bkg, nrm)
    np.testing.assert_array_equal(got["decoded"], ref["decoded"])
    _same_f16(got["distance"], ref["distance"], "distance")


@pytest.mark.parametrize("shape", [(5, 33, 47), (30, 40, 72), (2, 8, 8),
                                   # several tile rows / columns: partial last tile column, a single tile reflected on
                                   # both sides, a last tile row of one line
                                   (3, 200, 130), (2, 97, 40), (4, 131, 57), (2, 330, 64), (1, 68, 113)])
@pytest.mark.parametrize("mode2d", [False, True])
def test_lowpass_matches_scipy_bit_exact(torch, shape, mode2d):
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(41)
    vols = rng.integers(0, 4000, size=(3, *shape)).astype(np.uint16)
    got = ctx.lowpass(_dev(torch, vols), (3.0, 1.0, 1.0), mode2d).cpu().numpy()
    ref = orc.lowpass_stack(vols.astype(np.float32), (3.0, 1.0, 1.0), not mode2d)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("shape", [(5, 33, 47), (30, 40, 72)])
@pytest.mark.parametrize("mode2d", [False, True])
def test_lowpass_float32_accumulation_mode(torch, shape, mode2d):
    """Opt-in arithmetic (m3d_set_lowpass_mode(ctx, 1)): float32 weights, one float32 FMA per tap in ascending order --
    bit-exact against its restatement, templated radii (3,1,1) and the generic-radius path, with a predictor; and
    switching back restores SciPy's float64 arithmetic."""
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(45)
    vols = rng.integers(0, 4000, size=(2, *shape)).astype(np.uint16)
    pred = rng.uniform(0, 1, size=vols.shape).astype(np.float32)
    ctx.set_lowpass_accumulate("float32")
    for sigma in [(3.0, 1.0, 1.0), (1.3, 2.2, 0.8)]:
        got = ctx.lowpass(_dev(torch, vols), sigma, mode2d, predictor=_dev(torch, pred)).cpu().numpy()
        w = orc.weight_readout(vols, pred)
        ref = np.stack([orc.lowpass_image_float32_accumulation(v, sigma, not mode2d) for v in w])
        np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32), err_msg=str(sigma))
        f64 = orc.lowpass_stack(w, sigma, not mode2d)
        assert not np.array_equal(got, f64)  # it IS a different arithmetic ...
        np.testing.assert_allclose(got, f64, rtol=2e-6, atol=1e-3)  # ... a few float32 ulps away
    ctx.set_lowpass_accumulate("float64")
    got = ctx.lowpass(_dev(torch, vols), (3.0, 1.0, 1.0), mode2d).cpu().numpy()
    ref = orc.lowpass_stack(vols.astype(np.float32), (3.0, 1.0, 1.0), not mode2d)
    np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_lowpass_f32_with_predictor_and_other_sigmas(torch):
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(43)
    vols = rng.integers(0, 4000, size=(2, 9, 31, 40)).astype(np.uint16)
    pred = rng.uniform(0, 1, size=vols.shape).astype(np.float32)
    for sigma in [(3.0, 1.0, 1.0), (2.0, 1.5, 1.5), (1.0, 2.0, 0.7)]:
        got = ctx.lowpass(_dev(torch, vols), sigma, False, predictor=_dev(torch, pred)).cpu().numpy()
        w = orc.weight_readout(vols, pred)
        ref = orc.lowpass_stack(w, sigma, True)
        np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32), err_msg=str(sigma))
    f = (vols.astype(np.float32) * pred).astype(np.float32)
    got = ctx.lowpass(_dev(torch, f), (3.0, 1.0, 1.0), False).cpu().numpy()
    ref = ndi.gaussian_filter(f[0], (3.0, 1.0, 1.0))
    np.testing.assert_array_equal(got[0].view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("in_dtype", ["u16", "f32"])
def test_warp_affine_matches_scipy_bit_exact(torch, in_dtype):
    """decode-time warp = scipy.ndimage.affine_transform(order=1, mode='constant', cval=0)."""
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(71)
    vol = rng.integers(0, 60000, size=(14, 45, 52)).astype(np.uint16)
    if in_dtype == "f32":
        vol = (vol.astype(np.float32) * np.float32(0.37)).astype(np.float32)
    pred = rng.uniform(0, 1, size=vol.shape).astype(np.float32)
    spacing = (0.315, 0.098, 0.098)
    for k in range(4):
        xf = np.eye(4, dtype=np.float32)
        xf[:3, :3] += rng.normal(0, 0.004, (3, 3)).astype(np.float32)
        xf[:3, 3] = (rng.uniform(-1.0, 1.0, 3) * np.asarray(spacing) * (1 + k)).astype(np.float32)  # <= 4 px
        m, o = orc.warp_px_arguments(xf, spacing)
        for p in (None, pred):
            w = orc.weight_readout(vol, p)
            ref = orc.warp_to_reference(w, xf, spacing)
            got = ctx.warp_affine(_dev(torch, vol), m, o, predictor=None if p is None else _dev(torch, p)).cpu().numpy()
            np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))
            assert 0.05 < (ref != 0).mean() <= 1.0
        # plane window (z_range cropping / z-slab sharding)
        part = ctx.warp_affine(_dev(torch, vol), m, o, out_z0=3, out_nz=6).cpu().numpy()
        np.testing.assert_array_equal(part, orc.warp_to_reference(orc.weight_readout(vol, None), xf, spacing)[3:9])


def test_weight_kernel(torch):
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(47)
    r = rng.integers(0, 65535, size=(3, 17, 19)).astype(np.uint16)
    p = rng.uniform(0, 1, size=r.shape).astype(np.float32)
    got = ctx.weight(_dev(torch, r), _dev(torch, p)).cpu().numpy()
    np.testing.assert_array_equal(got, orc.weight_readout(r, p))
    # vector body / scalar tail, no predictor, views that do not start on a 16-byte boundary
    flat_r, flat_p = _dev(torch, r.ravel()), _dev(torch, p.ravel())
    for off, n in ((0, 969), (0, 968), (0, 7), (0, 1), (1, 64), (8, 801), (3, 9)):
        rr, pp = flat_r[off:off + n], flat_p[off:off + n]
        np.testing.assert_array_equal(ctx.weight(rr, pp).cpu().numpy(), orc.weight_readout(rr.cpu().numpy(), pp.cpu().numpy()))
        np.testing.assert_array_equal(ctx.weight(rr, None).cpu().numpy(), rr.cpu().numpy().astype(np.float32))


def _random_decoded(rng, shape, n_codes=6, fill=0.3):
    dec = np.full(shape, -1, dtype=np.int16)
    m = rng.uniform(size=shape) < fill
    dec[m] = rng.integers(0, n_codes, size=int(m.sum())).astype(np.int16)
    return dec


@pytest.mark.parametrize("mode2d", [False, True])
@pytest.mark.parametrize("fill", [0.08, 0.35])
def test_label_matches_oracle(torch, mode2d, fill):
    """equal-value CCL + both size filters; canonical ids = raster order of first voxel."""
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(53)
    dec = _random_decoded(rng, (7, 37, 45), fill=fill)
    # one oversized uniform block (> maximum_pixels) that must be dropped
    dec[1:5, 2:14, 2:14] = 3
    labels = torch.zeros(dec.shape, dtype=torch.int32, device="cuda")
    n = ctx.label(_dev(torch, dec), mode2d, 3.0, 500, labels=labels)
    raw = orc.label_decoded(dec, not mode2d)
    ref = cases.canonical_labels(orc.filter_label_sizes(raw, 3.0, 500))
    got = labels.cpu().numpy()
    # voxels of components dropped for exceeding maximum_pixels are marked -1 (z-slab sharding needs
    # to tell them from background); everything else is the oracle's canonical labelling
    counts = np.bincount(raw.ravel())
    oversized = np.isin(raw, np.flatnonzero(counts > 500)[1:] if counts[0] > 500 else np.flatnonzero(counts > 500))
    oversized &= raw != 0
    np.testing.assert_array_equal(got == -1, oversized)
    np.testing.assert_array_equal(np.where(got == -1, 0, got), ref)
    assert n == ref.max() and n > 5


@pytest.mark.parametrize("mode2d", [False, True])
@pytest.mark.parametrize("optimize", [False, True])
@pytest.mark.parametrize("capacity", [0, 37])
def test_fused_decode_label_equals_separate_calls(torch, mode2d, optimize, capacity):
    """fused path (gate->search raw hand-off, search->regionprops records) == separate calls;
    capacity=37 forces both hand-off buffers to overflow into the recompute paths."""
    _df, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(9, 40, 56), seed=67)
    bkg, nrm = cases.simple_vectors(16)
    ctx, d_stack, dec, got, ref = _decode_both(torch, cb, stack, bkg, nrm, dense=False)
    lab_a = torch.zeros(dec.shape, dtype=torch.int32, device="cuda")
    n_a = ctx.label(dec, mode2d, 4.0, 500, labels=lab_a)
    tab_a = ctx.features(d_stack, dec, optimize).cpu().numpy()
    ctx.set_sparse_capacity(capacity)
    dec_b = torch.empty_like(dec)
    lab_b = torch.zeros(dec.shape, dtype=torch.int32, device="cuda")
    n_b = ctx.decode_label(d_stack, dec_b, mode2d, 4.0, 500, labels=lab_b)
    tab_b = ctx.features(d_stack, dec_b, optimize).cpu().numpy()
    assert n_a == n_b and n_a > 10
    np.testing.assert_array_equal(dec_b.cpu().numpy(), ref["decoded"])
    np.testing.assert_array_equal(lab_a.cpu().numpy(), lab_b.cpu().numpy())
    np.testing.assert_array_equal(tab_a, tab_b)
    ref_lab = cases.canonical_labels(orc.filter_label_sizes(orc.label_decoded(ref["decoded"], not mode2d), 4.0, 500))
    np.testing.assert_array_equal(lab_b.cpu().numpy(), ref_lab)


def test_label_edge_cases(torch):
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    empty = np.full((3, 8, 8), -1, dtype=np.int16)
    assert ctx.label(_dev(torch, empty), False, 1.0, 500) == 0
    full = np.zeros((3, 8, 8), dtype=np.int16)  # one 192-voxel component
    lab = torch.zeros(full.shape, dtype=torch.int32, device="cuda")
    assert ctx.label(_dev(torch, full), False, 16.0, 500, labels=lab) == 1
    assert int(lab.min()) == 1 and int(lab.max()) == 1
    assert ctx.label(_dev(torch, full), False, 16.0, 100) == 0  # larger than maximum_pixels
    # float minimum_pixels: int() truncation like the reference (16.9 -> keeps area >= 16)
    dec = np.full((2, 8, 8), -1, dtype=np.int16)
    dec[0, 0, :8] = 1
    dec[0, 1, :8] = 1
    assert ctx.label(_dev(torch, dec), False, 16.9, 500) == 1
    assert ctx.label(_dev(torch, dec), False, 17.0, 500) == 0
    # odd, unaligned sizes
    rng = np.random.default_rng(59)
    dec = _random_decoded(rng, (3, 5, 7), fill=0.5)
    lab = torch.zeros(dec.shape, dtype=torch.int32, device="cuda")
    ctx.label(_dev(torch, dec), False, 1.0, 500, labels=lab)
    ref = cases.canonical_labels(orc.filter_label_sizes(orc.label_decoded(dec, True), 1.0, 500))
    np.testing.assert_array_equal(lab.cpu().numpy(), ref)


def _check_table(tab, ref_df, n_bits, optimize):
    assert tab.shape[0] == len(ref_df)
    np.testing.assert_array_equal(tab[:, 0].astype(np.int64), ref_df["first_voxel"].to_numpy())
    np.testing.assert_array_equal(tab[:, 1], ref_df["area"].to_numpy())
    for j, c in enumerate(("z", "y", "x")):
        np.testing.assert_allclose(tab[:, 3 + j], ref_df[c].to_numpy(dtype=np.float64), rtol=1e-12)
    np.testing.assert_allclose(tab[:, 12], ref_df["distance_min"].to_numpy(dtype=np.float64), rtol=REL)
    np.testing.assert_allclose(tab[:, 13], ref_df["magnitude_mean"].to_numpy(dtype=np.float64), rtol=REL)
    ref_bits = np.stack([ref_df[f"intensity_mean-{b}"].to_numpy(dtype=np.float64) for b in range(n_bits)], axis=1)
    np.testing.assert_allclose(tab[:, 14 : 14 + n_bits], ref_bits, rtol=REL, atol=1e-7)


@pytest.mark.parametrize("optimize", [False, True])
@pytest.mark.parametrize("in_dtype", ["u16", "f32"])
def test_features_match_oracle(torch, optimize, in_dtype):
    _df, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=(10, 40, 64), seed=23)
    if in_dtype == "f32":
        stack = orc.lowpass_stack(stack.astype(np.float32), (1.0, 0.6, 0.6), True)
    bkg, nrm = cases.simple_vectors(16, nrm=600.0 if in_dtype == "f32" else 900.0)
    ctx, d_stack, dec, _got, ref = _decode_both(torch, cb, stack, bkg, nrm, dense=False)
    n = ctx.label(dec, False, 3.0, 500)
    tab = ctx.features(d_stack, dec, optimize).cpu().numpy()
    labels = orc.filter_label_sizes(orc.label_decoded(ref["decoded"], True), 3.0, 500)
    intensity = np.asarray(stack, dtype=np.float32) if optimize else ref["scaled"]
    ref_df = orc.region_table(labels, ref["distance"], ref["magnitude"], intensity)
    assert n == len(ref_df) and n > 20
    _check_table(tab, ref_df, 16, optimize)
    # decoded id of every row = value at its first voxel
    np.testing.assert_array_equal(
        tab[:, 2].astype(np.int64), ref["decoded"].ravel()[ref_df["first_voxel"].to_numpy()].astype(np.int64)
    )
    # second central moments -> inertia eigenvalues (scikit-image semantics)
    ev = ctx.inertia_eigvals(torch.from_numpy(tab).cuda()).cpu().numpy()
    ref_ev = np.stack([ref_df[f"inertia_tensor_eigvals-{k}"].to_numpy(dtype=np.float64) for k in range(3)], axis=1)
    np.testing.assert_allclose(ev, ref_ev, rtol=1e-9, atol=1e-9)


def test_select_hist_order_statistics(torch):
    from merfish3d_analysis_b200 import normalization as nz

    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(61)
    a = rng.gamma(2.0, 150.0, size=(9, 33, 41)).astype(np.float32)
    a[0, 0, :5] = 0.0
    b = rng.gamma(2.0, 170.0, size=(5, 20, 31)).astype(np.float32)
    da, db = _dev(torch, a), _dev(torch, b)
    stats = nz.DeviceOrderStats(ctx)
    for q in (10.0, 50.0, 90.0, 99.9):
        assert float(stats.percentile(da, q)) == float(np.percentile(a.ravel(), q))
    assert float(stats.median([da, db])) == float(np.median(np.concatenate([a.ravel(), b.ravel()])))
    got = nz.global_normalization_vectors(ctx, [[da, db]])
    nrm, bkg = orc.global_normalization_vectors(
        [a[None], b[None]], 1, True, lowpass_sigma=None, hot_pixel_threshold=1e30
    )
    assert got[0][0] == nrm[0] and got[1][0] == bkg[0]
    # the histogram kernel's paths: 128-bit body / scalar tail, volumes that do not start on a 16-byte boundary,
    # sizes below one warp, values that all fall in ONE first-digit bin (warp-aggregated adds), negatives with and
    # without the clip, a predicate that rejects almost everything
    flat = torch.from_numpy(np.concatenate([
        rng.normal(300.0, 0.01, 70001), -rng.gamma(2.0, 50.0, 5003), np.zeros(17)]).astype(np.float32)).cuda()
    for off, n in ((0, flat.numel()), (1, 4096), (3, 31), (2, 1), (5, 65536 + 7), (0, 8)):
        v = flat[off:off + n]
        h = v.cpu().numpy()
        for q in (0.0, 10.0, 50.0, 100.0):
            assert float(stats.percentile(v, q)) == float(np.percentile(h, q)), (off, n, q)
        assert float(stats.percentile(v, 90.0, sub=250.0, clip0=True)) == float(
            np.percentile(np.clip(h - np.float32(250.0), 0, None), 90.0)), (off, n)
        sel = h[h < np.float32(-10.0)]
        m = stats.median([v], pred=nz.PRED_LT, cutoffs=[-10.0])
        assert (m is None and sel.size == 0) or float(m) == float(np.median(sel)), (off, n)
        assert stats.count([v], pred=nz.PRED_GT, cutoffs=[299.99]) == int((h > np.float32(299.99)).sum())


def test_select_hist_batch_equals_single_launches(torch):
    """m3d_select_hist_batch (a whole level of the optimiser's 2 x bits medians in one launch) fills the same rows as one
    m3d_select_hist launch per multiset -- empty multisets, more rows than one parameter block holds, every level's prefix
    -- and the medians built on it are np.median's."""
    from merfish3d_analysis_b200 import normalization as nm

    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    rng = np.random.default_rng(88)
    sets = [rng.gamma(2.0, 200.0, int(n)).astype(np.float32) for n in rng.integers(0, 5000, 150)]
    sets[3] = np.float32([])
    sets[5] = np.float32([7.5])
    sets[9] = -sets[9]
    dev = [_dev(torch, a) for a in sets]
    for shift, mask_bits in ((21, 0), (10, 11), (0, 22)):
        rows = []
        for a, d in zip(sets, dev):
            pm = (0xFFFFFFFF << (32 - mask_bits)) & 0xFFFFFFFF if mask_bits else 0
            key = 0
            if a.size and mask_bits:
                u = np.array([a[a.size // 2]], np.float32).view(np.uint32)[0]
                key = int(~u & 0xFFFFFFFF) if (u & 0x80000000) else int(u | 0x80000000)
            rows.append((d, pm, key & pm))
        hb = torch.zeros((len(rows), 2048), dtype=torch.int64, device="cuda")
        ctx.select_hist_batch(rows, hb, shift)
        hs = torch.zeros((len(rows), 2048), dtype=torch.int64, device="cuda")
        for r, (d, pm, pv) in enumerate(rows):
            if d.numel():
                ctx.select_hist(d, hs[r], prefix_mask=pm, prefix_value=pv, shift=shift)
        torch.cuda.synchronize()
        assert torch.equal(hb, hs), shift
        if mask_bits == 0:
            assert hb.sum(1).cpu().tolist() == [a.size for a in sets]
    got = nm.pooled_medians(dev, None, lambda n: torch.zeros((n, 2048), dtype=torch.int64, device="cuda"),
                            hist_batch_fn=ctx.select_hist_batch)
    for a, g in zip(sets, got):
        assert (np.isnan(g) and a.size == 0) or g == float(np.median(a.astype(np.float64)))
    # NaN = "no entry": a multiset handed over as a dense column with NaN in the rows that do not belong to it
    padded = []
    for a in sets[:40]:
        col = np.full(a.size + 37, np.nan, dtype=np.float32)
        col[rng.permutation(col.size)[: a.size]] = a
        padded.append(_dev(torch, col))
    got = nm.pooled_medians(padded, None, lambda n: torch.zeros((n, 2048), dtype=torch.int64, device="cuda"),
                            hist_batch_fn=ctx.select_hist_batch)
    for a, g in zip(sets[:40], got):
        assert (np.isnan(g) and a.size == 0) or g == float(np.median(a.astype(np.float64)))


def test_centroid_statistics_upstream_known_answer_and_random():
    """m3d_centroid_statistics against the reference's own known-answer case
    (tests/test_optimization_codeword_exclusions.py:153-205) and the oracle on a random volume."""
    import torch

    from merfish3d_analysis_b200._capi import DecodeContext
    from test_cpu_oracle_and_host import _upstream_centroid_case

    labels, intensity, expected = _upstream_centroid_case()
    ctx = DecodeContext(np.array([[1.0]], dtype=np.float32), (), device=0)  # one bit, one codeword lighting it
    code = torch.zeros(int(labels.max()) + 1, dtype=torch.int16, device="cuda")
    sums, peak = ctx.centroid_statistics(torch.from_numpy(labels).cuda(), torch.from_numpy(intensity[None]).cuda(),
                                         3, code)
    sums, peak = sums.cpu().numpy(), peak.cpu().numpy()
    for k in range(4):
        np.testing.assert_allclose(sums[1:, 0, k], expected[k][1:])
    np.testing.assert_allclose(peak[1:, 0], expected[4][1:])
    ctx.close()

    rng = np.random.default_rng(12)
    Z, Y, X, nb = 9, 33, 47, 6
    lab = np.zeros((Z, Y, X), dtype=np.int32)
    n_lab = 60
    for l in range(1, n_lab + 1):  # small boxes; later labels overwrite earlier ones
        z, y, x = rng.integers(0, Z - 1), rng.integers(0, Y - 3), rng.integers(0, X - 3)
        lab[z:z + rng.integers(1, 3), y:y + rng.integers(1, 4), x:x + rng.integers(1, 4)] = l
    lab[lab == 7] = -1  # an oversized component marked as dropped
    stack = rng.normal(50, 80, (nb, Z, Y, X)).astype(np.float32)
    cb = np.zeros((5, nb), dtype=np.float32)
    for k in range(5):
        cb[k, rng.choice(nb, 3, replace=False)] = 1 / np.sqrt(3)
    codes = rng.integers(0, 5, n_lab + 1).astype(np.int16)
    codes[0] = -1
    ctx = DecodeContext(cb, (), device=0)
    for z_support in (1, 3, 5, 7, 9):
        sums, peak = ctx.centroid_statistics(torch.from_numpy(lab).cuda(), torch.from_numpy(stack).cuda(), z_support,
                                             torch.from_numpy(codes).cuda())
        sums, peak = sums.cpu().numpy(), peak.cpu().numpy()
        pos = np.maximum(lab, 0)
        for b in range(nb):
            w, zs, ys, xs, pk = orc.plane_wise_centroid_statistics(pos, stack[b], z_support, n_lab + 1)
            on = np.array([l > 0 and cb[codes[l], b] > 0 for l in range(n_lab + 1)])
            np.testing.assert_allclose(sums[on, b, 0], w[on], rtol=1e-12)
            np.testing.assert_allclose(sums[on, b, 1], zs[on], rtol=1e-12)
            np.testing.assert_allclose(sums[on, b, 2], ys[on], rtol=1e-12)
            np.testing.assert_allclose(sums[on, b, 3], xs[on], rtol=1e-12)
            np.testing.assert_array_equal(peak[on, b], pk[on])
            assert not sums[~on, b].any() and not peak[~on, b].any()
    ctx.close()


def test_inertia_eigvals_match_lapack():
    """m3d_inertia_eigvals (Jacobi, float64) against numpy.linalg.eigvalsh on the scikit-image tensor,
    including degenerate regions (single voxel, straight lines, planes)."""
    import torch

    from merfish3d_analysis_b200._capi import DecodeContext

    rng = np.random.default_rng(5)
    rows = []
    for i in range(3000):
        n = int(rng.integers(1, 60))
        kind = i % 4
        pts = rng.integers(0, 12, (n, 3)).astype(np.float64)
        if kind == 1:
            pts[:, 1:] = 3.0  # line along z
        elif kind == 2:
            pts[:, 0] = 5.0  # plane
        elif kind == 3:
            pts = pts[:1]  # single voxel
        d = pts - pts.mean(axis=0)
        mu = d.T @ d
        rows.append([0, len(pts), 0, 0, 0, 0, mu[0, 0], mu[1, 1], mu[2, 2], mu[0, 1], mu[0, 2], mu[1, 2], 0, 0])
    tab = np.asarray(rows, dtype=np.float64)
    n = tab[:, 1]
    T = np.empty((len(tab), 3, 3))
    T[:, 0, 0], T[:, 1, 1], T[:, 2, 2] = (tab[:, 7] + tab[:, 8]) / n, (tab[:, 6] + tab[:, 8]) / n, (tab[:, 6] + tab[:, 7]) / n
    T[:, 0, 1] = T[:, 1, 0] = -tab[:, 9] / n
    T[:, 0, 2] = T[:, 2, 0] = -tab[:, 10] / n
    T[:, 1, 2] = T[:, 2, 1] = -tab[:, 11] / n
    want = np.sort(np.clip(np.linalg.eigvalsh(T), 0, None), axis=1)[:, ::-1]
    ctx = DecodeContext(np.array([[1.0]], dtype=np.float32), (), device=0)
    got = ctx.inertia_eigvals(torch.from_numpy(tab).cuda()).cpu().numpy()
    ctx.close()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_staged_upload_is_byte_identical():
    """m3d_upload_batch: pageable sources through the pinned ring (ragged sizes, more chunks than ring
    slots, several pieces in one pipeline, repeated calls reusing the ring) and pinned sources via DMA."""
    import torch

    from merfish3d_analysis_b200._capi import DecodeContext, M3dError

    ctx = DecodeContext(np.array([[1.0]], dtype=np.float32), (), device=0)
    rng = np.random.default_rng(3)
    sizes = [1, 4097, (1 << 20) + 3, (32 << 20), (32 << 20) * 9 + 12345, 7]
    for rep in range(2):
        srcs = [rng.integers(0, 256, n, dtype=np.uint8) for n in sizes]
        dsts = [torch.zeros(n, dtype=torch.uint8, device="cuda") for n in sizes]
        ctx.upload(list(zip(srcs, dsts)))
        torch.cuda.synchronize()
        for s, d in zip(srcs, dsts):
            assert np.array_equal(d.cpu().numpy(), s), (rep, s.size)
    # uint16 volume slices (what the tile loader passes) + a pinned source in the same batch
    vol = rng.integers(0, 65536, (5, 300, 1000), dtype=np.uint16)
    pinned = torch.empty((2, 300, 1000), dtype=torch.uint16, pin_memory=True)
    pinned.numpy()[...] = vol[:2]
    stack = torch.zeros((3, 3, 300, 1000), dtype=torch.uint16, device="cuda")
    ctx.upload([(vol[1:4], stack[0]), (np.ascontiguousarray(vol[2:5]), stack[1]), (pinned.numpy()[:], stack[2, :2])])
    torch.cuda.synchronize()
    got = stack.cpu().numpy()
    assert np.array_equal(got[0], vol[1:4]) and np.array_equal(got[1], vol[2:5]) and np.array_equal(got[2, :2], vol[:2])
    with pytest.raises(M3dError):
        ctx.upload([(vol[:, ::2], stack[0])])  # not contiguous
    with pytest.raises(M3dError):
        ctx.upload([(vol[0], stack[0])])  # size mismatch
    ctx.close()


def test_persistent_decoded_image_equals_fresh_fill(torch):
    """m3d_decode_label_persistent: the decoded image kept across calls (previous foreground reset, dense -1 fill
    skipped) must equal a fresh m3d_decode_label for every tile of a sequence, also after calls that invalidate
    the remembered foreground list (m3d_decode, m3d_label) and after the caller scribbles over the buffer and
    says so by making one non-persistent call."""
    _df, cb = cases.codebook16()
    ctx, _ = _ctx(cb)
    bkg, nrm = cases.simple_vectors(16)
    ctx.set_normalization(bkg, nrm)
    ctx.set_thresholds(cb["pixel_assignment_threshold"], 1.5, 10.0)
    shape = (10, 40, 64)
    stacks = [torch.from_numpy(cases.small_stack(cb["matrix"], shape=shape, seed=700 + i, density=4e-3)).cuda()
              for i in range(5)]
    keep = torch.empty(shape, dtype=torch.int16, device="cuda")
    for i, st in enumerate(stacks + stacks[::-1]):
        fresh = torch.empty(shape, dtype=torch.int16, device="cuda")
        n_ref = ctx.decode_label(st, fresh, False, 4.0, 500)
        t_ref = ctx.features(st, fresh, False, n_ref)
        if i == 3:
            ctx.decode(st, keep)  # dense rewrite of the kept buffer: must invalidate the remembered list
        if i == 5:
            ctx.label(fresh, False, 4.0, 500)  # rebuilds the foreground list from another image
        if i == 7:
            keep.fill_(5)  # caller wrote to the buffer -> one non-persistent call re-establishes it
            ctx.decode_label(st, keep, False, 4.0, 500)
        n = ctx.decode_label(st, keep, False, 4.0, 500, persistent=True)
        tab = ctx.features(st, keep, False, n)
        assert n == n_ref and n > 10
        assert torch.equal(keep, fresh), i
        assert torch.equal(tab, t_ref), i
    ctx.close()
