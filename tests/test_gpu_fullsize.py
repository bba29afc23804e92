"""Full-size (BASELINE configs[1]: 16 x 100 x 2048 x 2048 uint16) property tests.

The oracle cannot run at this size in reasonable time, so parity is checked through
size-independent properties: the fused production path against the dense (reference-complete)
kernel on a random million-voxel sample, bookkeeping identities between the decoded image, the
label image and the feature table, and z-slab sharding == unsharded on the whole volume."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPE = (100, 2048, 2048)


@pytest.fixture(scope="module")
def big():
    import torch

    from merfish3d_analysis_b200 import synthetic
    from merfish3d_analysis_b200._capi import DecodeContext

    free, _total = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~60 GB of free HBM")
    m = synthetic.mhd4_codebook_matrix(16)
    unit = (m / np.linalg.norm(m, axis=1, keepdims=True)).astype(np.float32)
    ctx = DecodeContext(unit, (), device=0)
    ctx.set_normalization(np.full(16, 200.0, np.float32), np.full(16, 900.0, np.float32))
    ctx.set_thresholds(0.7653668647, 1.5, 10.0)
    stack = synthetic.make_stack_device(m, SHAPE, 2002, device="cuda")
    yield ctx, stack, m
    ctx.close()


def test_fullsize_fused_path_properties(big):
    import torch

    ctx, stack, _m = big
    decoded = torch.empty(SHAPE, dtype=torch.int16, device="cuda")
    labels = torch.zeros(SHAPE, dtype=torch.int32, device="cuda")
    n = ctx.decode_label(stack, decoded, False, 16.0, 500, labels=labels)
    table = ctx.features(stack, decoded, False, n)
    assert n > 10000
    # bookkeeping identities
    area = table[:, 1]
    assert float(area.min()) >= 16 and float(area.max()) <= 500
    assert int(area.sum().item()) == int((labels > 0).sum().item())
    assert int(labels.max().item()) == n
    first = table[:, 0].to(torch.int64)
    assert bool((first[1:] > first[:-1]).all())  # canonical order = ascending first voxel
    assert bool((labels.reshape(-1)[first] == torch.arange(1, n + 1, device="cuda", dtype=torch.int32)).all())
    assert bool((decoded.reshape(-1)[first].to(torch.float64) == table[:, 2]).all())
    assert bool(((labels != 0) <= (decoded >= 0)).all())  # labelled voxels are decoded voxels
    # idempotence: a second run gives identical outputs
    decoded2 = torch.empty_like(decoded)
    n2 = ctx.decode_label(stack, decoded2, False, 16.0, 500)
    table2 = ctx.features(stack, decoded2, False, n2)
    assert n2 == n and torch.equal(decoded, decoded2) and torch.equal(table, table2)
    # persistent decoded image (previous foreground reset instead of the dense -1 fill): alternate two
    # normalisation settings with different foreground sets on ONE kept buffer; every result must equal a
    # fresh decode under the same setting
    keep = torch.empty_like(decoded)
    for rep, nrm_v in enumerate((900.0, 600.0, 900.0, 600.0)):
        ctx.set_normalization(np.full(16, 200.0, np.float32), np.full(16, nrm_v, np.float32))
        nk = ctx.decode_label(stack, keep, False, 16.0, 500, persistent=True)
        if nrm_v == 900.0:
            assert nk == n and torch.equal(keep, decoded), rep
        else:
            nf = ctx.decode_label(stack, decoded2, False, 16.0, 500)
            assert nk == nf and nk != n and torch.equal(keep, decoded2), rep
    ctx.set_normalization(np.full(16, 200.0, np.float32), np.full(16, 900.0, np.float32))
    del keep
    # fused production path == dense reference-complete kernel on a random sample of 2^20 voxels
    # (all decoded voxels of 64 random planes' rows would bias; take uniform voxels + all foreground of one plane)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    n_vox = decoded.numel()
    lin = torch.randint(0, n_vox, (1 << 20,), device="cuda", generator=g)
    fg = torch.nonzero(decoded.reshape(-1) >= 0).reshape(-1)
    lin = torch.cat([lin, fg[:: max(1, fg.numel() // (1 << 18))]])
    mini = stack.view(torch.int16).reshape(16, -1)[:, lin].contiguous().view(torch.uint16).reshape(16, 1, 1, -1)
    d_dense = torch.empty((1, 1, lin.numel()), dtype=torch.int16, device="cuda")
    mag = torch.empty((1, 1, lin.numel()), dtype=torch.float16, device="cuda")
    dist = torch.empty_like(mag)
    scaled = torch.empty(tuple(mini.shape), dtype=torch.float16, device="cuda")
    ctx.decode(mini, d_dense, mag, dist, scaled)
    assert torch.equal(d_dense.reshape(-1), decoded.reshape(-1)[lin])
    # distance_min of every component equals the dense kernel's distance at its argmin voxel set:
    # check the weaker, size-independent form on the sample: decoded voxels satisfy both gates
    sel = d_dense.reshape(-1) >= 0
    assert bool((dist.reshape(-1)[sel].float() <= 0.7656).all())
    assert bool((mag.reshape(-1)[sel].float() >= 1.5 - 1e-3).all())


def test_fullsize_sharded_equals_unsharded(big, tmp_path):
    """whole-volume z-slab sharding (4 sequential slabs on this GPU) == unsharded decode."""
    import pandas as pd
    import torch

    from merfish3d_analysis_b200 import synthetic
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    _ctx, stack, m = big
    host = torch.empty(stack.shape, dtype=torch.uint16, pin_memory=True)
    host.copy_(stack)
    torch.cuda.synchronize()
    ds = ArrayDataStore(tmp_path / "qi2labdatastore", codebook=synthetic.codebook_dataframe(m, n_blank=10))
    ds.add_tile(host.numpy())
    ds.save_decode_normalization_vectors(None, "global", np.full(16, 900.0, np.float32), np.full(16, 200.0, np.float32))
    a = PixelDecoder(ds, merfish_bits=16, verbose=0)
    a.decode_one_tile(0, lowpass_sigma=None, normalization_method="global")
    ref = a.decoded_barcodes
    ref_img = a.decoded_image
    a._cleanup()
    b = PixelDecoder(ds, merfish_bits=16, verbose=0)
    b.decode_one_tile_sharded(0, n_slabs=4, lowpass_sigma=None, normalization_method="global")
    assert len(ref) > 10000
    pd.testing.assert_frame_equal(b.decoded_barcodes, ref)
    np.testing.assert_array_equal(b.decoded_image, ref_img)


def test_fullsize_dense_candidate_regime_matches_oracle_on_sample(big):
    """Noise-level normalisation vectors (what the optimiser's percentile seed gives): every voxel passes the
    magnitude gate, clipped traces tie exactly among many codewords, and the search runs through the tensor-core
    candidate marking.  The decoded image of the whole tile is compared with the CPU oracle on 20 000 random
    voxels (the oracle's first-arg-min semantics decide the ties)."""
    import torch

    from oracle import decode_oracle as orc

    ctx, stack, m = big
    bkg = np.full(16, 187.0, np.float32)
    nrm = np.full(16, 17.0, np.float32)
    ctx.set_normalization(bkg, nrm)
    try:
        decoded = torch.empty(SHAPE, dtype=torch.int16, device="cuda")
        ctx.decode(stack, decoded)  # production path: gate (all candidates) + search
        frac = float((decoded >= 0).float().mean().item())
        assert 0.05 < frac < 0.6  # noise decodes: this really is the dense regime
        g = torch.Generator(device="cuda")
        g.manual_seed(11)
        lin = torch.randint(0, decoded.numel(), (20000,), device="cuda", generator=g)
        mini = stack.view(torch.int16).reshape(16, -1)[:, lin].contiguous().view(torch.uint16).reshape(16, 1, 1, -1)
        unit = orc.normalize_codebook(m)
        thr = float(np.sqrt(2 - 2 * ((4 - 2) / np.sqrt(4 * (4 - 2)))))
        ref = orc.decode_pixels(mini.cpu().numpy().astype(np.float32), unit, bkg, nrm, thr, (1.5, 10.0))
        got = decoded.reshape(-1)[lin].cpu().numpy()
        np.testing.assert_array_equal(got, ref["decoded"].reshape(-1))
        assert (got >= 0).sum() > 1000
    finally:
        ctx.set_normalization(np.full(16, 200.0, np.float32), np.full(16, 900.0, np.float32))


@pytest.mark.parametrize("origin", [(0, 0, 0), (37, 700, 1201), (84, 1792, 1792)])
def test_fullsize_production_path_matches_oracle_on_crops(big, origin):
    """The PRODUCTION path (gate + search + fused labelling + regionprops kernel) on the whole configs[1] tile against
    the CPU oracle on a 16 x 256 x 256 crop of the same tile: the decoded image of the crop voxel by voxel, and the
    feature-table row of every component the crop holds completely (components touching a crop face are cut by the
    crop, not by the decoder) -- first voxel, area, codeword, centroid, per-bit means, distance_min, magnitude_mean
    and inertia eigenvalues."""
    import torch

    import cases
    from oracle import decode_oracle as orc

    ctx, stack, _m = big
    _df_cb, cb = cases.codebook16()
    bkg, nrm = np.full(16, 200.0, np.float32), np.full(16, 900.0, np.float32)
    ctx.set_normalization(bkg, nrm)
    decoded = torch.empty(SHAPE, dtype=torch.int16, device="cuda")
    n = ctx.decode_label(stack, decoded, False, 16.0, 500)
    table = ctx.features(stack, decoded, False, n)
    eig = ctx.inertia_eigvals(table).cpu().numpy()
    table = table.cpu().numpy()
    z0, y0, x0 = origin
    cz, cy, cx = 16, 256, 256
    crop = stack[:, z0:z0 + cz, y0:y0 + cy, x0:x0 + cx].cpu().numpy()
    w = orc.weight_readout(crop, None)
    unit = orc.normalize_codebook(cb["matrix"][:, :16])
    out = orc.decode_pixels(w, unit, bkg, nrm, cb["pixel_assignment_threshold"], (1.5, 10.0), ())
    np.testing.assert_array_equal(decoded[z0:z0 + cz, y0:y0 + cy, x0:x0 + cx].cpu().numpy(), out["decoded"])
    raw = orc.label_decoded(out["decoded"], True)
    # crop faces that are not the tile's own border cut components: those labels are left out of the comparison
    cutting = [f for f in (
        raw[0].ravel() if z0 > 0 else None, raw[-1].ravel() if z0 + cz < SHAPE[0] else None,
        raw[:, 0].ravel() if y0 > 0 else None, raw[:, -1].ravel() if y0 + cy < SHAPE[1] else None,
        raw[:, :, 0].ravel() if x0 > 0 else None, raw[:, :, -1].ravel() if x0 + cx < SHAPE[2] else None) if f is not None]
    touching_raw = np.unique(np.concatenate(cutting))
    touching_raw = touching_raw[touching_raw > 0]
    labels = orc.filter_label_sizes(raw.copy(), 16.0)  # labels are not renumbered (PD:2987)
    labels[np.isin(labels, touching_raw)] = 0
    tab = orc.region_table(labels, out["distance"], out["magnitude"], out["scaled"])
    assert len(tab) >= 20
    zz, rem = np.divmod(tab["first_voxel"].to_numpy(np.int64), cy * cx)
    yy, xx = np.divmod(rem, cx)
    first_global = ((zz + z0) * SHAPE[1] + (yy + y0)) * SHAPE[2] + (xx + x0)
    pos = np.searchsorted(table[:, 0], first_global.astype(np.float64))
    assert (pos < len(table)).all() and np.array_equal(table[pos, 0], first_global.astype(np.float64)), \
        "a complete component of the crop is missing from the production table"
    got = table[pos]
    np.testing.assert_array_equal(got[:, 1], tab["area"].to_numpy())
    flat_dec = out["decoded"].ravel()
    np.testing.assert_array_equal(got[:, 2], flat_dec[tab["first_voxel"].to_numpy(np.int64)].astype(np.float64))
    np.testing.assert_allclose(got[:, 3], tab["z"].to_numpy() + z0, rtol=1e-12)
    np.testing.assert_allclose(got[:, 4], tab["y"].to_numpy() + y0, rtol=1e-12)
    np.testing.assert_allclose(got[:, 5], tab["x"].to_numpy() + x0, rtol=1e-12)
    np.testing.assert_array_equal(got[:, 12], tab["distance_min"].to_numpy(np.float64))
    np.testing.assert_array_equal(got[:, 13], tab["magnitude_mean"].to_numpy(np.float64))
    means = np.stack([tab[f"intensity_mean-{b}"].to_numpy(np.float64) for b in range(16)], axis=1)
    np.testing.assert_array_equal(got[:, 14:30], means)
    ev = np.stack([tab[f"inertia_tensor_eigvals-{k}"].to_numpy(np.float64) for k in range(3)], axis=1)
    np.testing.assert_allclose(eig[pos], ev, rtol=1e-9, atol=1e-9)
    # and nothing else: every production component whose first voxel lies strictly inside the crop (one voxel away
    # from the cutting faces) and that the crop holds completely is one of the oracle's rows
    fz, frem = np.divmod(table[:, 0].astype(np.int64), SHAPE[1] * SHAPE[2])
    fy, fx = np.divmod(frem, SHAPE[2])
    inside = (fz >= z0) & (fz < z0 + cz) & (fy >= y0) & (fy < y0 + cy) & (fx >= x0) & (fx < x0 + cx)
    prod_first_local = ((fz[inside] - z0) * cy + (fy[inside] - y0)) * cx + (fx[inside] - x0)
    raw_at = raw.ravel()[prod_first_local]
    complete = ~np.isin(raw_at, touching_raw)
    assert set(prod_first_local[complete]) == set(tab["first_voxel"].to_numpy(np.int64))
