"""world_size-2 gloo tests (CPU) of the multi-rank host logic: contiguous tile sharding, the
all-gather of variable-length transcript tables and the rank-0 -> JSON -> other-ranks hand-off
of the normalisation vectors.  The device work is replaced by a deterministic fake so that no
GPU is needed; the kernels themselves are covered by the `gpu` tests."""
import socket
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_table(decoder, tile_idx, vectors):
    """Deterministic transcript table that depends on the tile and on the vectors in use."""
    rng = np.random.default_rng(1000 + tile_idx)
    n = 40 + 7 * tile_idx
    words = rng.integers(0, len(decoder._gene_ids), n)
    on = np.argsort(~decoder._codebook_matrix.astype(bool), axis=1)[:, :4] + 1
    scale = float(np.mean(vectors)) if vectors is not None else 1.0
    data = {f"bit{i:02d}_mean_intensity": (rng.gamma(2.0, 100.0, n) + scale).astype(np.float32) for i in range(1, 17)}
    df = pd.DataFrame(data)
    df["gene_id"] = [decoder._gene_ids[w] for w in words]
    df["tile_idx"] = tile_idx
    for k in range(4):
        df[f"on_bit_{k + 1}"] = on[words, k]
    df["distance_min"] = 0.1
    return df


def _patch(decoder, log):
    import types

    def fake_global(self, **kw):
        nv = np.linspace(500, 800, 16, dtype=np.float32)
        bv = np.linspace(90, 120, 16, dtype=np.float32)
        self._datastore.save_decode_normalization_vectors(None, "global", nv, bv, decode_mode="3d")
        self._global_normalization_vector, self._global_background_vector = nv, bv
        self._global_normalization_loaded = True

    def fake_decode(self, tile_idx=0, gpu_id=0, normalization_method=None, **kw):
        self._prepare_normalization_state(normalization_method, True, gpu_id, kw.get("lowpass_sigma"))
        self._tile_idx = tile_idx
        _b, nrm = self._active_vectors()
        self._df_barcodes = _fake_table(self, tile_idx, nrm)
        log.append((tile_idx, normalization_method))

    decoder._global_normalization_vectors = types.MethodType(fake_global, decoder)
    decoder.decode_one_tile = types.MethodType(fake_decode, decoder)
    # the digit histograms of the pooled medians: host restatement of select_hist_kernel, reduced over gloo
    from merfish3d_analysis_b200 import normalization as _norm

    decoder._order_stats_backend = _norm._numpy_hist_backend


def _make_store(path, n_tiles):
    import cases
    from merfish3d_analysis_b200.datastore import ArrayDataStore

    df_cb, _cb = cases.codebook16()
    ds = ArrayDataStore(path, codebook=df_cb)
    for _ in range(n_tiles):
        ds.add_tile(np.zeros((16, 2, 4, 4), dtype=np.uint16))
    return ds


def _worker(rank, world, port, root, n_tiles, out_q):
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ds = ArrayDataStore(root)
        dec = PixelDecoder(ds, merfish_bits=16, num_gpus=world, verbose=0)
        log = []
        _patch(dec, log)
        dec.optimize_normalization_by_decoding(n_iterations=3, tile_indices=list(range(n_tiles)))
        nv, bv = ds.load_decode_normalization_vectors(None, "iterative")
        out_q.put((rank, log, nv.tolist(), bv.tolist(), dec._iterative_normalization_vector.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_optimiser_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    n_tiles = 5
    # single-process reference
    ds1 = _make_store(tmp_path / "single", n_tiles)
    dec1 = PixelDecoder(ds1, merfish_bits=16, num_gpus=1, verbose=0)
    log1 = []
    _patch(dec1, log1)
    dec1.optimize_normalization_by_decoding(n_iterations=3, tile_indices=list(range(n_tiles)))
    nv1, bv1 = ds1.load_decode_normalization_vectors(None, "iterative")
    assert [t for t, _ in log1[:n_tiles]] == list(range(n_tiles))
    assert [m for _, m in log1] == ["global"] * n_tiles + ["iterative"] * (2 * n_tiles)

    # two ranks over gloo
    _make_store(tmp_path / "dist", n_tiles)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path / "dist"), n_tiles, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, log_a, nv_a, bv_a, mem_a), (r1, log_b, nv_b, bv_b, mem_b) = results
    # contiguous chunks like PD:4811-4818: ceil(5/2) = 3 -> rank 0: tiles 0-2, rank 1: tiles 3-4
    assert sorted({t for t, _ in log_a}) == [0, 1, 2] and sorted({t for t, _ in log_b}) == [3, 4]
    # every rank ends with the same vectors, equal to the single-process result and to the host median of the pooled
    # table (the exchange is an all-reduce of digit histograms -- an exact order statistic, not a sum of means)
    from merfish3d_analysis_b200 import normalization as _norm

    np.testing.assert_array_equal(np.float32(nv_a), nv1)
    np.testing.assert_array_equal(np.float32(bv_a), bv1)
    assert nv_a == nv_b and bv_a == bv_b and mem_a == mem_b == nv_a


# ---------------------------------------------------------------------------------------------- z-slab exchange
def _fake_summary(rank):
    rng = np.random.default_rng(50 + rank)
    out = {}
    for k in range(1 + rank):  # a rank may own several slabs
        r = 10 * rank + k
        n = int(rng.integers(0, 9))
        out[r] = dict(z0=7 * r, z1=7 * r + 7, shape_yx=(40, 48), areas=rng.integers(1, 600, n).astype(np.float64),
                      pairs=rng.integers(0, 50, (int(rng.integers(0, 5)), 2)).astype(np.int64),
                      poisoned_here=rng.integers(0, 50, int(rng.integers(0, 3))).astype(np.int64),
                      poisoned_prev=rng.integers(0, 50, int(rng.integers(0, 3))).astype(np.int64))
    return out


def _fake_records(rank, n_bits=16):
    rng = np.random.default_rng(90 + rank)
    records, tabs = {}, {}
    for cid in range(rank + 1):
        n = int(rng.integers(1, 30))
        vals = rng.random((n, n_bits)).astype(np.float32 if cid % 2 else np.float16)
        records[(rank, cid)] = dict(lin=np.sort(rng.integers(0, 2**40, n)).astype(np.int64), vals=vals,
                                    mag=rng.random(n).astype(np.float16), dist=rng.random(n).astype(np.float16),
                                    dec=int(rng.integers(0, 140)))
    tabs[rank] = rng.random((int(rng.integers(0, 6)), 14 + n_bits))
    return records, tabs


def _assert_summary_equal(a, b):
    assert sorted(a) == sorted(b)
    for r in a:
        assert (a[r]["z0"], a[r]["z1"], tuple(a[r]["shape_yx"])) == (b[r]["z0"], b[r]["z1"], tuple(b[r]["shape_yx"]))
        for k in ("areas", "pairs", "poisoned_here", "poisoned_prev"):
            np.testing.assert_array_equal(np.asarray(a[r][k]).reshape(-1), np.asarray(b[r][k]).reshape(-1))


def _assert_records_equal(a, b):
    assert sorted(a) == sorted(b)
    for k in a:
        assert a[k]["dec"] == b[k]["dec"]
        for f in ("lin", "vals", "mag", "dist"):
            assert a[k][f].dtype == b[k][f].dtype, (k, f)
            np.testing.assert_array_equal(a[k][f], b[k][f])


def test_zslab_exchange_packing_round_trips():
    from merfish3d_analysis_b200 import sharded as sh

    for rank in range(3):
        s = _fake_summary(rank)
        _assert_summary_equal(sh.unpack_summary(sh.pack_summary(s)), s)
        rec, tabs = _fake_records(rank)
        rec2, tabs2 = sh.unpack_records(sh.pack_records(rec, tabs, 16))
        _assert_records_equal(rec2, rec)
        assert sorted(tabs2) == sorted(tabs)
        for r in tabs:
            np.testing.assert_array_equal(tabs2[r], tabs[r])
    assert sh.unpack_summary(sh.pack_summary({})) == {}
    assert sh.unpack_records(sh.pack_records({}, {}, 16)) == ({}, {})


def _exchange_worker(rank, world, port, out_q):
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    from merfish3d_analysis_b200 import sharded as sh

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        gathered = sh.all_gather_vectors(dist, sh.pack_summary(_fake_summary(rank)))
        merged = {k: v for g in gathered for k, v in sh.unpack_summary(g).items()}
        rec, tabs = _fake_records(rank)
        got = sh.gather_vectors(dist, sh.pack_records(rec, tabs, 16), dst=0)
        out_q.put((rank, merged, None if got is None else [sh.unpack_records(g) for g in got]))
    finally:
        dist.destroy_process_group()


def test_zslab_exchange_over_two_gloo_ranks():
    """The summaries (all ranks) and the records (rank 0) of the z-slab path travel as padded float64 tensors."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = {**_fake_summary(0), **_fake_summary(1)}
    for _rank, merged, _rec in results:
        _assert_summary_equal(merged, want)
    assert results[1][2] is None
    for r, (rec, tabs) in enumerate(results[0][2]):
        w_rec, w_tabs = _fake_records(r)
        _assert_records_equal(rec, w_rec)
        for k in w_tabs:
            np.testing.assert_array_equal(tabs[k], w_tabs[k])


def test_z_window_volume_serves_only_its_planes():
    from merfish3d_analysis_b200.datastore import ZWindowVolume

    block = np.arange(3 * 4 * 5, dtype=np.uint16).reshape(3, 4, 5)
    v = ZWindowVolume((10, 4, 5), 4, block)
    assert v.shape == (10, 4, 5) and v.dtype == np.uint16 and len(v) == 10
    np.testing.assert_array_equal(v[4:7], block)
    np.testing.assert_array_equal(v[5:6], block[1:2])
    assert v[6:6].shape[0] == 0
    with pytest.raises(IndexError):
        v[3:6]
    with pytest.raises(IndexError):
        v[6:8]
