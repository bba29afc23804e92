"""Image store -> HBM on the GPU: `m3d_zarr_read_chunks` (host entropy decode into pinned slots,
un-shuffle + placement kernels) against arrays written by the NumPy/pyarrow oracle, and the whole decode path fed
from a reference-layout datastore against the golden vectors produced by the reference's own code."""
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

import cases
from oracle import zarr_oracle as zo
from scenarios import SCENARIOS, scenario_inputs, warp_tile_kwargs
from test_cpu_reference_golden import compare_with_reference_table, golden_table
from test_cpu_zarr_store import CASES, _image

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def ctx():
    import torch
    from merfish3d_analysis_b200._capi import DecodeContext

    _df, cb = cases.codebook16()
    m = np.asarray(cb["matrix"], dtype=np.float32)
    unit = m / np.linalg.norm(m, axis=1, keepdims=True)
    c = DecodeContext(unit.astype(np.float32))
    yield c
    torch.cuda.synchronize()


@pytest.mark.parametrize("zstd_on", ["device", "host"])
@pytest.mark.parametrize("shape,dtype,chunks,compression,shards", CASES)
def test_device_reader_matches_oracle_written_images(tmp_path, ctx, shape, dtype, chunks, compression, shards, zstd_on,
                                                     monkeypatch):
    import torch
    from merfish3d_analysis_b200 import zarr_store as zs

    if zstd_on == "host":
        if compression != "blosc-zstd":
            pytest.skip("only Blosc-zstd frames have two decoders")
        monkeypatch.setenv("M3D_ZARR_GPU_ZSTD", "0")  # entropy stage on host threads (libzstd)

    rng = np.random.default_rng(shape[-1])
    a = _image(rng, shape, dtype)
    zo.write_ome_image(tmp_path / "img.ome.zarr", a, chunks=chunks, compression=compression, shards=shards)
    img = zs.ZarrImage(tmp_path / "img.ome.zarr")
    tdt = {np.dtype(np.uint16): torch.uint16, np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8,
           np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}[a.dtype]
    dst = torch.full(a.shape, 3, dtype=tdt, device=ctx.device)
    before = ctx.launches_by_kernel()
    zs.transfer(ctx, [(img, dst)])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dst.cpu().numpy(), a)
    after = ctx.launches_by_kernel()
    assert after.get("zarr_unshuffle_place_kernel", 0) > before.get("zarr_unshuffle_place_kernel", 0)
    chunk_bytes = int(np.prod(img.array.chunks)) * a.dtype.itemsize  # Blosc stores buffers under 128 bytes as they are
    if compression == "blosc-lz4" and chunk_bytes >= 128:  # LZ4 frames are decoded on the device, one warp per stream
        assert after.get("blosc_lz4_decode_kernel", 0) > before.get("blosc_lz4_decode_kernel", 0)
        mine = tmp_path / "mine"  # and the un-split frames this package's writer produces
        zs.write_ome_image(mine, a, chunks=chunks, compression=compression)
        dst.fill_(1)
        zs.transfer(ctx, [(zs.ZarrImage(tmp_path / "mine.ome.zarr"), dst)])
        torch.cuda.synchronize()
        np.testing.assert_array_equal(dst.cpu().numpy(), a)
    if a.ndim == 3:
        for z0, z1 in ((0, 1), (3, min(11, shape[0])), (shape[0] - 2, shape[0])):
            win = torch.full((z1 - z0,) + a.shape[1:], 9, dtype=tdt, device=ctx.device)
            zs.transfer(ctx, [(img.window(z0, z1), win)])
            torch.cuda.synchronize()
            np.testing.assert_array_equal(win.cpu().numpy(), a[z0:z1])


@pytest.mark.parametrize("env", [{"M3D_ZARR_HOST_LZ4": "1"}, {"M3D_ZARR_MMAP": "0"}, {"M3D_IO_THREADS": "1"},
                                 {"M3D_IO_THREADS": "3", "LOCAL_WORLD_SIZE": "8"}])
def test_alternative_host_paths_give_the_same_volume(tmp_path, ctx, monkeypatch, env):
    """LZ4 frames decoded by the host threads instead of the device, chunk files read instead of mapped, one worker
    (every chunk waits for the previous one's slot), few workers and many chunks (slot reuse)."""
    import torch
    from merfish3d_analysis_b200 import zarr_store as zs

    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(8)
    a = rng.poisson(120, (40, 100, 90)).astype(np.uint16)
    for comp in ("blosc-lz4", "blosc-zstd"):
        zo.write_ome_image(tmp_path / f"{comp}.ome.zarr", a, chunks=(4, 16, 16), compression=comp)  # 420 chunks
        dst = torch.zeros(a.shape, dtype=torch.uint16, device=ctx.device)
        zs.transfer(ctx, [(zs.ZarrImage(tmp_path / f"{comp}.ome.zarr"), dst)])
        torch.cuda.synchronize()
        np.testing.assert_array_equal(dst.cpu().numpy(), a)


def test_unwritten_chunks_and_corrupt_chunks_on_the_device(tmp_path, ctx):
    import json
    import torch
    from merfish3d_analysis_b200 import zarr_store as zs
    from merfish3d_analysis_b200._capi import M3dError

    a = np.zeros((20, 64, 64), dtype=np.uint16)
    a[:8, :32, :32] = 7
    a[12:, 40:, 5:] = 900
    for shards in (None, (16, 64, 64)):
        p = tmp_path / f"s{shards is not None}"
        zo.write_zarr3_array(p / "0", a, (8, 32, 32), shards=shards, skip_fill_chunks=True)
        (p / "zarr.json").write_text(json.dumps({"zarr_format": 3, "node_type": "group", "attributes": {}}))
        dst = torch.full(a.shape, 5, dtype=torch.uint16, device=ctx.device)
        zs.transfer(ctx, [(zs.ZarrImage(p), dst)])
        torch.cuda.synchronize()
        np.testing.assert_array_equal(dst.cpu().numpy(), a)
    assert ctx.launches_by_kernel().get("zarr_fill_chunk_kernel", 0) > 0
    # a damaged chunk is an error, and the context keeps working afterwards
    rng = np.random.default_rng(0)
    b = rng.poisson(100, (16, 64, 64)).astype(np.uint16)
    zo.write_ome_image(tmp_path / "img.ome.zarr", b, chunks=(8, 32, 32))
    f = tmp_path / "img.ome.zarr" / "0" / "c" / "1" / "1" / "0"
    good = f.read_bytes()
    f.write_bytes(good[: len(good) // 2])
    dst = torch.zeros(b.shape, dtype=torch.uint16, device=ctx.device)
    with pytest.raises(M3dError):
        zs.transfer(ctx, [(zs.ZarrImage(tmp_path / "img.ome.zarr"), dst)])
    torch.cuda.synchronize()
    f.write_bytes(good)
    zs.transfer(ctx, [(zs.ZarrImage(tmp_path / "img.ome.zarr"), dst)])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dst.cpu().numpy(), b)
    # the same for frames the device decodes: corruption is reported through the kernel's error flag
    zo.write_ome_image(tmp_path / "lz.ome.zarr", b, chunks=(8, 32, 32), compression="blosc-lz4")
    f = tmp_path / "lz.ome.zarr" / "0" / "c" / "1" / "1" / "0"
    good = f.read_bytes()
    raised = 0
    for trial in range(6):
        bad = bytearray(good)
        bad[len(bad) // 2 + trial] ^= 0xFF
        bad[-1 - trial] ^= 0x5A
        f.write_bytes(bytes(bad))
        try:
            zs.transfer(ctx, [(zs.ZarrImage(tmp_path / "lz.ome.zarr"), dst)])
        except M3dError:
            raised += 1
        torch.cuda.synchronize()
    assert raised > 0
    f.write_bytes(good[: len(good) // 2])
    with pytest.raises(M3dError):
        zs.transfer(ctx, [(zs.ZarrImage(tmp_path / "lz.ome.zarr"), dst)])
    torch.cuda.synchronize()
    f.write_bytes(good)
    zs.transfer(ctx, [(zs.ZarrImage(tmp_path / "lz.ome.zarr"), dst)])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dst.cpu().numpy(), b)


def test_transfer_mixed_sources_completes_pieces_in_order(tmp_path, ctx):
    import torch
    from merfish3d_analysis_b200 import zarr_store as zs

    rng = np.random.default_rng(3)
    vols = [rng.poisson(50 + 10 * i, (21, 70, 66)).astype(np.uint16) for i in range(5)]
    for i in (0, 1, 3):
        zo.write_ome_image(tmp_path / f"v{i}.ome.zarr", vols[i], chunks=(8, 32, 32))
    srcs = [zs.ZarrImage(tmp_path / "v0.ome.zarr").window(2, 19), zs.ZarrImage(tmp_path / "v1.ome.zarr").window(2, 19),
            np.ascontiguousarray(vols[2][2:19]), zs.ZarrImage(tmp_path / "v3.ome.zarr").window(2, 19),
            np.ascontiguousarray(vols[4][2:19])]
    dsts = [torch.zeros((17, 70, 66), dtype=torch.uint16, device=ctx.device) for _ in srcs]
    order = []
    sums = []

    def arrived(i):  # called with piece i completely enqueued on the current stream: reading it here is ordered
        order.append(i)
        sums.append(dsts[i].to(torch.int64).sum())

    zs.transfer(ctx, list(zip(srcs, dsts)), on_piece=arrived)
    torch.cuda.synchronize()
    assert order == [0, 1, 2, 3, 4]
    for i in range(5):
        np.testing.assert_array_equal(dsts[i].cpu().numpy(), vols[i][2:19])
        assert int(sums[i]) == int(vols[i][2:19].astype(np.int64).sum())


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_reference_goldens_from_a_reference_layout_store(tmp_path, name):
    """Every golden scenario of the reference, with the tile coming out of `<image>.ome.zarr` chunk files."""
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder
    from merfish3d_analysis_b200.zarr_store import Qi2labZarrDataStore, ZarrImage

    sc = SCENARIOS[name]
    g = np.load(GOLDEN / f"reference_{name}.npz")
    df_cb, _cb, stack, pred, bkg, nrm, excluded = scenario_inputs(sc)
    ds = Qi2labZarrDataStore.create(tmp_path / "qi2labdatastore", df_cb, microscope_type=sc.get("microscope", "3D"))
    extra = warp_tile_kwargs(sc)[0] if sc.get("warp") else {}
    ds.add_tile(stack, predictors=pred, stage_origin_zyx_um=sc.get("origin"), chunks=(4, 24, 40), **extra)
    ds = Qi2labZarrDataStore(tmp_path / "qi2labdatastore")  # re-opened from disk only
    assert isinstance(ds.load_local_readout_image(0, 0).result(), ZarrImage)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    ref = golden_table(g)
    kw = dict(lowpass_sigma=sc["lowpass"], minimum_pixels=sc["min_px"], magnitude_threshold=sc.get("mag"),
              normalization_method=sc["norm"])
    dec = PixelDecoder(ds, merfish_bits=int(sc.get("bits", 16)), verbose=0, z_range=sc.get("z_range"),
                       excluded_gene_ids=excluded)
    dec._optimize_normalization_weights = dec._collect_chromatic_centroids = bool(sc.get("chroma"))
    image, scaled, magnitude, distance, decoded = dec.decode_one_tile(0, return_results=True, **kw)
    if not sc.get("slim"):  # medium fixtures hold a checksum of the input instead of the float32 copy / scaled images
        np.testing.assert_array_equal(image, g["image"])
        np.testing.assert_array_equal(scaled, g["scaled"])
    np.testing.assert_array_equal(decoded, g["decoded"])
    np.testing.assert_array_equal(distance, g["distance"])
    compare_with_reference_table(dec.decoded_barcodes, ref, rel=1e-5)
    launches = dec._ctx(0).launches_by_kernel()
    assert launches.get("zarr_unshuffle_place_kernel", 0) >= stack.shape[0]


def test_optimizer_and_all_tiles_from_a_reference_layout_store(tmp_path):
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder
    from merfish3d_analysis_b200.zarr_store import Qi2labZarrDataStore

    g = np.load(GOLDEN / "reference_optimizer.npz")
    df_cb, _cb = cases.codebook16()
    zds = Qi2labZarrDataStore.create(tmp_path / "zarr" / "qi2labdatastore", df_cb)
    ads = ArrayDataStore(tmp_path / "npy" / "qi2labdatastore", codebook=df_cb)
    for i, st in enumerate(g["stacks"]):
        origin = (0.0, 30.0 * i, 0.0)
        zds.add_tile(st, stage_origin_zyx_um=origin, chunks=(8, 32, 32), compression=("blosc-zstd", "blosc-lz4", "zstd")[i % 3])
        ads.add_tile(st, stage_origin_zyx_um=origin)
    dec = PixelDecoder(zds, merfish_bits=16, verbose=0)
    dec.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=None,
                                           magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2])
    i_n, i_b = zds.load_decode_normalization_vectors(None, "iterative")
    np.testing.assert_array_equal(i_n, g["iterative_normalization"])
    np.testing.assert_array_equal(i_b, g["iterative_background"])
    g_n, g_b = zds.load_decode_normalization_vectors(None, "global")
    ads.save_decode_normalization_vectors(None, "global", g_n, g_b)
    ads.save_decode_normalization_vectors(None, "iterative", i_n, i_b)
    outs = []
    for ds in (zds, ads):
        d = PixelDecoder(ds, merfish_bits=16, verbose=0)
        d.decode_all_tiles(lowpass_sigma=(3.0, 1.0, 1.0), minimum_pixels=4, magnitude_threshold=(0.9, 10.0))
        outs.append([ds.load_local_decoded_spots(t) for t in ds.tile_ids] + [ds.load_global_filtered_decoded_spots()])
    assert sum(len(t) for t in outs[0][:-1]) > 20
    for a, b in zip(*outs):
        pd.testing.assert_frame_equal(a.reset_index(drop=True), b.reset_index(drop=True))
    # z-slab sharding reads only each slab's planes (+ halo) from the chunk files
    d1 = PixelDecoder(zds, merfish_bits=16, verbose=0)
    d1.decode_one_tile_sharded(1, n_slabs=3, lowpass_sigma=None, minimum_pixels=4, magnitude_threshold=(0.9, 10.0))
    d2 = PixelDecoder(ads, merfish_bits=16, verbose=0)
    d2.decode_one_tile(1, lowpass_sigma=None, minimum_pixels=4, magnitude_threshold=(0.9, 10.0))
    pd.testing.assert_frame_equal(d1.decoded_barcodes, d2.decoded_barcodes)


@pytest.mark.parametrize("mode", ["1", "2"])
def test_device_zstd_decoder_modes(tmp_path, ctx, monkeypatch, capsys, mode):
    """M3D_ZARR_GPU_ZSTD=1|2 (2 is the default): Blosc-zstd frames cross PCIe compressed and the library's own zstd decoder
    (csrc/zstd_decode.cuh, pinned to libzstd on the CPU) runs on the device -- mode 1 lane-serial
    (`blosc_zstd_decode_kernel`), mode 2 a warp as a team (`blosc_zstd_decode_kernel_v2`, csrc/zstd_lanes.cuh).
    Both must reproduce the host decode bit for bit and report a damaged frame."""
    import time

    import torch
    from merfish3d_analysis_b200 import zarr_store as zs
    from merfish3d_analysis_b200._capi import M3dError

    monkeypatch.setenv("M3D_ZARR_GPU_ZSTD", mode)
    rng = np.random.default_rng(21)
    for shape, dtype, chunks, tdt in (((24, 64, 64), np.uint16, (8, 32, 32), torch.uint16),
                                      ((20, 70, 90), np.float32, (8, 32, 48), torch.float32),
                                      ((16, 512, 512), np.uint16, (16, 512, 512), torch.uint16)):
        a = _image(rng, shape, dtype)
        if shape[1] == 512:
            a[:, 100:110, 200:230] += 4000
        p = tmp_path / f"z{shape[1]}"
        zs.write_ome_image(p, a, chunks=chunks, compression="blosc-zstd")
        img = zs.ZarrImage(tmp_path / f"z{shape[1]}.ome.zarr")
        dst = torch.full(a.shape, 7, dtype=tdt, device=ctx.device)
        before = ctx.launches_by_kernel().get("blosc_zstd_decode_kernel", 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        zs.transfer(ctx, [(img, dst)])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        np.testing.assert_array_equal(dst.cpu().numpy(), a)
        assert ctx.launches_by_kernel().get("blosc_zstd_decode_kernel", 0) > before
        with capsys.disabled():
            print(f"\n[device zstd mode {mode}] {shape} {np.dtype(dtype).name}: {a.nbytes / 1e6:.2f} MB in {dt * 1e3:.1f} ms")
    f = tmp_path / "z64.ome.zarr" / "0" / "c" / "1" / "1" / "0"
    good = f.read_bytes()
    f.write_bytes(good[:40] + bytes(len(good) - 40))
    dst = torch.zeros((24, 64, 64), dtype=torch.uint16, device=ctx.device)
    with pytest.raises(M3dError):
        zs.transfer(ctx, [(zs.ZarrImage(tmp_path / "z64.ome.zarr"), dst)])
    torch.cuda.synchronize()


def test_device_zstd_mode_values_are_validated(tmp_path, ctx, monkeypatch):
    """Only 0, 1 and 2 select a decoder; anything else is refused instead of silently picking one."""
    import torch
    from merfish3d_analysis_b200 import zarr_store as zs
    from merfish3d_analysis_b200._capi import M3dError

    a = _image(np.random.default_rng(3), (8, 32, 32), np.uint16)
    zs.write_ome_image(tmp_path / "v", a, chunks=(8, 32, 32), compression="blosc-zstd")
    img = zs.ZarrImage(tmp_path / "v.ome.zarr")
    dst = torch.zeros(a.shape, dtype=torch.uint16, device=ctx.device)
    for bad in ("3", "-1", "yes"):
        monkeypatch.setenv("M3D_ZARR_GPU_ZSTD", bad)
        with pytest.raises(M3dError):
            zs.transfer(ctx, [(img, dst)])
    monkeypatch.setenv("M3D_ZARR_GPU_ZSTD", "0")
    zs.transfer(ctx, [(img, dst)])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dst.cpu().numpy(), a)


def test_truncated_shard_is_an_error_on_the_device_path(tmp_path, ctx):
    """The device reader maps chunk byte ranges from the page cache: a shard index entry past the end of a truncated
    shard must come back as an error (the mapping would raise SIGBUS)."""
    import torch
    from merfish3d_analysis_b200 import zarr_store as zs
    from merfish3d_analysis_b200._capi import M3dError

    a = _image(np.random.default_rng(4), (16, 64, 64), np.uint16)
    zs.write_ome_image(tmp_path / "s", a, chunks=(8, 32, 32), shards=(16, 64, 64), compression="blosc-zstd")
    shard = next(f for f in (tmp_path / "s.ome.zarr" / "0" / "c").rglob("*") if f.is_file())
    good = shard.read_bytes()
    arr = zs.ZarrImage(tmp_path / "s.ome.zarr").array
    dst = torch.zeros(a.shape, dtype=torch.uint16, device=ctx.device)
    table = arr.chunk_table(dst.data_ptr())
    n_index = 8 * 16 + (4 if arr._index_crc else 0)
    shard.write_bytes(good[: (len(good) - n_index) // 2] + good[-n_index:])
    with pytest.raises(M3dError):
        ctx.zarr_read(table)
    torch.cuda.synchronize()
    shard.write_bytes(good)
    ctx.zarr_read(table)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dst.cpu().numpy(), a)
