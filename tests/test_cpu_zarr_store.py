"""Image store without a GPU: the C library's Blosc / Zarr v3 host decoder and encoder against the
NumPy + pyarrow oracle (an independent implementation: different shuffle code, different zstd / lz4 builds),
the datastore's metadata conventions, and malformed input."""

import json
import struct

import numpy as np
import pandas as pd
import pytest

from merfish3d_analysis_b200 import _capi, zarr_store as zs
from oracle import zarr_oracle as zo


def _image(rng, shape, dtype):
    if np.dtype(dtype).kind == "f":
        return (rng.gamma(2.0, 50.0, shape) * (rng.random(shape) > 0.3)).astype(dtype)
    return rng.poisson(180, shape).astype(dtype)


@pytest.mark.parametrize("typesize", [1, 2, 4, 8])
def test_blosc_frames_both_ways(typesize):
    rng = np.random.default_rng(typesize)
    for n in (0, 5, 127, 128, 1000, 33000):
        data = rng.integers(0, 30, n * typesize, dtype=np.uint8).tobytes()
        for cname in ("zstd", "lz4"):
            for shuffle in ("noshuffle", "shuffle", "bitshuffle"):
                for blocksize in (0, 2048 * typesize, 1000 * typesize):
                    for split in (None, True, False):  # lz4 frames are split into `typesize` streams by c-blosc
                        frame = zo.blosc_compress(data, typesize, cname=cname, shuffle=shuffle, blocksize=blocksize,
                                                  split=split)
                        assert _capi.blosc_decode_host(frame) == data, (n, cname, shuffle, blocksize, split)
                    mine = _capi.blosc_encode_host(data, typesize, cname, 5, shuffle, blocksize)
                    assert zo.blosc_decompress(mine) == data, (n, cname, shuffle, blocksize)
                    info = _capi.blosc_info(mine)
                    assert info["nbytes"] == len(data) and info["typesize"] == typesize and info["cbytes"] == len(mine)


def test_bitshuffle_layout_known_answer():
    # element j of 8 uint16 carries bit j in its low byte and bit (7 - j) in its high byte: row (byte b, bit i) of the
    # bit-shuffled block is one byte whose bit j says whether element j has bit i of byte b set
    elems = np.array([(1 << j) | (1 << (15 - j)) for j in range(8)], dtype="<u2")
    rows = np.frombuffer(zo.bit_shuffle(elems.tobytes(), 2), dtype=np.uint8)
    assert rows.tolist() == [1 << i for i in range(8)] + [1 << (7 - i) for i in range(8)]
    frame = zo.blosc_compress(np.tile(elems, 64).tobytes(), 2)
    assert _capi.blosc_decode_host(frame) == np.tile(elems, 64).tobytes()
    assert zo.crc32c(b"123456789") == 0xE3069283 == zs.crc32c(b"123456789")


def test_zstd_codec_both_ways():
    data = np.random.default_rng(0).poisson(3, 50000).astype(np.uint8).tobytes()
    assert zo._decompress(4, _capi.zstd_host(data, True), len(data)) == data
    assert _capi.zstd_host(zo._compress("zstd", data, 3), False, len(data)) == data


CASES = [
    ((37, 150, 140), np.uint16, (16, 64, 64), "blosc-zstd", None),
    ((37, 150, 140), np.uint16, (16, 64, 64), "blosc-lz4", None),
    ((20, 70, 90), np.float32, (8, 32, 48), "blosc-zstd", None),
    ((20, 70, 90), np.float32, (8, 32, 48), "zstd", None),
    ((9, 33, 31), np.uint16, (4, 16, 16), "none", None),
    ((37, 150, 140), np.uint16, (8, 32, 32), "blosc-zstd", (16, 64, 64)),
    ((5, 40, 40), np.uint16, None, "blosc-zstd", None),  # reference default chunks, clipped to the image
    ((3, 6, 40, 44), np.float32, (1, 4, 16, 16), "blosc-zstd", None),  # SOFIMA flow field layout
    ((50, 60), np.uint16, (32, 32), "blosc-zstd", None),
    ((17, 19, 23), np.uint8, (8, 8, 8), "blosc-zstd", None),
    ((10, 30, 34), np.int32, (4, 16, 16), "blosc-lz4", None),  # label images
    ((6, 20, 24), np.int64, (4, 8, 8), "blosc-zstd", None),
    ((5, 7, 3), np.uint16, (4, 4, 2), "blosc-lz4", None),  # chunks smaller than one 8-element shuffle group row
]


@pytest.mark.parametrize("shape,dtype,chunks,compression,shards", CASES)
def test_reader_matches_oracle_written_images(tmp_path, shape, dtype, chunks, compression, shards):
    rng = np.random.default_rng(len(shape) * 7 + shape[-1])
    a = _image(rng, shape, dtype)
    zo.write_ome_image(tmp_path / "img.ome.zarr", a, chunks=chunks, compression=compression, shards=shards,
                       extra_attributes={"round_linker": 2, "emission_um": 0.67})
    img = zs.ZarrImage(tmp_path / "img.ome.zarr")
    assert img.shape == a.shape and img.dtype == a.dtype
    np.testing.assert_array_equal(np.asarray(img), a)
    assert img.extra_attributes == {"round_linker": 2, "emission_um": 0.67}
    if a.ndim == 3:  # z windows touch only their chunks and crop them
        for z0, z1 in ((0, 1), (3, min(11, shape[0])), (shape[0] - 2, shape[0]), (4, 4)):
            np.testing.assert_array_equal(img.read(z0, z1), a[z0:z1])
            np.testing.assert_array_equal(img[z0:z1], a[z0:z1])
        np.testing.assert_array_equal(img[2], a[2])
    # the handle behaves like the array a reference loader returns
    np.testing.assert_array_equal(img * 2, a * 2)
    np.testing.assert_array_equal(1 + img, 1 + a)
    np.testing.assert_array_equal(img.astype(np.float64), a.astype(np.float64))
    assert img.max() == a.max() and float(np.mean(img)) == float(np.mean(a)) and len(img) == len(a)
    np.testing.assert_array_equal(img > a.mean(), a > a.mean())


@pytest.mark.parametrize("shape,dtype,chunks,compression,shards", CASES)
def test_writer_is_read_by_the_oracle(tmp_path, shape, dtype, chunks, compression, shards):
    rng = np.random.default_rng(11)
    a = _image(rng, shape, dtype)
    zs.write_ome_image(tmp_path / "img", a, chunks=chunks, compression=compression, shards=shards,
                       extra_attributes={"psf_idx": 1})
    b, attrs = zo.read_ome_image(tmp_path / "img.ome.zarr")
    np.testing.assert_array_equal(b, a)
    assert attrs == {"psf_idx": 1}
    np.testing.assert_array_equal(np.asarray(zs.ZarrImage(tmp_path / "img.ome.zarr")), a)
    meta = json.loads((tmp_path / "img.ome.zarr" / "zarr.json").read_text())
    assert meta["attributes"]["ome"]["version"] == "0.5"
    assert meta["attributes"]["ome"]["multiscales"][0]["datasets"][0]["path"] == "0"


def test_level_zero_is_taken_from_the_ome_metadata(tmp_path):
    a = np.arange(4 * 6 * 8, dtype=np.uint16).reshape(4, 6, 8)
    zo.write_ome_image(tmp_path / "img.ome.zarr", a, chunks=(2, 4, 4))
    root = tmp_path / "img.ome.zarr"
    (root / "0").rename(root / "s0")
    meta = json.loads((root / "zarr.json").read_text())
    meta["attributes"]["ome"]["multiscales"][0]["datasets"][0]["path"] = "s0"
    (root / "zarr.json").write_text(json.dumps(meta))
    np.testing.assert_array_equal(np.asarray(zs.ZarrImage(root)), a)


def test_unwritten_chunks_are_fill_value(tmp_path):
    a = np.zeros((20, 64, 64), dtype=np.uint16)
    a[:8, :32, :32] = 7
    for shards in (None, (16, 64, 64)):
        p = tmp_path / f"s{shards is not None}"
        zo.write_zarr3_array(p / "0", a, (8, 32, 32), shards=shards, skip_fill_chunks=True)
        (p / "zarr.json").write_text(json.dumps({"zarr_format": 3, "node_type": "group", "attributes": {}}))
        np.testing.assert_array_equal(np.asarray(zs.ZarrImage(p)), a)
    # a whole shard file missing
    (tmp_path / "sTrue" / "0" / "c" / "1" / "0" / "0").unlink()
    np.testing.assert_array_equal(np.asarray(zs.ZarrImage(tmp_path / "sTrue"))[16:], 0)


def test_malformed_input_raises(tmp_path):
    a = np.random.default_rng(0).poisson(100, (16, 64, 64)).astype(np.uint16)
    zo.write_ome_image(tmp_path / "img.ome.zarr", a, chunks=(16, 32, 32))
    f = tmp_path / "img.ome.zarr" / "0" / "c" / "0" / "1" / "0"
    good = f.read_bytes()
    for bad in (good[: len(good) // 2], good[:10], b"", good[:16] + bytes(len(good) - 16),
                struct.pack("<BBBBIII", 2, 1, 0x94, 2, 32768, 4096, len(good)) + good[16:]):
        f.write_bytes(bad)
        with pytest.raises(_capi.M3dError):
            np.asarray(zs.ZarrImage(tmp_path / "img.ome.zarr"))
    f.write_bytes(good)
    np.testing.assert_array_equal(np.asarray(zs.ZarrImage(tmp_path / "img.ome.zarr")), a)
    # what the reader does not handle is refused when the metadata is read, not guessed at
    mp = tmp_path / "img.ome.zarr" / "0" / "zarr.json"
    meta = json.loads(mp.read_text())
    for key, val in (("codecs", [{"name": "bytes", "configuration": {"endian": "big"}}]),
                     ("codecs", [{"name": "transpose", "configuration": {"order": [2, 1, 0]}}, {"name": "bytes"}]),
                     ("codecs", [{"name": "bytes"}, {"name": "gzip", "configuration": {"level": 5}}]),
                     ("chunk_grid", {"name": "rectilinear", "configuration": {}}),
                     ("data_type", "complex64")):
        mp.write_text(json.dumps(dict(meta, **{key: val})))
        with pytest.raises(ValueError):
            zs.ZarrImage(tmp_path / "img.ome.zarr")
    with pytest.raises(FileNotFoundError):
        zs.ZarrImage(tmp_path / "nothing.ome.zarr")


def test_truncated_shard_with_a_valid_index_is_an_error_not_a_fault(tmp_path):
    """A shard whose index survives (index at the start, or re-appended) but whose chunk bytes are cut off: the
    byte range of an entry lies past the end of the file.  Refused when the chunk table is built and again by the C
    reader (which maps the range: touching it would be SIGBUS)."""
    a = np.random.default_rng(5).poisson(100, (16, 64, 64)).astype(np.uint16)
    zo.write_ome_image(tmp_path / "img.ome.zarr", a, chunks=(8, 32, 32), shards=(16, 64, 64))
    shard = next(f for f in (tmp_path / "img.ome.zarr" / "0" / "c").rglob("*") if f.is_file())
    good = shard.read_bytes()
    img = zs.ZarrImage(tmp_path / "img.ome.zarr")
    arr = img.array
    n_index = 8 * 16 + (4 if arr._index_crc else 0)
    assert arr._index_at_end
    cut = good[: (len(good) - n_index) // 2] + good[-n_index:]  # half of the chunk bytes gone, index intact
    shard.write_bytes(cut)
    with pytest.raises(ValueError):
        np.asarray(zs.ZarrImage(tmp_path / "img.ome.zarr"))
    # the C reader on its own (a table built from the intact file, read after the truncation)
    shard.write_bytes(good)
    dst = np.zeros(a.shape, dtype=a.dtype)
    table = arr.chunk_table(dst.ctypes.data)
    shard.write_bytes(cut)
    with pytest.raises(_capi.M3dError):
        _capi.zarr_read_chunks_host(table)
    shard.write_bytes(good)
    _capi.zarr_read_chunks_host(table)
    np.testing.assert_array_equal(dst, a)


def test_unshuffled_typesize_one_frame_with_odd_block_size(tmp_path):
    """A Blosc frame written with typesize 1 and no shuffle for a uint16 array, block size not a multiple of the
    element size: the decoded bytes are linear, so elements straddling a block edge must survive placement."""
    a = np.random.default_rng(6).poisson(300, (8, 32, 32)).astype(np.uint16)
    zo.write_ome_image(tmp_path / "img.ome.zarr", a, chunks=(8, 32, 32), compression="blosc-zstd")
    f = tmp_path / "img.ome.zarr" / "0" / "c" / "0" / "0" / "0"
    f.write_bytes(zo.blosc_compress(a.tobytes(), typesize=1, cname="zstd", shuffle="noshuffle", blocksize=4099))
    np.testing.assert_array_equal(np.asarray(zs.ZarrImage(tmp_path / "img.ome.zarr")), a)
    f.write_bytes(zo.blosc_compress(a.tobytes(), typesize=1, cname="lz4", shuffle="noshuffle", blocksize=4099))
    np.testing.assert_array_equal(np.asarray(zs.ZarrImage(tmp_path / "img.ome.zarr")), a)


def _codebook(n_bits=8):
    rows = [["gene%d" % i] + [int((i >> b) & 1) for b in range(n_bits)] for i in range(1, 6)]
    return pd.DataFrame(rows, columns=["gene_id"] + [f"bit{i:02d}" for i in range(1, n_bits + 1)])


def test_datastore_reference_layout_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    cb = _codebook()
    ds = zs.Qi2labZarrDataStore.create(tmp_path / "qi2labdatastore", cb, voxel_size_zyx_um=(0.3, 0.1, 0.1))
    ro = rng.poisson(150, (8, 6, 40, 48)).astype(np.uint16)
    pr = rng.random((8, 6, 40, 48)).astype(np.float32)
    xf = np.eye(4)
    xf[:3, 3] = (0.1, -0.2, 0.3)
    flow = rng.normal(0, 0.2, (3, 2, 5, 6)).astype(np.float32)
    ds.add_tile(ro, pr, stage_origin_zyx_um=(1.0, 2.0, 3.0), global_xform=(np.eye(4), (4, 5, 6), (0.3, 0.1, 0.1)),
                wavelengths_um=[(0.561, 0.58)] * 4 + [(0.638, 0.67)] * 4, bit_round=[1, 1, 2, 2, 1, 1, 2, 2],
                round_transforms_zyx_um={2: xf}, chunks=(4, 16, 16),
                sofima_flow_fields={2: (flow, {"block_size": [8, 8, 8], "block_stride": [4, 4, 4]})})
    ds.add_tile(ro[:, ::-1].copy(), None, chunks=(4, 16, 16), bit_round=[1, 1, 2, 2, 1, 1, 2, 2])

    again = zs.Qi2labZarrDataStore(tmp_path / "qi2labdatastore", validate=True)
    assert again.tile_ids == ["tile0000", "tile0001"] and len(again.bit_ids) == 8
    assert again.round_ids == ["round001", "round002"]
    pd.testing.assert_frame_equal(again.codebook, cb, check_dtype=False)
    np.testing.assert_allclose(again.voxel_size_zyx_um, (0.3, 0.1, 0.1))
    # layout of docs/datastore.md:211-300
    root = tmp_path / "qi2labdatastore"
    assert (root / "readouts/tile0000/bit003/corrected_data.ome.zarr/0/zarr.json").exists()
    assert (root / "readouts/tile0000/bit003/feature_predictor_data.ome.zarr/zarr.json").exists()
    assert (root / "fiducial/tile0000/round002/local_sofima_flow_field.ome.zarr/0/zarr.json").exists()
    assert (root / "fiducial/tile0000/round001/attributes.json").exists()
    for b in range(8):
        img = again.load_local_readout_image("tile0000", b)
        np.testing.assert_array_equal(np.asarray(img.result()), ro[b])
        np.testing.assert_array_equal(np.asarray(again.load_local_feature_predictor_image(0, b, return_future=False)),
                                      pr[b])
    assert type(again.load_local_feature_predictor_image("tile0001", 0)).__name__ == "UnitPredictor"
    assert again.load_local_round_linker("tile0000", "bit003") == 2
    assert again.load_local_wavelengths_um("tile0000", bit="bit005") == (0.638, 0.67)
    stage, cam = again.load_local_stage_position_zyx_um("tile0000", round=0)
    np.testing.assert_allclose(stage, (1, 2, 3))
    np.testing.assert_allclose(cam, np.eye(4))
    np.testing.assert_allclose(again.load_local_round_transform_zyx_um("tile0000", "round002"), xf, rtol=1e-6)
    aff, org, spc = again.load_global_coord_xforms_um("tile0000")
    np.testing.assert_allclose(org, (4, 5, 6))
    assert again.load_global_coord_xforms_um("tile0001") == (None, None, None)
    f2, fattrs = again.load_local_sofima_flow_field("tile0000", "round002")
    np.testing.assert_array_equal(f2, flow)
    assert fattrs["block_size"] == [8, 8, 8]
    assert again.load_local_sofima_flow_field("tile0000", "round001") is None
    assert not again.has_identity_decode_transforms
    np.testing.assert_array_equal(again.load_chromatic_affine_transform_zyx_um(wavelength_um=0.67), np.eye(4))
    # decon_data wins over corrected_data (DS:4709-4745)
    zo.write_ome_image(root / "readouts/tile0001/bit001/decon_data.ome.zarr", ro[0] + 1, chunks=(4, 16, 16))
    np.testing.assert_array_equal(np.asarray(again.load_local_readout_image("tile0001", "bit001")), ro[0] + 1)
    # decode-stage outputs keep working on this store (inherited, same files as the reference)
    again.save_decode_normalization_vectors(None, "global", np.ones(8), np.zeros(8))
    nv, bv = zs.Qi2labZarrDataStore(root).load_decode_normalization_vectors(None, "global")
    np.testing.assert_array_equal(nv, np.ones(8, dtype=np.float32))


def test_store_written_in_reference_conventions_by_hand(tmp_path):
    """A store assembled file by file the way the reference's conversion stage leaves it (metadata split between image
    extra attributes and sidecars, DS:1860-1901) opens without this package's writer."""
    root = tmp_path / "qi2labdatastore"
    cb = _codebook(4)
    (root / "calibrations").mkdir(parents=True)
    (root / "datastore_state.json").write_text(json.dumps({"Version": 0.6, "Initialized": True, "Calibrations": True,
                                                           "Corrected": True}))
    (root / "calibrations" / "attributes.json").write_text(json.dumps({
        "num_rounds": 2, "num_tiles": 1, "num_bits": 4, "codebook": cb.to_numpy(dtype=object).tolist(),
        "voxel_size_zyx_um": [0.315, 0.098, 0.098], "microscope_type": "3D",
        "chromatic_affine_transforms_zyx_um": {"channels": {"far_red": {
            "channel_index": 2, "wavelength_um": 0.67, "affine_zyx_um": (np.eye(4) * 1.001).tolist()}}}}))
    rng = np.random.default_rng(2)
    vols = rng.poisson(90, (4, 18, 40, 36)).astype(np.uint16)
    for b in range(4):
        d = root / "readouts" / "tile0000" / f"bit{b + 1:03d}"
        zo.write_ome_image(d / "corrected_data.ome.zarr", vols[b], extra_attributes={
            "round_linker": 1 + b // 2, "excitation_um": 0.638, "psf_idx": 2})
        (d / "attributes.json").write_text(json.dumps({"emission_um": 0.67}))  # the sidecar completes / overrides
    for r in (1, 2):
        d = root / "fiducial" / "tile0000" / f"round{r:03d}"
        d.mkdir(parents=True)
        (d / "attributes.json").write_text(json.dumps({
            "stage_zyx_um": [0, 10, 20], "affine_zyx_px": np.eye(4).tolist(),
            "local_round_transform_zyx_um": np.eye(4).tolist(), "bit_linker": [2 * r - 1, 2 * r]}))
    ds = zs.Qi2labZarrDataStore(root, validate=True)
    assert ds.bit_ids == ["bit001", "bit002", "bit003", "bit004"]
    assert ds.load_local_wavelengths_um(0, bit=3) == (0.638, 0.67)
    assert ds.load_local_round_linker(0, 3) == 2
    np.testing.assert_array_equal(np.asarray(ds.load_local_readout_image(0, 2)), vols[2])
    np.testing.assert_allclose(ds.load_chromatic_affine_transform_zyx_um(wavelength_um=0.67), np.eye(4) * 1.001)
    np.testing.assert_array_equal(ds.load_chromatic_affine_transform_zyx_um(wavelength_um=0.52), np.eye(4))
    assert not ds.has_identity_decode_transforms
    with pytest.raises(FileNotFoundError):
        zs.Qi2labZarrDataStore(tmp_path / "elsewhere")


@pytest.mark.parametrize("shape,chunks", [((37, 150, 140), (16, 64, 64)), ((3, 6, 40, 44), (1, 4, 16, 16)),
                                          ((50, 60), (32, 32)), ((2, 3, 9, 20, 20), (1, 1, 4, 8, 8))])
def test_vectorised_chunk_table_equals_the_record_list(tmp_path, shape, chunks):
    import ctypes

    zo.write_ome_image(tmp_path / "img.ome.zarr", np.zeros(shape, dtype=np.float32), chunks=chunks)
    arr = zs.ZarrImage(tmp_path / "img.ome.zarr").array
    vz = arr.volume_shape[0]
    for z0, z1 in ((0, vz), (1, vz - 1), (vz // 2, vz // 2 + 1)):
        if not 0 <= z0 < z1 <= vz:
            continue
        fast = arr.chunk_table(0x10000, z0, z1, piece=3)
        slow = _capi.ChunkTable.from_dicts(arr.chunk_records(0x10000, z0, z1, piece=3))
        assert len(fast) == len(slow) > 0
        for name in _capi.ZARR_CHUNK_DTYPE.names:
            if name != "path":
                np.testing.assert_array_equal(fast.records[name], slow.records[name], err_msg=name)
        paths = [[ctypes.string_at(int(p)).decode() for p in t.records["path"]] for t in (fast, slow)]
        assert paths[0] == paths[1]


def test_corrupted_lz4_and_zstd_frames_never_escape_their_buffers():
    """The LZ4 parser is the library's own (shared with the device kernel): flip bytes anywhere in valid frames and
    the decoder must either report an error or return exactly nbytes bytes -- guard pages around NumPy buffers are not
    available, so this is a crash / hang test, plus a check that untouched frames still decode."""
    rng = np.random.default_rng(42)
    data = (rng.poisson(5, 40000) * (rng.random(40000) > 0.5)).astype(np.uint16).tobytes()
    for cname in ("lz4", "zstd"):
        for split in (True, False):
            frame = bytearray(zo.blosc_compress(data, 2, cname=cname, blocksize=8192, split=split))
            assert _capi.blosc_decode_host(bytes(frame)) == data
            for _ in range(300):
                bad = bytearray(frame)
                for _k in range(int(rng.integers(1, 4))):
                    bad[int(rng.integers(16, len(bad)))] = int(rng.integers(0, 256))
                try:
                    out = _capi.blosc_decode_host(bytes(bad))
                    assert len(out) == len(data)
                except _capi.M3dError:
                    pass


def _zstd_datasets():
    rng = np.random.default_rng(0)
    yield "zeros", bytes(100000)
    yield "one", b"x"
    yield "short", b"hello world hello world hello"
    yield "text", b"the quick brown fox jumps over the lazy dog. " * 3000
    yield "random", rng.integers(0, 256, 200000, dtype=np.uint8).tobytes()
    yield "low_entropy", rng.integers(0, 4, 300000, dtype=np.uint8).tobytes()
    yield "poisson_u16", rng.poisson(100, 150000).astype(np.uint16).tobytes()
    a = (rng.poisson(100, 131072) + 100).astype(np.uint16)
    a[1000:1200] += 3000
    yield "bitshuffled_block", zo.bit_shuffle(a.tobytes(), 2)  # what a Blosc block of a readout image holds
    yield "mixed", b"".join([bytes(5000), rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(), b"abc" * 4000,
                             rng.integers(0, 16, 70000, dtype=np.uint8).tobytes()])
    yield "multi_block", rng.poisson(3, 1 << 20).astype(np.uint8).tobytes()  # 8 blocks: repeat tables / offsets
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(500)]
    yield "words", b" ".join(words[int(i)] for i in rng.integers(0, 500, 60000))


def test_builtin_zstd_decoder_equals_libzstd():
    """csrc/zstd_decode.cuh (the decoder that is to move onto the device) against frames written by the system
    libzstd at fast, default, high and negative levels and by pyarrow's bundled zstd: raw / RLE / compressed blocks,
    Huffman literals in 1 and 4 streams with direct and FSE-coded weights, predefined / RLE / described / repeated
    sequence tables, repeat offsets."""
    for name, data in _zstd_datasets():
        for level in (-5, 1, 3, 5, 9, 15, 19):
            frame = _capi.zstd_host(data, True, level=level)
            assert _capi.zstd_decode_builtin(frame, len(data)) == data, (name, level)
            assert _capi.zstd_decode_builtin(frame, len(data), lanes=True) == data, (name, level, "lanes")
        for lanes in (False, True):
            assert _capi.zstd_decode_builtin(zo._compress("zstd", data, 3), len(data), lanes=lanes) == data, name
    assert _capi.zstd_decode_builtin(_capi.zstd_host(b"", True), 0) == b""
    with pytest.raises(_capi.M3dError):  # does not fit
        _capi.zstd_decode_builtin(_capi.zstd_host(bytes(1000), True), 999)


def test_builtin_zstd_decoder_survives_corruption():
    rng = np.random.default_rng(7)
    data = b"the quick brown fox " * 2000 + rng.integers(0, 8, 30000, dtype=np.uint8).tobytes()
    frame = bytearray(_capi.zstd_host(data, True, level=5))
    reported = 0
    for _ in range(1500):
        bad = bytearray(frame)
        for _k in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
        outcomes = []
        for lanes in (False, True):  # the two arrangements of the decoder agree on what is damage
            try:
                out = _capi.zstd_decode_builtin(bytes(bad), len(data), lanes=lanes)
                assert len(out) <= len(data)
                outcomes.append(out)
            except _capi.M3dError:
                outcomes.append(None)
        assert outcomes[0] == outcomes[1]
        reported += outcomes[0] is None
    assert reported > 500
    for cut in (0, 3, 5, 9, len(frame) // 2, len(frame) - 1):
        with pytest.raises(_capi.M3dError):
            _capi.zstd_decode_builtin(bytes(frame[:cut]), len(data))


def test_run_scoped_normalization_metadata_round_trips_on_both_stores(tmp_path):
    """Mirror of the reference's tests/test_optimization_codeword_exclusions.py:275-299 (DS:1179-1271)."""
    from merfish3d_analysis_b200.datastore import ArrayDataStore

    cb = _codebook(4)
    stores = [ArrayDataStore(tmp_path / "a" / "qi2labdatastore", codebook=cb),
              zs.Qi2labZarrDataStore.create(tmp_path / "z" / "qi2labdatastore", cb)]
    metadata = {"scope": "iterative_optimization", "excluded_gene_ids": ["GeneB"], "codebook_sha256": "abc123"}
    for ds in stores:
        ds.save_decode_normalization_vectors("run1", "iterative", np.ones(4, dtype=np.float32),
                                             np.zeros(4, dtype=np.float32), decode_mode="3d", metadata=metadata)
        assert ds.load_decode_normalization_metadata("run1", "iterative") == metadata
        assert ds.load_decode_normalization_metadata(None, "iterative") is None
        n, b = type(ds)(ds._datastore_path).load_decode_normalization_vectors("run1", "iterative")
        np.testing.assert_array_equal(n, np.ones(4, dtype=np.float32))
        np.testing.assert_array_equal(b, np.zeros(4, dtype=np.float32))
        with pytest.raises(ValueError):
            ds.save_decode_normalization_vectors("bad/key", "iterative", np.ones(4), np.zeros(4))
