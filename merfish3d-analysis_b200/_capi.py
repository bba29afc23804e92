"""ctypes binding of ``libm3d_b200.so`` (the C ABI declared in ``include/m3d_b200.h``).

PyTorch is used only for device memory and streams: every array crossing this boundary is
a raw device pointer taken from a torch tensor.  There is NO CPU fallback -- if the shared
library is missing or no CUDA device is visible, calls raise.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libm3d_b200.so"

M3D_DTYPE_U16 = 0
M3D_DTYPE_F32 = 1
M3D_TABLE_FIXED_COLS = 14
M3D_MAX_BITS = 32

_c_i64_3 = C.c_int64 * 3
_c_f64_3 = C.c_double * 3

# name -> (restype, argtypes); must list every symbol the header declares
SIGNATURES = {
    "m3d_abi_version": (C.c_int, []),
    "m3d_last_error": (C.c_char_p, []),
    "m3d_create": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)],
    ),
    "m3d_destroy": (C.c_int, [C.c_void_p]),
    "m3d_set_normalization": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "m3d_set_thresholds": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "m3d_set_lowpass_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "m3d_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "m3d_upload_batch": (
        C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p]
    ),
    "m3d_upload_batch_cb": (
        C.c_int,
        [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "m3d_zarr_read_chunks": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "m3d_zarr_read_chunks_host": (C.c_int, [C.c_int, C.c_void_p]),
    "m3d_blosc_info": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "m3d_blosc_decode_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "m3d_blosc_encode_bound": (C.c_int64, [C.c_int64, C.c_int64]),
    "m3d_blosc_encode_host": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64,
         C.POINTER(C.c_int64)],
    ),
    "m3d_zstd_host": (
        C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    ),
    "m3d_zstd_decode_builtin": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "m3d_zstd_decode_builtin_lanes": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "m3d_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "m3d_warp_affine": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double),
         C.POINTER(C.c_double), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p],
    ),
    "m3d_warp_flow": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float),
         C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float),
         C.POINTER(C.c_int64), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p],
    ),
    "m3d_lowpass": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int64),
         C.POINTER(C.c_double), C.c_int, C.c_void_p, C.c_void_p],
    ),
    "m3d_decode": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "m3d_label": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_double, C.c_int,
         C.c_void_p, C.POINTER(C.c_int64), C.c_void_p],
    ),
    "m3d_decode_label": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_int, C.c_double,
         C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p],
    ),
    "m3d_decode_label_persistent": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_int, C.c_double,
         C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p],
    ),
    "m3d_interface_pairs": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
         C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p],
    ),
    "m3d_features": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_int,
         C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "m3d_select_hist": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_int, C.c_float,
         C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p],
    ),
    "m3d_select_hist_batch": (
        C.c_int,
        [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_uint32),
         C.POINTER(C.c_uint32), C.c_int, C.c_void_p, C.c_void_p],
    ),
    "m3d_replace_above": (
        C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p]
    ),
    "m3d_table_hist3d": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
         C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "m3d_overlap_duplicates": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p],
    ),
    "m3d_within_tile_duplicates": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double,
         C.c_void_p, C.c_void_p],
    ),
    "m3d_centroid_statistics": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_int64,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "m3d_assign_cells": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
         C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p],
    ),
    "m3d_inertia_eigvals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "m3d_set_sparse_capacity": (C.c_int, [C.c_void_p, C.c_int64]),
    "m3d_launch_count": (C.c_int64, [C.c_void_p]),
    "m3d_kernel_name": (C.c_char_p, [C.c_int]),
    "m3d_kernel_launches": (C.c_int64, [C.c_void_p, C.c_int]),
    "m3d_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "m3d_reset_counters": (C.c_int, [C.c_void_p]),
    "m3d_kernel_time_ms": (C.c_double, [C.c_void_p, C.c_int]),
}

_PIECE_CB = C.CFUNCTYPE(None, C.c_int, C.c_void_p)

M3D_ZARR_RAW, M3D_ZARR_BLOSC, M3D_ZARR_ZSTD, M3D_ZARR_ABSENT = 0, 1, 2, 3


class ZarrChunk(C.Structure):
    """``m3d_zarr_chunk`` (include/m3d_b200.h)."""

    _fields_ = [
        ("path", C.c_char_p),
        ("offset", C.c_int64),
        ("length", C.c_int64),
        ("codec", C.c_int32),
        ("elem_size", C.c_int32),
        ("chunk_shape", C.c_int64 * 3),
        ("origin", C.c_int64 * 3),
        ("dst", C.c_void_p),
        ("dst_shape", C.c_int64 * 3),
        ("fill_bits", C.c_uint64),
        ("piece", C.c_int32),
        ("reserved", C.c_int32),
    ]


# the same record as a NumPy structured dtype: chunk tables of whole tiles are built with array arithmetic
ZARR_CHUNK_DTYPE = np.dtype({
    "names": ["path", "offset", "length", "codec", "elem_size", "chunk_shape", "origin", "dst", "dst_shape",
              "fill_bits", "piece", "reserved"],
    "formats": ["<u8", "<i8", "<i8", "<i4", "<i4", ("<i8", 3), ("<i8", 3), "<u8", ("<i8", 3), "<u8", "<i4", "<i4"],
    "offsets": [0, 8, 16, 24, 28, 32, 56, 80, 88, 112, 120, 124],
    "itemsize": 128,
})
assert C.sizeof(ZarrChunk) == ZARR_CHUNK_DTYPE.itemsize


class ChunkTable:
    """``m3d_zarr_chunk[n]`` as a structured array + the NUL-terminated path strings its pointers refer to."""

    def __init__(self, records: np.ndarray, path_blobs: list):
        self.records = records
        self.path_blobs = path_blobs  # keep-alive

    def __len__(self):
        return int(self.records.shape[0])

    @staticmethod
    def from_paths(paths: list[str], n: int) -> "ChunkTable":
        """n zeroed records whose ``path`` fields point at ``paths`` (len n)."""
        enc = [p.encode() for p in paths]
        blob = np.frombuffer(b"\0".join(enc) + b"\0", dtype=np.uint8)
        starts = np.zeros(n, dtype=np.uint64)
        if n > 1:
            starts[1:] = np.cumsum([len(e) + 1 for e in enc[:-1]], dtype=np.uint64)
        rec = np.zeros(n, dtype=ZARR_CHUNK_DTYPE)
        rec["path"] = np.uint64(blob.ctypes.data) + starts
        return ChunkTable(rec, [blob])

    @staticmethod
    def from_dicts(chunks) -> "ChunkTable":
        t = ChunkTable.from_paths([c["path"] for c in chunks], len(chunks))
        r = t.records
        for k in ("offset", "length", "codec", "elem_size", "dst", "fill_bits", "piece"):
            default = {"offset": 0, "length": -1, "fill_bits": 0, "piece": 0}.get(k)
            r[k] = [c.get(k, default) for c in chunks]
        for k in ("chunk_shape", "origin", "dst_shape"):
            r[k] = np.asarray([c[k] for c in chunks], dtype=np.int64).reshape(len(chunks), 3)
        return t

    @staticmethod
    def concatenate(tables) -> "ChunkTable":
        tables = [t for t in tables if len(t)]
        if not tables:
            return ChunkTable(np.zeros(0, dtype=ZARR_CHUNK_DTYPE), [])
        return ChunkTable(np.ascontiguousarray(np.concatenate([t.records for t in tables])),
                          [b for t in tables for b in t.path_blobs])


def _as_chunk_table(chunks) -> ChunkTable:
    return chunks if isinstance(chunks, ChunkTable) else ChunkTable.from_dicts(list(chunks))


def zarr_read_chunks_host(chunks) -> None:
    """Decode chunks into HOST destinations (``dst`` = host addresses); needs no GPU."""
    t = _as_chunk_table(chunks)
    if len(t):
        _check(load_library().m3d_zarr_read_chunks_host(len(t), C.c_void_p(t.records.ctypes.data)),
               "m3d_zarr_read_chunks_host")


def blosc_info(frame: bytes) -> dict:
    out = (C.c_int64 * 6)()
    buf = np.frombuffer(frame, dtype=np.uint8)
    _check(load_library().m3d_blosc_info(buf.ctypes.data, buf.size, out), "m3d_blosc_info")
    return dict(zip(("nbytes", "blocksize", "cbytes", "typesize", "flags", "codec"), [int(v) for v in out]))


def blosc_decode_host(frame: bytes) -> bytes:
    info = blosc_info(frame)
    buf = np.frombuffer(frame, dtype=np.uint8)
    out = np.empty(max(info["nbytes"], 1), dtype=np.uint8)
    _check(load_library().m3d_blosc_decode_host(buf.ctypes.data, buf.size, out.ctypes.data, out.size),
           "m3d_blosc_decode_host")
    return out[: info["nbytes"]].tobytes()


def blosc_encode_host(data, typesize: int, cname: str = "zstd", clevel: int = 5, shuffle: str = "bitshuffle",
                      blocksize: int = 0) -> bytes:
    lib = load_library()
    src = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    src = np.ascontiguousarray(src)
    cap = int(lib.m3d_blosc_encode_bound(src.size, blocksize))
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_int64(0)
    _check(lib.m3d_blosc_encode_host(src.ctypes.data, src.size, int(typesize), {"zstd": 4, "lz4": 1}[cname], int(clevel),
                                     {"noshuffle": 0, "shuffle": 1, "bitshuffle": 2}[shuffle], int(blocksize),
                                     out.ctypes.data, cap, C.byref(n)), "m3d_blosc_encode_host")
    return out[: n.value].tobytes()


def zstd_decode_builtin(frame, capacity: int, lanes: bool = False) -> bytes:
    """One zstd frame through the library's own decoder (``csrc/zstd_decode.cuh``), not libzstd; ``lanes=True`` takes
    the team-of-lanes arrangement (``csrc/zstd_lanes.cuh``) with a team of one."""
    src = np.ascontiguousarray(np.frombuffer(frame, dtype=np.uint8))
    out = np.empty(max(int(capacity), 1), dtype=np.uint8)
    n = C.c_int64(0)
    lib = load_library()
    fn = lib.m3d_zstd_decode_builtin_lanes if lanes else lib.m3d_zstd_decode_builtin
    _check(fn(src.ctypes.data, src.size, out.ctypes.data, int(capacity), C.byref(n)), "m3d_zstd_decode_builtin")
    return out[: n.value].tobytes()


def zstd_host(data, compress: bool, decoded_size: int = 0, level: int = 3) -> bytes:
    src = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8))
    cap = int(src.size + src.size // 128 + 1024) if compress else int(decoded_size)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = C.c_int64(0)
    _check(load_library().m3d_zstd_host(int(bool(compress)), src.ctypes.data, src.size, out.ctypes.data, cap, int(level),
                                        C.byref(n)), "m3d_zstd_host")
    return out[: n.value].tobytes()
_lib = None


class M3dError(RuntimeError):
    """Raised when a libm3d_b200 call returns a negative status."""


def load_library():
    """dlopen the in-tree shared library and attach signatures.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise M3dError(
            f"{LIB_PATH} is missing: build it with `python -m merfish3d_analysis_b200.build` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().m3d_last_error()
        raise M3dError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def _ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise M3dError("libm3d_b200 takes device memory only (no CPU fallback)")
    if not t.is_contiguous():
        raise M3dError("libm3d_b200 takes contiguous tensors")
    return C.c_void_p(t.data_ptr())


def _stream(device):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dtype_code(t) -> int:
    import torch

    if t.dtype == torch.uint16 or t.dtype == torch.int16:
        return M3D_DTYPE_U16
    if t.dtype == torch.float32:
        return M3D_DTYPE_F32
    raise M3dError(f"unsupported stack dtype {t.dtype}; expected uint16 or float32")


class DecodeContext:
    """One ``m3d_ctx`` per process/GPU: resident codebook, thresholds, scratch."""

    def __init__(self, codebook_unit: np.ndarray, excluded=(), device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise M3dError("no CUDA device visible: the B200 decode path has no CPU fallback")
        self._lib = load_library()
        cb = np.ascontiguousarray(codebook_unit, dtype=np.float32)
        if cb.ndim != 2:
            raise ValueError("codebook_unit must be (K, bits)")
        self.n_codewords, self.n_bits = int(cb.shape[0]), int(cb.shape[1])
        ex = np.ascontiguousarray(np.asarray(list(excluded), dtype=np.int32))
        self.device = torch.device("cuda", int(device))
        handle = C.c_void_p()
        _check(
            self._lib.m3d_create(
                int(device), self.n_bits, self.n_codewords, cb.ctypes.data_as(C.c_void_p),
                ex.ctypes.data_as(C.c_void_p) if ex.size else None, int(ex.size), C.byref(handle),
            ),
            "m3d_create",
        )
        self._h = handle
        self._n_features = -1
        import os

        accum = os.environ.get("M3D_LOWPASS_ACCUM")
        if accum:  # "float32": the opt-in CuPy-style low-pass arithmetic (see set_lowpass_accumulate)
            self.set_lowpass_accumulate(accum)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.m3d_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ state
    def set_normalization(self, background, normalization):
        if background is None or normalization is None:
            _check(self._lib.m3d_set_normalization(self._h, None, None), "m3d_set_normalization")
            return
        b = np.ascontiguousarray(np.asarray(background, dtype=np.float32)[: self.n_bits])
        n = np.ascontiguousarray(np.asarray(normalization, dtype=np.float32)[: self.n_bits])
        if b.size != self.n_bits or n.size != self.n_bits:
            raise ValueError("normalisation vectors shorter than the bit count")
        _check(
            self._lib.m3d_set_normalization(
                self._h, b.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p)
            ),
            "m3d_set_normalization",
        )

    def set_thresholds(self, pixel_threshold: float, magnitude_lo: float, magnitude_hi: float):
        # the reference compares float32 arrays with weak Python scalars (NEP 50): thresholds
        # act as float32 values
        _check(
            self._lib.m3d_set_thresholds(
                self._h, float(np.float32(pixel_threshold)), float(np.float32(magnitude_lo)),
                float(np.float32(magnitude_hi)),
            ),
            "m3d_set_thresholds",
        )

    # ------------------------------------------------------------------ kernels
    @staticmethod
    def _dims(shape_zyx):
        return _c_i64_3(*[int(s) for s in shape_zyx])

    def upload(self, pieces, on_piece=None):
        """Host -> device copies of ``[(numpy C-contiguous array, device tensor of the same byte size), ...]``
        through the library's pinned staging ring (pageable sources) or straight DMA (pinned sources).
        Returns once the sources are no longer needed; the copies complete on the current stream.
        ``on_piece(i)`` is called (on this thread) as soon as piece i is completely enqueued on the stream,
        while later pieces are still being staged."""
        pieces = list(pieces)
        if not pieces:
            return
        n = len(pieces)
        srcs, dsts, sizes = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int64 * n)()
        for i, (src, dst) in enumerate(pieces):
            if not src.flags.c_contiguous:
                raise M3dError("upload() takes C-contiguous host arrays")
            if src.nbytes != dst.numel() * dst.element_size() or not dst.is_cuda or not dst.is_contiguous():
                raise M3dError("upload(): destination must be a contiguous device tensor of the source's byte size")
            srcs[i], dsts[i], sizes[i] = src.ctypes.data, dst.data_ptr(), src.nbytes
        if on_piece is None:
            _check(self._lib.m3d_upload_batch(self._h, n, srcs, dsts, sizes, _stream(self.device)), "m3d_upload_batch")
            return
        failure = []

        def trampoline(piece, _user):
            if failure:
                return
            try:
                on_piece(int(piece))
            except BaseException as e:  # noqa: BLE001 - re-raised below, ctypes would swallow it
                failure.append(e)

        cb = _PIECE_CB(trampoline)
        rc = self._lib.m3d_upload_batch_cb(self._h, n, srcs, dsts, sizes, _stream(self.device),
                                           C.cast(cb, C.c_void_p), None)
        if failure:
            raise failure[0]
        _check(rc, "m3d_upload_batch_cb")

    def zarr_read(self, chunks, on_piece=None):
        """``m3d_zarr_read_chunks``: decode the listed chunks (dicts with the ``m3d_zarr_chunk`` fields, ``dst`` a
        device address) into their device volumes on the current stream.  ``on_piece(i)`` as in :meth:`upload`."""
        t = _as_chunk_table(chunks)
        if not len(t):
            return
        failure = []

        def trampoline(piece, _user):
            if failure or on_piece is None:
                return
            try:
                on_piece(int(piece))
            except BaseException as e:  # noqa: BLE001 - re-raised below, ctypes would swallow it
                failure.append(e)

        cb = _PIECE_CB(trampoline)
        rc = self._lib.m3d_zarr_read_chunks(self._h, len(t), C.c_void_p(t.records.ctypes.data), _stream(self.device),
                                            C.cast(cb, C.c_void_p), None)
        if failure:
            raise failure[0]
        _check(rc, "m3d_zarr_read_chunks")

    def weight(self, readout, predictor, out=None):
        import torch

        if out is None:
            out = torch.empty(readout.shape, dtype=torch.float32, device=readout.device)
        _check(
            self._lib.m3d_weight(self._h, _ptr(readout), _ptr(predictor), readout.numel(),
                                 _ptr(out), _stream(self.device)),
            "m3d_weight",
        )
        return out

    def warp_affine(self, volume, matrix_px, offset_px, predictor=None, out_z0: int = 0, out_nz=None, out=None):
        """scipy-style order-1 affine resampling of one (z, y, x) volume -> float32 planes."""
        import torch

        if volume.dim() != 3:
            raise ValueError("volume must be (z, y, x)")
        nz = int(volume.shape[0]) if out_nz is None else int(out_nz)
        if out is None:
            out = torch.empty((nz, *volume.shape[1:]), dtype=torch.float32, device=volume.device)
        m = (C.c_double * 9)(*[float(v) for v in np.asarray(matrix_px, dtype=np.float64).reshape(9)])
        o = (C.c_double * 3)(*[float(v) for v in np.asarray(offset_px, dtype=np.float64).reshape(3)])
        _check(
            self._lib.m3d_warp_affine(self._h, _ptr(volume), _dtype_code(volume), _ptr(predictor),
                                      self._dims(volume.shape), m, o, int(out_z0), nz, _ptr(out),
                                      _stream(self.device)),
            "m3d_warp_affine",
        )
        return out

    def warp_flow(self, volume, transform_zyx_um, spacing_zyx_um, flow, stride_zyx, box_start_zyx, out_shape,
                  predictor=None, origin_zyx_um=(0.0, 0.0, 0.0), out_z0: int = 0, out_nz=None, out=None):
        """Affine + SOFIMA flow warp of one (z, y, x) volume; ``flow`` = (3, fz, fy, fx) float32 device tensor."""
        import torch

        if volume.dim() != 3 or flow.dim() != 4 or flow.shape[0] != 3:
            raise ValueError("volume must be (z, y, x) and flow (3, fz, fy, fx)")
        nz = int(out_shape[0]) if out_nz is None else int(out_nz)
        if out is None:
            out = torch.empty((nz, int(out_shape[1]), int(out_shape[2])), dtype=torch.float32, device=volume.device)
        f3 = C.c_float * 3
        t = (C.c_float * 16)(*[float(v) for v in np.asarray(transform_zyx_um, dtype=np.float32).reshape(16)])
        _check(
            self._lib.m3d_warp_flow(
                self._h, _ptr(volume), _dtype_code(volume), _ptr(predictor), self._dims(volume.shape), t,
                f3(*[float(np.float32(v)) for v in spacing_zyx_um]), f3(*[float(np.float32(v)) for v in origin_zyx_um]),
                _ptr(flow), self._dims(flow.shape[1:]), f3(*[float(np.float32(v)) for v in stride_zyx]),
                f3(*[float(np.float32(v)) for v in box_start_zyx]), self._dims(out_shape), int(out_z0), nz, _ptr(out),
                _stream(self.device),
            ),
            "m3d_warp_flow",
        )
        return out

    def set_lowpass_accumulate(self, kind: str) -> None:
        """``"float64"`` (default: SciPy's arithmetic, pinned by the reference-generated goldens) or ``"float32"``
        (float32 weights + FMA accumulation as CuPy's filter is believed to compute; opt-in, not pinned)."""
        if kind not in ("float64", "float32"):
            raise ValueError("lowpass accumulation must be 'float64' or 'float32'")
        _check(self._lib.m3d_set_lowpass_mode(self._h, 1 if kind == "float32" else 0), "m3d_set_lowpass_mode")

    def lowpass(self, stack, sigma, mode2d: bool, predictor=None, out=None):
        """stack: (n_vols, z, y, x) uint16/float32 device tensor -> float32 low-passed."""
        import torch

        if stack.dim() != 4:
            raise ValueError("stack must be (n_vols, z, y, x)")
        if out is None:
            out = torch.empty(stack.shape, dtype=torch.float32, device=stack.device)
        sg = _c_f64_3(*[float(s) for s in sigma])
        _check(
            self._lib.m3d_lowpass(
                self._h, _ptr(stack), _dtype_code(stack), _ptr(predictor), int(stack.shape[0]),
                self._dims(stack.shape[1:]), sg, 1 if mode2d else 0, _ptr(out),
                _stream(self.device),
            ),
            "m3d_lowpass",
        )
        return out

    def decode(self, stack, decoded=None, magnitude=None, distance=None, scaled=None):
        """stack: (bits, z, y, x).  Writes int16 decoded (z, y, x); optional float16 images."""
        import torch

        if stack.dim() != 4 or stack.shape[0] != self.n_bits:
            raise ValueError(f"stack must be ({self.n_bits}, z, y, x)")
        if decoded is None:
            decoded = torch.empty(stack.shape[1:], dtype=torch.int16, device=stack.device)
        _check(
            self._lib.m3d_decode(
                self._h, _ptr(stack), _dtype_code(stack), self._dims(stack.shape[1:]),
                _ptr(decoded), _ptr(magnitude), _ptr(distance), _ptr(scaled),
                _stream(self.device),
            ),
            "m3d_decode",
        )
        return decoded

    def label(self, decoded, mode2d: bool, minimum_pixels: float, maximum_pixels: int = 500,
              labels=None) -> int:
        n = C.c_int64(-1)
        _check(
            self._lib.m3d_label(
                self._h, _ptr(decoded), self._dims(decoded.shape), 1 if mode2d else 0,
                float(minimum_pixels), int(maximum_pixels), _ptr(labels), C.byref(n),
                _stream(self.device),
            ),
            "m3d_label",
        )
        self._n_features = int(n.value)
        return self._n_features

    def decode_label(self, stack, decoded, mode2d: bool, minimum_pixels: float,
                     maximum_pixels: int = 500, labels=None, persistent: bool = False) -> int:
        """Fused production path: decode (no result images) + connected components.  ``persistent``: ``decoded``
        is a buffer kept across calls and untouched in between (see m3d_decode_label_persistent)."""
        if stack.dim() != 4 or stack.shape[0] != self.n_bits:
            raise ValueError(f"stack must be ({self.n_bits}, z, y, x)")
        n = C.c_int64(-1)
        fn = self._lib.m3d_decode_label_persistent if persistent else self._lib.m3d_decode_label
        _check(
            fn(
                self._h, _ptr(stack), _dtype_code(stack), self._dims(stack.shape[1:]), _ptr(decoded),
                1 if mode2d else 0, float(minimum_pixels), int(maximum_pixels), _ptr(labels),
                C.byref(n), _stream(self.device),
            ),
            "m3d_decode_label",
        )
        self._n_features = int(n.value)
        return self._n_features

    def interface_pairs(self, dec_lo, lab_lo, dec_hi, lab_hi):
        """(label_hi, label_lo) pairs joined across a z-slab interface, de-duplicated (int32 (n, 2))."""
        import torch

        Y, X = int(dec_hi.shape[-2]), int(dec_hi.shape[-1])
        cap = 1 << 16
        while True:
            pairs = torch.empty((cap, 2), dtype=torch.int32, device=dec_hi.device)
            n = C.c_int64(0)
            _check(
                self._lib.m3d_interface_pairs(self._h, _ptr(dec_lo), _ptr(lab_lo), _ptr(dec_hi), _ptr(lab_hi),
                                              Y, X, _ptr(pairs), cap, C.byref(n), _stream(self.device)),
                "m3d_interface_pairs",
            )
            if n.value <= cap:
                break
            cap = int(n.value)
        if n.value == 0:
            return np.zeros((0, 2), dtype=np.int32)
        return np.unique(pairs[: n.value].cpu().numpy(), axis=0)

    def features(self, stack, decoded, optimize_mode: bool, n_rows: int | None = None):
        """Feature table (n_rows, 14 + bits) float64 on the device for the last ``label``."""
        import torch

        n = self._n_features if n_rows is None else int(n_rows)
        if n < 0:
            raise M3dError("features() before label()")
        table = torch.empty((n, M3D_TABLE_FIXED_COLS + self.n_bits), dtype=torch.float64,
                            device=stack.device)
        if n == 0:
            return table
        _check(
            self._lib.m3d_features(
                self._h, _ptr(stack), _dtype_code(stack), self._dims(stack.shape[1:]),
                _ptr(decoded), 1 if optimize_mode else 0, _ptr(table), n, _stream(self.device),
            ),
            "m3d_features",
        )
        return table

    def select_hist(self, data, hist, sub=0.0, clip0=False, pred=0, cutoff=0.0, prefix_mask=0,
                    prefix_value=0, shift=21):
        _check(
            self._lib.m3d_select_hist(
                self._h, _ptr(data), data.numel(), float(sub), 1 if clip0 else 0, int(pred),
                float(cutoff), int(prefix_mask), int(prefix_value), int(shift), _ptr(hist),
                _stream(self.device),
            ),
            "m3d_select_hist",
        )

    def select_hist_batch(self, rows, hist, shift: int):
        """One launch for a whole level of a multi-query radix select: ``rows`` = list of
        ``(data tensor, prefix_mask, prefix_value)``; row r's digit histogram is added to ``hist[r]``
        (``hist``: zeroed (len(rows), 2048) int64 device tensor).  NaN values are not counted."""
        n = len(rows)
        if n == 0:
            return
        ptrs = (C.c_void_p * n)(*[(_ptr(d) if d.numel() else None) for d, _pm, _pv in rows])
        sizes = (C.c_int64 * n)(*[int(d.numel()) for d, _pm, _pv in rows])
        pms = (C.c_uint32 * n)(*[int(pm) & 0xFFFFFFFF for _d, pm, _pv in rows])
        pvs = (C.c_uint32 * n)(*[int(pv) & 0xFFFFFFFF for _d, _pm, pv in rows])
        _check(
            self._lib.m3d_select_hist_batch(self._h, n, ptrs, sizes, pms, pvs, int(shift), _ptr(hist),
                                            _stream(self.device)),
            "m3d_select_hist_batch",
        )

    def replace_above(self, data, threshold: float, value: float):
        _check(
            self._lib.m3d_replace_above(self._h, _ptr(data), data.numel(), float(threshold),
                                        float(value), _stream(self.device)),
            "m3d_replace_above",
        )

    # ------------------------------------------------------------------ post-decode table stage
    def table_hist3d(self, v0, v1, v2, blank, edges0, edges1, edges2):
        """Row bins + all / blank histograms of the blank-fraction filter (PD:3656-3742).
        v*: float32 device vectors, blank: uint8 device vector, edges*: float32 host arrays.
        Returns (flat_bin int32 (n,), all_hist int32, blank_hist int32) device tensors."""
        import torch

        n = int(v0.numel())
        es = [np.ascontiguousarray(e, dtype=np.float32) for e in (edges0, edges1, edges2)]
        shape = tuple(int(e.size) - 1 for e in es)
        flat = torch.empty(n, dtype=torch.int32, device=self.device)
        all_h = torch.zeros(shape, dtype=torch.int32, device=self.device)
        blank_h = torch.zeros(shape, dtype=torch.int32, device=self.device)
        _check(
            self._lib.m3d_table_hist3d(
                self._h, _ptr(v0), _ptr(v1), _ptr(v2), _ptr(blank), n,
                es[0].ctypes.data_as(C.c_void_p), int(es[0].size), es[1].ctypes.data_as(C.c_void_p), int(es[1].size),
                es[2].ctypes.data_as(C.c_void_p), int(es[2].size), _ptr(flat), _ptr(all_h), _ptr(blank_h),
                _stream(self.device),
            ),
            "m3d_table_hist3d",
        )
        return flat, all_h, blank_h

    def overlap_duplicates(self, zyx, tile, distance_min, radius: float):
        """uint8 drop flags of ``_remove_duplicates_in_tile_overlap`` (PD:4137-4177)."""
        import torch

        n = int(tile.numel())
        drop = torch.zeros(n, dtype=torch.uint8, device=self.device)
        _check(
            self._lib.m3d_overlap_duplicates(self._h, _ptr(zyx), _ptr(tile), _ptr(distance_min), n, float(radius),
                                             _ptr(drop), _stream(self.device)),
            "m3d_overlap_duplicates",
        )
        return drop

    def within_tile_duplicates(self, zyx, tile, gene, distance_min, radius_xy: float, radius_z: float):
        """uint8 drop flags of ``_remove_duplicates_within_tile`` (PD:4179-4363)."""
        import torch

        n = int(tile.numel())
        drop = torch.zeros(n, dtype=torch.uint8, device=self.device)
        _check(
            self._lib.m3d_within_tile_duplicates(self._h, _ptr(zyx), _ptr(tile), _ptr(gene), _ptr(distance_min), n,
                                                 float(radius_xy), float(radius_z), _ptr(drop),
                                                 _stream(self.device)),
            "m3d_within_tile_duplicates",
        )
        return drop

    def centroid_statistics(self, labels, stack, z_support: int, label_code):
        """Per-label, per-on-bit weighted centroid sums (PD:2833-2906) in one pass.
        labels: int32 (z,y,x) ids + 1; stack: (bits,z,y,x); label_code: int16 (minlength,) codeword row
        per label.  Returns (sums float64 (minlength, bits, 4) = {w, wz, wy, wx}, peak float32
        (minlength, bits)) device tensors."""
        import torch

        minlength = int(label_code.numel())
        sums = torch.empty((minlength, self.n_bits, 4), dtype=torch.float64, device=self.device)
        peak = torch.empty((minlength, self.n_bits), dtype=torch.float32, device=self.device)
        _check(
            self._lib.m3d_centroid_statistics(
                self._h, _ptr(labels), _ptr(stack), _dtype_code(stack), self._dims(labels.shape), int(z_support),
                _ptr(label_code), minlength, _ptr(sums), _ptr(peak), _stream(self.device),
            ),
            "m3d_centroid_statistics",
        )
        return sums, peak

    def assign_cells(self, yx, verts_yx, poly_offsets, bbox, cell_start, cell_polys, origin_yx, cell_size, grid_yx):
        """int32 cell ids (0 = none) of the points ``yx`` ((n, 2) float64 device tensor); see m3d_assign_cells."""
        import torch

        n = int(yx.shape[0])
        out = torch.zeros(n, dtype=torch.int32, device=self.device)
        _check(
            self._lib.m3d_assign_cells(
                self._h, _ptr(yx), n, _ptr(verts_yx), _ptr(poly_offsets), _ptr(bbox), _ptr(cell_start), _ptr(cell_polys),
                float(origin_yx[0]), float(origin_yx[1]), float(cell_size), int(grid_yx[0]), int(grid_yx[1]), _ptr(out),
                _stream(self.device),
            ),
            "m3d_assign_cells",
        )
        return out

    def inertia_eigvals(self, table):
        """(n, 3) float64 inertia-tensor eigenvalues (descending, clipped at 0) of a feature table."""
        import torch

        ev = torch.empty((int(table.shape[0]), 3), dtype=torch.float64, device=self.device)
        _check(
            self._lib.m3d_inertia_eigvals(self._h, _ptr(table), int(table.shape[0]), int(table.shape[1]), _ptr(ev),
                                          _stream(self.device)),
            "m3d_inertia_eigvals",
        )
        return ev

    def set_sparse_capacity(self, entries: int) -> None:
        _check(self._lib.m3d_set_sparse_capacity(self._h, int(entries)), "m3d_set_sparse_capacity")

    # ------------------------------------------------------------------ accounting
    def launch_count(self) -> int:
        return int(self._lib.m3d_launch_count(self._h))

    def set_timing(self, enable: bool) -> None:
        _check(self._lib.m3d_set_timing(self._h, 1 if enable else 0), "m3d_set_timing")

    def reset_counters(self) -> None:
        _check(self._lib.m3d_reset_counters(self._h), "m3d_reset_counters")

    def kernel_times_ms(self) -> dict:
        """Accumulated device milliseconds per kernel family (needs ``set_timing(True)``)."""
        out = {}
        i = 0
        while True:
            name = self._lib.m3d_kernel_name(i)
            if name is None:
                break
            t = float(self._lib.m3d_kernel_time_ms(self._h, i))
            if t:
                out[name.decode()] = t
            i += 1
        return out

    def launches_by_kernel(self) -> dict:
        out = {}
        i = 0
        while True:
            name = self._lib.m3d_kernel_name(i)
            if name is None:
                break
            n = int(self._lib.m3d_kernel_launches(self._h, i))
            if n:
                out[name.decode()] = n
            i += 1
        return out
