"""Duck-typed stand-in for ``qi2labDataStore`` covering the surface ``PixelDecoder`` calls.

The decoder never imports this module: any object with the same attributes/methods works
(the real ``qi2labDataStore`` included -- SURVEY.md section 8b lists the surface).  The store that reads
the reference's own ``<image>.ome.zarr`` chunk files straight into HBM is ``zarr_store.Qi2labZarrDataStore``
(it inherits the output side from this class); this array-fed one serves tests, benchmarks and callers that
already hold the tile in memory.  It keeps
the reference's directory layout and file names for everything the decode stage WRITES
(``docs/datastore.md:211-300`` of the reference):

    <root>/calibrations/attributes.json      normalisation vectors + metadata + codebook
    <root>/decoded/<tile>_decoded_features.parquet
    <root>/decoded/temporary/iteration_XXX/tileNNN_temp_decoded.parquet
    <root>/all_tiles_filtered_decoded_features/decoded_features.{parquet,csv.gz}

and, for what it READS, stores each image as ``.npy`` next to where the reference keeps the
``.ome.zarr`` group (``readouts/<tile>/<bit>/corrected_data.npy`` uint16 and
``feature_predictor_data.npy`` float32) so worker processes can re-open the store by path
exactly like the reference's workers do (PD:249, PD:354).
"""

from __future__ import annotations

import json
import os
import re
import uuid
from collections.abc import Mapping
from pathlib import Path
from typing import Any, Sequence

import numpy as np
import pandas as pd


class _Ready:
    """Minimal future: the reference calls ``.result()`` on loader return values (PD:1872)."""

    def __init__(self, value):
        self._value = value

    def result(self):
        return self._value


class ZWindowVolume:
    """A (z, y, x) image of which this process holds only planes [z0, z0 + len(block)): what a rank of a z-slab
    sharded decode needs of a volume too large for one host / one GPU (configs[4]: 16 x 400 x 4096 x 4096).
    Quacks like the array a loader returns (``shape``, ``dtype``, ``ndim``, ``[a:b]``); slicing outside the held
    window raises."""

    def __init__(self, full_shape, z0: int, block: np.ndarray):
        self.shape = tuple(int(v) for v in full_shape)
        self.dtype = block.dtype
        self.ndim = 3
        self._z0 = int(z0)
        self._block = block
        if block.shape[1:] != self.shape[1:] or self._z0 < 0 or self._z0 + block.shape[0] > self.shape[0]:
            raise ValueError("block does not fit the volume")

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        if not isinstance(key, slice):
            raise TypeError("ZWindowVolume supports z slices only")
        a, b, step = key.indices(self.shape[0])
        if step != 1:
            raise ValueError("ZWindowVolume supports unit-step z slices only")
        if b <= a:
            return self._block[0:0]
        if a < self._z0 or b > self._z0 + self._block.shape[0]:
            raise IndexError(f"planes [{a}, {b}) are not held by this process (window starts at {self._z0})")
        return self._block[a - self._z0 : b - self._z0]


class ArrayDataStore:
    """In-memory / ``.npy``-backed datastore with the qi2labDataStore decode-stage surface."""

    def __init__(
        self,
        datastore_path: str | Path,
        codebook: pd.DataFrame | None = None,
        voxel_size_zyx_um: Sequence[float] = (0.315, 0.098, 0.098),
        microscope_type: str = "3D",
        validate: bool = False,
    ) -> None:
        self._datastore_path = Path(datastore_path)
        self._calibrations_zarr_path = self._datastore_path / "calibrations"
        self._decoded_root_path = self._datastore_path / "decoded"
        self._readouts_root_path = self._datastore_path / "readouts"
        for p in (self._datastore_path, self._calibrations_zarr_path, self._decoded_root_path):
            p.mkdir(parents=True, exist_ok=True)
        self._mem_readout: dict[tuple[str, str], np.ndarray] = {}
        self._mem_predictor: dict[tuple[str, str], np.ndarray | None] = {}
        attrs = self._load_calibrations_attributes()
        if codebook is not None:
            attrs["codebook"] = {
                "columns": [str(c) for c in codebook.columns],
                "rows": codebook.to_numpy(dtype=object).tolist(),
            }
            attrs["voxel_size_zyx_um"] = [float(v) for v in voxel_size_zyx_um]
            attrs["microscope_type"] = str(microscope_type)
            attrs.setdefault("tile_ids", [])
            attrs.setdefault("bit_ids", [f"bit{i:03d}" for i in range(1, codebook.shape[1])])
            attrs.setdefault("tile_meta", {})
            self._save_calibrations_attributes(attrs)
        elif "codebook" not in attrs:
            raise ValueError(f"{self._datastore_path} holds no codebook; pass one to create it")
        self._refresh(attrs)

    # ------------------------------------------------------------------ JSON sidecar
    def _calibrations_attributes_path(self) -> Path:
        return self._calibrations_zarr_path / "attributes.json"

    def _load_calibrations_attributes(self) -> dict[str, Any]:
        p = self._calibrations_attributes_path()
        if not p.exists():
            return {}
        with open(p) as f:
            d = json.load(f)
        if not isinstance(d, dict):
            raise ValueError("calibrations/attributes.json is invalid.")
        return d

    def _save_calibrations_attributes(self, attributes: Mapping[str, Any]) -> None:
        def clean(v):
            if isinstance(v, np.ndarray):
                return v.tolist()
            if isinstance(v, (np.floating, np.integer)):
                return v.item()
            if isinstance(v, Mapping):
                return {str(k): clean(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return [clean(x) for x in v]
            return v

        p = self._calibrations_attributes_path()
        # one temporary file per writer (ranks / worker threads may save at the same time), then an atomic replace
        tmp = p.with_name(f"{p.name}.{os.getpid()}.{uuid.uuid4().hex}.tmp")
        try:
            with open(tmp, "w") as f:
                json.dump(clean(dict(attributes)), f)
            tmp.replace(p)
        finally:
            if tmp.exists():
                tmp.unlink()

    def _refresh(self, attrs=None) -> None:
        attrs = self._load_calibrations_attributes() if attrs is None else attrs
        cb = attrs["codebook"]
        self._codebook = pd.DataFrame(cb["rows"], columns=cb["columns"])
        self._tile_ids = list(attrs.get("tile_ids", []))
        self._bit_ids = list(attrs.get("bit_ids", []))
        self._voxel_size_zyx_um = np.asarray(attrs["voxel_size_zyx_um"], dtype=float)
        self._microscope_type = attrs.get("microscope_type", "3D")
        self._tile_meta = dict(attrs.get("tile_meta", {}))

    # ------------------------------------------------------------------ attributes
    @property
    def microscope_type(self) -> str:
        return self._microscope_type

    @property
    def codebook(self) -> pd.DataFrame:
        return self._codebook.copy()

    @property
    def tile_ids(self):
        return list(self._tile_ids)

    @property
    def bit_ids(self):
        return list(self._bit_ids)

    @property
    def round_ids(self):
        n = 1
        for meta in self._tile_meta.values():
            n = max(n, max(meta.get("bit_round", [1]) or [1]))
        return [f"round{i:03d}" for i in range(1, n + 1)]

    @property
    def voxel_size_zyx_um(self):
        return self._voxel_size_zyx_um

    @property
    def _num_tiles(self) -> int:
        return len(self._tile_ids)

    def _vector(self, key):
        v = self._load_calibrations_attributes().get(key)
        return None if v is None else np.asarray(v, dtype=np.float32)

    def _set_vector(self, key, value):
        a = self._load_calibrations_attributes()
        a[key] = np.asarray(value, dtype=np.float32)
        self._save_calibrations_attributes(a)

    global_normalization_vector = property(
        lambda s: s._vector("global_normalization_vector"),
        lambda s, v: s._set_vector("global_normalization_vector", v),
    )
    global_background_vector = property(
        lambda s: s._vector("global_background_vector"),
        lambda s, v: s._set_vector("global_background_vector", v),
    )
    iterative_normalization_vector = property(
        lambda s: s._vector("iterative_normalization_vector"),
        lambda s, v: s._set_vector("iterative_normalization_vector", v),
    )
    iterative_background_vector = property(
        lambda s: s._vector("iterative_background_vector"),
        lambda s, v: s._set_vector("iterative_background_vector", v),
    )

    # ------------------------------------------------------------------ tiles
    def _tile_id(self, tile: int | str) -> str | None:
        if isinstance(tile, (int, np.integer)):
            if tile < 0 or tile >= self._num_tiles:
                print("Set tile index >=0 and <" + str(self._num_tiles))
                return None
            return self._tile_ids[int(tile)]
        if isinstance(tile, str):
            if tile not in self._tile_ids:
                print("set valid tiled id")
                return None
            return tile
        print("'tile' must be integer index or string identifier")
        return None

    def _bit_id(self, bit: int | str) -> str | None:
        if isinstance(bit, (int, np.integer)):
            return self._bit_ids[int(bit)] if 0 <= bit < len(self._bit_ids) else None
        return bit if bit in self._bit_ids else None

    def add_tile(
        self,
        readouts: np.ndarray,
        predictors: np.ndarray | None = None,
        tile_id: str | None = None,
        stage_origin_zyx_um: Sequence[float] | None = None,
        camera_to_stage_affine: np.ndarray | None = None,
        global_xform: tuple | None = None,
        wavelengths_um: Sequence[tuple[float, float]] | None = None,
        persist: bool = False,
        bit_round: Sequence[int] | None = None,
        round_transforms_zyx_um: Mapping[int, np.ndarray] | None = None,
        sofima_flow_fields: Mapping[int, tuple] | None = None,
    ) -> str:
        """Register one tile: ``readouts`` (bits, z, y, x) uint16, ``predictors`` float32/None.  ``readouts`` may
        also be a list of per-bit array-likes (``shape``, ``dtype``, z slicing) that are kept as they are -- e.g.
        :class:`ZWindowVolume` for a volume of which this process holds only some planes."""
        lazy = isinstance(readouts, (list, tuple)) and all(
            not isinstance(r, np.ndarray) and len(getattr(r, "shape", ())) == 3 for r in readouts)
        if lazy:
            if predictors is not None or persist:
                raise ValueError("lazy per-bit images take no predictors and cannot be persisted")
            readouts = list(readouts)
        else:
            readouts = np.asarray(readouts)
            if readouts.ndim != 4:
                raise ValueError("readouts must be (bits, z, y, x)")
        if tile_id is None:
            tile_id = f"tile{len(self._tile_ids):04d}"
        attrs = self._load_calibrations_attributes()
        n_bits = len(readouts)
        if len(attrs.get("bit_ids", [])) < n_bits:
            attrs["bit_ids"] = [f"bit{i:03d}" for i in range(1, n_bits + 1)]
        if tile_id not in attrs["tile_ids"]:
            attrs["tile_ids"].append(tile_id)
        meta = {}
        if stage_origin_zyx_um is not None:
            meta["stage_origin_zyx_um"] = [float(v) for v in stage_origin_zyx_um]
            meta["camera_to_stage_affine"] = np.asarray(
                np.eye(4) if camera_to_stage_affine is None else camera_to_stage_affine, dtype=float
            ).tolist()
        if global_xform is not None:
            aff, org, spc = global_xform
            meta["global_affine"] = np.asarray(aff, dtype=float).tolist()
            meta["global_origin"] = [float(v) for v in org]
            meta["global_spacing"] = [float(v) for v in spc]
        if wavelengths_um is not None:
            meta["wavelengths_um"] = [[float(a), float(b)] for a, b in wavelengths_um]
        if bit_round is not None:  # 1-based imaging round of every bit (round 1 = reference frame)
            meta["bit_round"] = [int(r) for r in bit_round]
            meta["round_transforms_zyx_um"] = {
                str(int(k)): np.asarray(v, dtype=float).tolist() for k, v in (round_transforms_zyx_um or {}).items()
            }
        if sofima_flow_fields:  # {round (1-based): (flow (3, fz, fy, fx) float32 XYZ channels, attrs dict)}, DS:4282-4330
            self._mem_flow = getattr(self, "_mem_flow", {})
            meta["sofima_flow_attrs"] = {}
            for rnd, (field, fattrs) in sofima_flow_fields.items():
                self._mem_flow[(tile_id, int(rnd))] = np.ascontiguousarray(field, dtype=np.float32)
                meta["sofima_flow_attrs"][str(int(rnd))] = {
                    k: (np.asarray(v).tolist() if not isinstance(v, str) else v) for k, v in dict(fattrs).items()
                }
                if persist:
                    d = self._datastore_path / "fiducial" / tile_id / f"round{int(rnd):03d}"
                    d.mkdir(parents=True, exist_ok=True)
                    np.save(d / "local_sofima_flow_field.npy", self._mem_flow[(tile_id, int(rnd))])
        attrs.setdefault("tile_meta", {})[tile_id] = meta
        self._save_calibrations_attributes(attrs)
        self._refresh(attrs)
        for b in range(n_bits):
            bit_id = self._bit_ids[b]
            r = readouts[b] if lazy else np.ascontiguousarray(readouts[b], dtype=np.uint16)
            p = None if predictors is None else np.ascontiguousarray(predictors[b], dtype=np.float32)
            if persist:
                d = self._readouts_root_path / tile_id / bit_id
                d.mkdir(parents=True, exist_ok=True)
                np.save(d / "corrected_data.npy", r)
                if p is not None:
                    np.save(d / "feature_predictor_data.npy", p)
            else:
                self._mem_readout[(tile_id, bit_id)] = r
                self._mem_predictor[(tile_id, bit_id)] = p
        return tile_id

    # ------------------------------------------------------------------ image loaders
    def load_local_readout_image(self, tile, bit, return_future: bool = True):
        """uint16 (z, y, x) readout image (DS:4709)."""
        tile_id, bit_id = self._tile_id(tile), self._bit_id(bit)
        if tile_id is None or bit_id is None:
            return None
        arr = self._mem_readout.get((tile_id, bit_id))
        if arr is None:
            p = self._readouts_root_path / tile_id / bit_id / "corrected_data.npy"
            if not p.exists():
                print("Readout image not found.")
                return None
            arr = np.load(p, mmap_mode="c")  # copy-on-write: read-only on disk, writable view for torch
        return _Ready(arr) if return_future else arr

    def load_local_feature_predictor_image(self, tile, bit, return_future: bool = True):
        """float32 (z, y, x) predictor weight image (DS:4838); ones when none was stored."""
        tile_id, bit_id = self._tile_id(tile), self._bit_id(bit)
        if tile_id is None or bit_id is None:
            return None
        if (tile_id, bit_id) in self._mem_readout:
            arr = self._mem_predictor.get((tile_id, bit_id))
        else:
            p = self._readouts_root_path / tile_id / bit_id / "feature_predictor_data.npy"
            arr = np.load(p, mmap_mode="c") if p.exists() else None
        if arr is None:
            arr = UnitPredictor(self.load_local_readout_image(tile, bit, False).shape)
        return _Ready(arr) if return_future else arr

    def load_local_wavelengths_um(self, tile, bit):
        tile_id = self._tile_id(tile)
        meta = self._tile_meta.get(tile_id, {})
        wl = meta.get("wavelengths_um")
        if wl is None:
            return (0.561, 0.580)
        bit_id = self._bit_id(bit)
        return tuple(wl[self._bit_ids.index(bit_id)])

    def load_local_stage_position_zyx_um(self, tile, round=0):
        meta = self._tile_meta.get(self._tile_id(tile), {})
        if "stage_origin_zyx_um" not in meta:
            return None
        return (
            np.asarray(meta["stage_origin_zyx_um"], dtype=np.float32),
            np.asarray(meta["camera_to_stage_affine"], dtype=np.float32),
        )

    def load_global_coord_xforms_um(self, tile):
        meta = self._tile_meta.get(self._tile_id(tile), {})
        if "global_affine" not in meta:
            return None, None, None
        return (
            np.asarray(meta["global_affine"], dtype=np.float32),
            np.asarray(meta["global_origin"], dtype=np.float32),
            np.asarray(meta["global_spacing"], dtype=np.float32),
        )

    # decode-time warping inputs (utils/decode_warping.py:39-52): imaging round of each bit and the
    # physical transform from the reference round into that round; identity unless add_tile got them
    def load_local_round_linker(self, tile, bit):
        meta = self._tile_meta.get(self._tile_id(tile), {})
        br = meta.get("bit_round")
        return 1 if br is None else int(br[self._bit_ids.index(self._bit_id(bit))])

    def load_local_round_transform_zyx_um(self, tile, round):
        meta = self._tile_meta.get(self._tile_id(tile), {})
        idx = round if isinstance(round, (int, np.integer)) else int(str(round)[-3:])
        xf = meta.get("round_transforms_zyx_um", {}).get(str(int(idx)))
        return np.eye(4, dtype=np.float32) if xf is None else np.asarray(xf, dtype=np.float32)

    def load_chromatic_affine_transform_zyx_um(self, *a, **k):
        return np.eye(4, dtype=np.float32)

    def load_local_sofima_flow_field(self, tile, round, return_future: bool | None = True):
        """(flow field (3, fz, fy, fx) float32 with X, Y, Z channels, attributes) or None (DS:4203-4280)."""
        tile_id = self._tile_id(tile)
        idx = round if isinstance(round, (int, np.integer)) else int(str(round)[-3:])
        fattrs = self._tile_meta.get(tile_id, {}).get("sofima_flow_attrs", {}).get(str(int(idx)))
        if fattrs is None:
            return None
        field = getattr(self, "_mem_flow", {}).get((tile_id, int(idx)))
        if field is None:
            p = self._datastore_path / "fiducial" / tile_id / f"round{int(idx):03d}" / "local_sofima_flow_field.npy"
            if not p.exists():
                return None
            field = np.load(p)
        return field, dict(fattrs)

    @property
    def has_identity_decode_transforms(self) -> bool:
        return not any("bit_round" in m or "sofima_flow_attrs" in m for m in self._tile_meta.values())

    # ------------------------------------------------------------------ normalisation vectors
    @staticmethod
    def _validate_decode_run_key(decode_run_key):
        if decode_run_key is None:
            return None
        decode_run_key = str(decode_run_key)
        if not re.fullmatch(r"[A-Za-z0-9_.-]+", decode_run_key):
            raise ValueError("decode_run_key may only contain letters, numbers, '.', '_', and '-'.")
        return decode_run_key

    @staticmethod
    def _vector_keys(kind):
        if kind not in {"global", "iterative"}:
            raise ValueError("kind must be one of 'global' or 'iterative'.")
        return f"{kind}_normalization_vector", f"{kind}_background_vector"

    def load_decode_normalization_vectors(self, decode_run_key, kind):
        nk, bk = self._vector_keys(kind)
        attrs = self._load_calibrations_attributes()
        if decode_run_key is not None:
            key = self._validate_decode_run_key(decode_run_key)
            attrs = attrs.get("decode_normalization_runs", {}).get(key, {})
        n, b = attrs.get(nk), attrs.get(bk)
        if n is None or b is None:
            return None, None
        return np.asarray(n, dtype=np.float32), np.asarray(b, dtype=np.float32)

    def save_decode_normalization_vectors(
        self, decode_run_key, kind, normalization_vector, background_vector, decode_mode=None,
        metadata=None,
    ) -> None:
        nk, bk = self._vector_keys(kind)
        attrs = self._load_calibrations_attributes()
        n = np.asarray(normalization_vector, dtype=np.float32)
        b = np.asarray(background_vector, dtype=np.float32)
        if decode_run_key is None:
            attrs[nk], attrs[bk] = n, b
            if metadata is not None:
                md = dict(attrs.get("decode_normalization_metadata", {}))
                md[kind] = dict(metadata)
                attrs["decode_normalization_metadata"] = md
        else:
            key = self._validate_decode_run_key(decode_run_key)
            runs = dict(attrs.get("decode_normalization_runs", {}))
            run = dict(runs.get(key, {}))
            if decode_mode is not None:
                run["decode_mode"] = str(decode_mode)
            if metadata is not None:
                run[f"{kind}_metadata"] = dict(metadata)
            run[nk], run[bk] = n, b
            runs[key] = run
            attrs["decode_normalization_runs"] = runs
        self._save_calibrations_attributes(attrs)

    def load_decode_normalization_metadata(self, decode_run_key, kind):
        self._vector_keys(kind)
        attrs = self._load_calibrations_attributes()
        if decode_run_key is None:
            md = attrs.get("decode_normalization_metadata", {}).get(kind)
        else:
            key = self._validate_decode_run_key(decode_run_key)
            md = attrs.get("decode_normalization_runs", {}).get(key, {}).get(f"{kind}_metadata")
        return dict(md) if isinstance(md, Mapping) else None

    # ------------------------------------------------------------------ decoded outputs
    def _decoded_run_root(self, decode_run_key=None) -> Path:
        key = self._validate_decode_run_key(decode_run_key)
        return self._decoded_root_path if key is None else self._decoded_root_path / key

    def decoded_temporary_dir(self, decode_run_key=None, iteration=None) -> Path:
        root = self._decoded_run_root(decode_run_key) / "temporary"
        if iteration is not None:
            root = root / f"iteration_{int(iteration):03d}"
        return root

    def _global_filtered_decoded_root(self, decode_run_key=None) -> Path:
        root = self._datastore_path / "all_tiles_filtered_decoded_features"
        key = self._validate_decode_run_key(decode_run_key)
        return root if key is None else root / key

    def save_local_decoded_spots(self, features_df, tile, decode_run_key=None) -> None:
        tile_id = self._tile_id(tile)
        if tile_id is None:
            return None
        root = self._decoded_run_root(decode_run_key)
        root.mkdir(parents=True, exist_ok=True)
        features_df.to_parquet(root / f"{tile_id}_decoded_features.parquet")

    def load_local_decoded_spots(self, tile, decode_run_key=None):
        tile_id = self._tile_id(tile)
        if tile_id is None:
            return None
        p = self._decoded_run_root(decode_run_key) / f"{tile_id}_decoded_features.parquet"
        if not p.exists():
            print("Decoded spots not found.")
            return None
        return pd.read_parquet(p)

    def save_global_filtered_decoded_spots(self, filtered_decoded_df, decode_run_key=None) -> None:
        root = self._global_filtered_decoded_root(decode_run_key)
        root.mkdir(parents=True, exist_ok=True)
        filtered_decoded_df.to_parquet(root / "decoded_features.parquet")
        filtered_decoded_df.to_csv(root / "decoded_features.csv.gz", index=False, compression="gzip")

    def load_global_filtered_decoded_spots(self, decode_run_key=None, gene_ids=None, columns=None):
        p = self._global_filtered_decoded_root(decode_run_key) / "decoded_features.parquet"
        if not p.exists():
            print("Global, filtered, decoded spots not found.")
            return None
        df = pd.read_parquet(p, columns=None if columns is None else list(columns))
        if gene_ids is not None and "gene_id" in df.columns:
            df = df[df["gene_id"].astype(str).isin([str(g) for g in gene_ids])].reset_index(drop=True)
        return df


class UnitPredictor:
    """Marker for 'predictor weight is exactly 1.0 everywhere' (lets the loader skip the
    float32 multiply, which is the identity).  Materialises on ``np.asarray``."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.dtype = np.dtype(np.float32)

    def __array__(self, dtype=None, copy=None):
        return np.ones(self.shape, dtype=np.float32 if dtype is None else dtype)
