"""B200-native ``PixelDecoder``: the reference class' decode stage behind the same API.

Mirrors ``src/merfish3danalysis/PixelDecoder.py`` (0.13.0, "PD") for the hot path only
(SURVEY.md section 8): constructor and method names, argument meaning, return tuples and error
behaviour are the reference's (PD:449-461, PD:4471-4483, PD:4581-4592, PD:4759-4772);
the arithmetic runs in hand-written sm_100a kernels behind the C ABI of
``libm3d_b200.so`` (``include/m3d_b200.h``).  There is no CPU fallback: without a CUDA
device or without the shared library every decode call raises.

What differs from the reference by design
  * the whole tile stays resident in HBM: one H2D of the uint16 stack, one small D2H of the
    feature table (the reference round-trips every z plane, PD:2561-2632);
  * result images (float16 magnitude / distance / scaled) are only materialised for
    ``return_results=True``; otherwise the feature kernel recomputes what it needs;
  * multi-GPU work runs on persistent ranks (``torch.distributed``) or, without a
    launcher, on one host thread per device -- never a process spawn per iteration.

Also built behind this class (SURVEY.md 8f "next" rows): the decode-time warp of unregistered bits (affine and
SOFIMA flow), the pooled table stage (blank-fraction / LR filter, de-duplication, cell assignment) and the
optimiser's per-on-bit weighted centroids.  Out of scope: chromatic-affine ESTIMATION, cell-restricted
normalisation, napari display.
"""

from __future__ import annotations

import hashlib
from dataclasses import dataclass
import shutil
import tempfile
import warnings
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from random import sample
from typing import Literal, Sequence

import numpy as np
import pandas as pd

from . import normalization as _norm
from . import table_stage
from . import zarr_store as _zs
from ._capi import M3D_TABLE_FIXED_COLS, DecodeContext, M3dError

DEFAULT_DECODE_LOWPASS_SIGMA = (3.0, 1.0, 1.0)  # PD:128
DEFAULT_DECODE_MAGNITUDE_THRESHOLD = (1.5, 10.0)  # PD:129
DEFAULT_2D_MINIMUM_PIXELS = 7.0  # PD:130
DEFAULT_3D_MINIMUM_PIXELS = 16.0  # PD:131
MAXIMUM_PIXELS = 500  # PD:2909
_CENTROID_SUFFIXES = ("center_z", "center_y", "center_x", "intensity_sum", "intensity_peak", "voxel_count")


@dataclass(frozen=True)
class ChromaticAffineEstimationConfig:
    """PD:43-68.  Only the two centroid-collection fields are consumed by this build
    (``_add_on_bit_weighted_centroids``); the estimator that uses the rest is outside the hot path."""

    centroid_z_support: int = 7
    centroid_weight_epsilon: float = 1e-6

# feature-table columns written by m3d_features (include/m3d_b200.h)
_COL_FIRST, _COL_AREA, _COL_DEC = 0, 1, 2
_COL_CZ, _COL_CY, _COL_CX = 3, 4, 5
_COL_MU = 6  # zz, yy, xx, zy, zx, yx
_COL_DMIN, _COL_MAGMEAN = 12, 13


_UNSET = object()


def _is_identity_store(datastore) -> bool:
    return bool(getattr(datastore, "has_identity_decode_transforms", False))


class PixelDecoder:
    """Decode one or many tiles of a qi2labdatastore-shaped dataset on B200 GPUs.

    Parameters are the reference's (PD:449-461).  ``use_mask``, ``estimate_chromatic_affines``
    and ``chromatic_affine_config`` are accepted for signature compatibility; the opt-in
    chromatic estimation is outside the hot path and raises ``NotImplementedError`` when
    requested.
    """

    def __init__(
        self,
        datastore,
        merfish_bits: int = 16,
        num_gpus: int = 1,
        verbose: int = 1,
        use_mask: bool | None = False,
        z_range: Sequence[int] | None = None,
        decode_mode: Literal["auto", "2d", "3d"] = "auto",
        estimate_chromatic_affines: bool = False,
        chromatic_affine_config=None,
        excluded_gene_ids: Sequence[str] | None = None,
    ) -> None:
        self._datastore_path = Path(datastore._datastore_path)
        self._datastore = datastore
        self._num_gpus = num_gpus
        self._verbose = verbose
        self._barcodes_filtered = False
        self._n_merfish_bits = merfish_bits
        if decode_mode not in {"auto", "2d", "3d"}:
            raise ValueError("decode_mode must be one of 'auto', '2d', or '3d'.")
        self._decode_mode = decode_mode
        self._estimate_chromatic_affines = bool(estimate_chromatic_affines)
        self._chromatic_affine_config = chromatic_affine_config or ChromaticAffineEstimationConfig()
        if decode_mode == "auto":
            effective = "2d" if self._datastore.microscope_type == "2D" else "3d"
        else:
            effective = decode_mode
        self._effective_decode_mode = effective
        self._is_3D = effective != "2d"
        self._decode_run_key = None
        if z_range is None:
            self._z_crop = False
            self._z_range = [0, None]
        else:
            self._z_crop = True
            self._z_range = [z_range[0], z_range[1]]
        self._z_slice = slice(self._z_range[0], self._z_range[1])

        self._load_codebook()
        self._excluded_gene_ids, self._excluded_codeword_indices = self._resolve_excluded_gene_ids(
            excluded_gene_ids
        )
        self._optimization_excluded_gene_ids = self._excluded_gene_ids
        self._decoding_matrix_no_errors = self._normalize_codebook(include_errors=False)
        self._decoding_matrix = self._decoding_matrix_no_errors.copy()
        self._barcode_count = self._decoding_matrix.shape[0]
        self._bit_count = self._decoding_matrix.shape[1]
        self._mask_image = None
        self._codebook_style = 1
        self._optimize_normalization_weights = False
        self._collect_chromatic_centroids = False
        self._global_normalization_loaded = False
        self._iterative_normalization_loaded = False
        self._global_normalization_vector = None
        self._global_background_vector = None
        self._iterative_normalization_vector = None
        self._iterative_background_vector = None
        self._load_tile_decoding = False
        self._fuse_label_args = None
        self._copy_streams: dict[tuple, object] = {}
        self._prefetch_streams: dict[int, object] = {}
        self._prefetch_pool = None
        self._prefetched = None
        self._buffers: dict[tuple, dict] = {}
        self._slot: dict[int, int] = {}
        self._next_tile_hint = None
        self._contexts: dict[int, DecodeContext] = {}
        self._context_excluded: dict[int, tuple] = {}
        self._device_state: dict[int, dict] = {}
        # decode inputs kept in HBM across the optimiser's iterations (None = caching off)
        self._tile_cache: dict[tuple, tuple] | None = None
        self._tile_cache_budget = 0
        self._tile_cache_used = 0
        self._tile_cache_last_bytes = 0
        self._tile_cache_stats = {"hits": 0, "misses": 0, "resident_tiles": 0, "resident_bytes": 0, "budget_bytes": 0}
        self._profile: dict[str, float] | None = None  # host wall-clock split of the per-tile call (bench / probes)

    # ================================================================== codebook (PD:756-932)
    def _load_codebook(self) -> None:
        """PD:756-800: drop 1-on-bit rows, derive both thresholds from the code geometry."""
        self._df_codebook = self._datastore.codebook.copy()
        self._df_codebook = self._df_codebook.fillna(0)
        bit_columns = self._df_codebook.columns[1 : self._n_merfish_bits + 1]
        on_counts = self._df_codebook.loc[:, bit_columns].to_numpy(dtype=np.int8).sum(axis=1)
        self._df_codebook = self._df_codebook.loc[on_counts != 1].reset_index(drop=True)
        self._codebook_matrix = self._df_codebook.loc[:, bit_columns].to_numpy(dtype=int)
        on = int(np.median(on_counts[on_counts != 1]))
        self._pixel_assignment_threshold = float(
            np.sqrt(2.0 - 2.0 * ((on - 2.0) / np.sqrt(on * (on - 2.0))))
        )
        self._transcript_distance_threshold = float(
            np.sqrt(2.0 - 2.0 * (on / np.sqrt(on * (on + 2.0))))
        )
        self._blank_count = int(
            self._df_codebook.iloc[:, 0].astype("string").str.lower().str.startswith("blank", na=False).sum()
        )
        self._gene_ids = self._df_codebook.iloc[:, 0].tolist()

    def _resolve_excluded_gene_ids(self, excluded_gene_ids):
        """PD:802-841: requested gene IDs -> (resolved ids, full-codebook row indices)."""
        if not excluded_gene_ids:
            return (), ()
        requested, seen = [], set()
        for value in excluded_gene_ids:
            gene_id = str(value).strip()
            if not gene_id or gene_id in seen:
                continue
            requested.append(gene_id)
            seen.add(gene_id)
        index_by_gene: dict[str, list[int]] = {}
        for index, value in enumerate(self._gene_ids):
            index_by_gene.setdefault(str(value), []).append(index)
        unknown = [g for g in requested if g not in index_by_gene]
        if unknown:
            raise ValueError(
                "Optimization exclusion gene IDs are not present in the active "
                f"codebook: {', '.join(unknown)}. Matching is case-sensitive."
            )
        excluded_indices = tuple(i for g in requested for i in index_by_gene[g])
        if len(excluded_indices) >= len(self._gene_ids):
            raise ValueError("Optimization exclusions cannot remove every codeword.")
        resolved_set = set(requested)
        resolved = tuple(str(g) for g in self._gene_ids if str(g) in resolved_set)
        return tuple(dict.fromkeys(resolved)), excluded_indices

    def _codebook_fingerprint(self) -> str:
        """PD:843-853."""
        digest = hashlib.sha256()
        matrix = np.ascontiguousarray(self._codebook_matrix, dtype=np.int8)
        digest.update(np.asarray(matrix.shape, dtype=np.int64).tobytes())
        for gene_id in self._gene_ids:
            encoded = str(gene_id).encode("utf-8")
            digest.update(len(encoded).to_bytes(8, byteorder="little"))
            digest.update(encoded)
        digest.update(matrix.tobytes())
        return digest.hexdigest()

    def _iterative_normalization_metadata(self) -> dict:
        """PD:855-861."""
        return {
            "scope": "iterative_optimization",
            "excluded_gene_ids": list(self._optimization_excluded_gene_ids),
            "codebook_sha256": self._codebook_fingerprint(),
        }

    @staticmethod
    def _suppress_excluded_codeword_assignments(decoded_trace, codebook_index_trace,
                                                excluded_codeword_indices) -> None:
        """PD:863-877 (host arrays; the device path fuses this into the decode kernels).

        Works on NumPy arrays and torch tensors alike: excluded *winners* become background,
        the signal never falls through to another codeword."""
        if not excluded_codeword_indices:
            return
        if isinstance(decoded_trace, np.ndarray):
            ex = np.asarray(excluded_codeword_indices, dtype=codebook_index_trace.dtype)
            decoded_trace[np.isin(codebook_index_trace, ex)] = -1
        else:
            import torch

            ex = torch.as_tensor(list(excluded_codeword_indices), dtype=codebook_index_trace.dtype,
                                 device=codebook_index_trace.device)
            decoded_trace[torch.isin(codebook_index_trace, ex)] = -1

    def _normalize_codebook(self, gpu_id: int = 0, include_errors: bool = False) -> np.ndarray:
        """PD:879-906: rows / ||row||_2 (zero norm -> 1), float64 on the host.  The reference's
        ``include_errors=True`` expansion is dead code in 0.13 (PD:537) and not offered."""
        if include_errors:
            raise NotImplementedError("single-bit-error codebook expansion is not on the 0.13 decode path")
        m = np.asarray(self._codebook_matrix[:, 0 : self._n_merfish_bits])
        mag = np.linalg.norm(m, axis=1, keepdims=True)
        mag[mag == 0] = 1
        return m / mag

    # ================================================================== device contexts
    def _ctx(self, gpu_id: int = 0) -> DecodeContext:
        """One C-ABI context per device; rebuilt when the exclusion set changes."""
        key = tuple(self._excluded_codeword_indices)
        ctx = self._contexts.get(gpu_id)
        if ctx is None or self._context_excluded.get(gpu_id) != key:
            if ctx is not None:
                ctx.close()
            ctx = DecodeContext(self._decoding_matrix.astype(np.float32), key, device=gpu_id)
            self._contexts[gpu_id] = ctx
            self._context_excluded[gpu_id] = key
        return ctx

    def launch_count(self) -> int:
        """Kernels launched so far by this decoder's contexts (bench accounting)."""
        return sum(c.launch_count() for c in self._contexts.values())

    # ================================================================== normalisation state
    def _effective_lowpass_sigma(self, sigma):
        """PD:2026-2035."""
        if sigma is None:
            return None
        sigma_zyx = tuple(float(v) for v in sigma)
        if len(sigma_zyx) != 3:
            raise ValueError("lowpass_sigma must contain three values: z, y, x.")
        return sigma_zyx

    @staticmethod
    def _lowpass_active(sigma) -> bool:
        return sigma is not None and not np.any(np.asarray(sigma, dtype=float) == 0)

    def _default_minimum_pixels(self) -> float:
        """PD:2037-2041."""
        return DEFAULT_3D_MINIMUM_PIXELS if self._is_3D else DEFAULT_2D_MINIMUM_PIXELS

    def _load_global_normalization_vectors(self, gpu_id: int = 0, recalculate: bool = False,
                                           tile_indices=None,
                                           lowpass_sigma=DEFAULT_DECODE_LOWPASS_SIGMA,
                                           collective: bool = False) -> None:
        """PD:934-979."""
        nv, bv = self._datastore.load_decode_normalization_vectors(self._decode_run_key, "global")
        if not recalculate and nv is not None and bv is not None:
            self._global_normalization_vector = np.asarray(nv, dtype=np.float32)
            self._global_background_vector = np.asarray(bv, dtype=np.float32)
            self._global_normalization_loaded = True
        elif collective:
            self._global_normalization_vectors(gpu_id=gpu_id, tile_indices=tile_indices,
                                               lowpass_sigma=lowpass_sigma, collective=True)
        else:
            self._global_normalization_vectors(gpu_id=gpu_id, tile_indices=tile_indices,
                                               lowpass_sigma=lowpass_sigma)

    def _global_normalization_vectors(self, low_percentile_cut: float = 10.0,
                                      high_percentile_cut: float = 90.0,
                                      hot_pixel_threshold: int = 50000, gpu_id: int = 0,
                                      tile_indices=None,
                                      lowpass_sigma=DEFAULT_DECODE_LOWPASS_SIGMA,
                                      collective: bool = False) -> None:
        """PD:981-1199 on the device: per bit, hot-pixel replace -> z-crop -> low-pass ->
        percentile-gated medians through a radix select (no sort, no host copy).

        ``collective`` (every rank of the process group calls this together; the optimiser does): the sampled
        tiles are sharded over the ranks in contiguous chunks, each rank loads / filters only its own, and the
        pooled medians come from the radix select with its 2048-bin digit histograms ALL-REDUCED over NCCL --
        exact, so every rank ends with the vectors a single process computes.  Rank 0 saves them."""
        import torch

        sigma = self._effective_lowpass_sigma(lowpass_sigma)
        rank, world, dist = self._dist()
        collective = bool(collective and dist is not None and world > 1)
        all_ids = list(self._datastore.tile_ids)
        if tile_indices is not None:
            tiles = [all_ids[t] for t in tile_indices]
        elif len(all_ids) > 5:
            tiles = sample(all_ids, 5)
            if collective:  # one sample for the whole group (the reference's is unseeded, PD:1018)
                idx = torch.tensor([all_ids.index(t) for t in tiles], dtype=torch.int64)
                dev_c = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
                idx = idx.to(dev_c)
                dist.broadcast(idx, src=0)
                tiles = [all_ids[int(i)] for i in idx.cpu()]
        else:
            tiles = all_ids
        reduce = None
        if collective:
            tiles = self._contiguous_chunks(tiles, world)[rank]

            def reduce(hist):
                dist.all_reduce(hist)

        ctx = self._ctx(gpu_id)
        dev = ctx.device
        stats = _norm.DeviceOrderStats(ctx)
        bit_ids = list(self._datastore.bit_ids)
        per_bit = []
        # The seed's filtered volumes ARE the decode inputs of the optimiser's first iteration whenever the hot-pixel
        # replacement (PD:1072-1074) changes nothing: same weighting, warp, z crop and low-pass.  With the tile cache on
        # they are written straight into per-tile stacks and handed to the cache, so iteration 0 neither uploads nor
        # filters those tiles again.  A volume that does hold a value above the threshold disqualifies its tile.
        nb = self._n_merfish_bits
        lp_on = self._lowpass_active(sigma)
        keep: dict = {}
        threads_mode = not collective and max(1, min(int(self._num_gpus), torch.cuda.device_count() or 1)) > 1
        same_tiles = tile_indices is not None or len(all_ids) <= 5  # the tiles the iterations will decode on this rank
        if self._tile_cache is not None and lp_on and same_tiles and not threads_mode:
            keep = {t: {"stack": None, "em": [], "ok": True} for t in tiles}
        import threading

        skey = (dev.index, threading.get_ident(), "seed")
        copy = self._copy_streams.get(skey)
        if copy is None:
            copy = self._copy_streams[skey] = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        order = [(bi, bit_id, tile_id) for bi, bit_id in enumerate(bit_ids) for tile_id in tiles]

        def start(item):
            _bi, bit_id_, tile_id_ = item
            readout = self._datastore.load_local_readout_image(tile=tile_id_, bit=bit_id_, return_future=False)
            predictor = self._datastore.load_local_feature_predictor_image(tile=tile_id_, bit=bit_id_, return_future=False)
            # measured A/B on one box (2 tiles of 16 x 64 x 2048 x 2048): seed 0.80 s with the copies serialised, 0.58 s ahead
            return self._upload_volume_ahead(readout, predictor, ctx, copy)

        ahead = start(order[0]) if order else None
        pos = 0
        for bi, bit_id in enumerate(bit_ids):
            vols = []
            for tile_id in tiles:
                r_dev, p_dev, ev = ahead
                pos += 1
                copy.wait_stream(main)  # the next volume's buffers may reuse memory the main stream is still reading
                ahead = start(order[pos]) if pos < len(order) else None
                main.wait_event(ev)
                _ex, em = self._datastore.load_local_wavelengths_um(tile=tile_id, bit=bit_id)
                img = self._finish_weighted_volume(ctx, r_dev, p_dev, self._bit_warp_px(tile_id, bit_id, em))
                del r_dev, p_dev
                kp = keep.get(tile_id) if bi < nb else None
                if kp is not None and kp["ok"] and bool((img > float(hot_pixel_threshold)).any()):
                    kp["ok"], kp["stack"] = False, None  # the replacement below will change this volume
                # PD:1072-1074: hot pixels -> median of the middle plane
                mid = img[img.shape[0] // 2]
                med = stats.median([mid])
                ctx.replace_above(img, float(hot_pixel_threshold), float(med))
                full_z = int(img[self._z_slice].shape[0])
                img = img[self._z_slice].contiguous()
                if lp_on and img.numel():
                    out = None
                    if kp is not None and kp["ok"]:
                        if kp["stack"] is None:
                            need = nb * img.numel() * 4
                            if self._tile_cache_used + self._seed_reserved(keep) + need <= self._tile_cache_resolve_budget():
                                kp["stack"] = torch.empty((nb, *img.shape), dtype=torch.float32, device=dev)
                                kp["full_z"] = full_z
                            else:
                                kp["ok"] = False
                        if kp["ok"]:
                            out = kp["stack"][bi : bi + 1]
                            kp["em"].append(em)
                    img = ctx.lowpass(img[None], sigma, not self._is_3D, out=out)[0]
                vols.append(img)
            per_bit.append(vols)
            # one bit at a time keeps the peak at <= 5 volumes (+ low-pass temporaries)
            nv1, bv1 = _norm.global_normalization_vectors(ctx, [vols], low_percentile_cut, high_percentile_cut,
                                                          reduce=reduce)
            per_bit[-1] = (nv1[0], bv1[0])
            del vols
        torch.cuda.synchronize(dev)
        normalization_vector = np.asarray([p[0] for p in per_bit], dtype=np.float32)
        background_vector = np.asarray([p[1] for p in per_bit], dtype=np.float32)
        if not collective or rank == 0:
            self._datastore.save_decode_normalization_vectors(
                self._decode_run_key, "global", normalization_vector, background_vector,
                decode_mode=self._effective_decode_mode,
            )
        self._global_background_vector = background_vector
        self._global_normalization_vector = normalization_vector
        self._global_normalization_loaded = True
        for tile_id, kp in keep.items():
            if kp["ok"] and kp["stack"] is not None and len(kp["em"]) == nb:
                key = self._stage_key(all_ids.index(tile_id), gpu_id, None, sigma)
                state = {"readout": None, "predictor": None, "stack": kp["stack"], "lowpass_done": True}
                self._tile_cache_insert(key, state, {"em_wvl": kp["em"], "full_z": kp["full_z"]})
                self._tile_cache_stats["seeded_tiles"] = self._tile_cache_stats.get("seeded_tiles", 0) + 1

    @staticmethod
    def _seed_reserved(keep: dict) -> int:
        """bytes of the per-tile stacks the percentile seed already holds for the tile cache"""
        return sum(int(k["stack"].numel()) * 4 for k in keep.values() if k.get("stack") is not None)

    def _load_iterative_normalization_vectors(self, gpu_id: int = 0) -> None:
        """PD:1201-1248 (stale-codebook fingerprint -> ValueError)."""
        load_md = getattr(self._datastore, "load_decode_normalization_metadata", None)
        md = load_md(self._decode_run_key, "iterative") if callable(load_md) else None
        expected = md.get("codebook_sha256") if isinstance(md, dict) else None
        if expected is not None and expected != self._codebook_fingerprint():
            raise ValueError(
                "Cached iterative normalization vectors were fitted with a "
                "different active codebook. Re-run iterative optimization."
            )
        nv, bv = self._datastore.load_decode_normalization_vectors(self._decode_run_key, "iterative")
        if nv is not None and bv is not None:
            bv = np.nan_to_num(bv, 0.0)
            nv = np.nan_to_num(nv, 1.0)
            self._iterative_normalization_vector = np.asarray(nv)
            self._iterative_background_vector = np.asarray(bv)
            self._iterative_normalization_loaded = True
        else:
            self._iterative_normalization_vectors(gpu_id=gpu_id)

    def _iterative_normalization_vectors(self, gpu_id: int = 0, precomputed=_UNSET) -> None:
        """PD:1250-1421: per-bit medians over the pooled transcript table.  ``precomputed``: the
        ``(normalization, background)`` pair (or None = keep the previous vectors) already found by
        ``_pooled_iterative_vectors`` over every rank's rows."""
        if not hasattr(self, "_df_barcodes_loaded"):
            raise ValueError("No decoded transcripts loaded: run optimize_normalization_by_decoding first.")
        if self._iterative_background_vector is None and self._iterative_normalization_vector is None:
            old_b = np.round(np.asarray(self._global_background_vector[0 : self._n_merfish_bits]), 1)
            old_n = np.round(np.asarray(self._global_normalization_vector[0 : self._n_merfish_bits]), 1)
        else:
            old_b = np.asarray(self._iterative_background_vector)
            old_n = np.asarray(self._iterative_normalization_vector)
        if precomputed is _UNSET:
            res = _norm.iterative_normalization_vectors(self._df_barcodes_loaded, self._n_merfish_bits)
        else:
            res = precomputed
        if res is None:
            self._datastore.save_decode_normalization_vectors(
                self._decode_run_key, "iterative", old_n.astype(np.float32), old_b.astype(np.float32),
                decode_mode=self._effective_decode_mode, metadata=self._iterative_normalization_metadata(),
            )
            return
        nv, bv = res
        if self._verbose > 1:
            print("Background delta:", np.round(np.abs(bv - old_b), 1))
            print("Foreground delta:", np.round(np.abs(nv - old_n), 1))
        self._iterative_normalization_vector = nv
        self._iterative_background_vector = bv
        self._datastore.save_decode_normalization_vectors(
            self._decode_run_key, "iterative", nv, bv, decode_mode=self._effective_decode_mode,
            metadata=self._iterative_normalization_metadata(),
        )
        self._iterative_normalization_loaded = True

    def _prepare_normalization_state(self, normalization_method, use_normalization, gpu_id: int = 0,
                                     lowpass_sigma=DEFAULT_DECODE_LOWPASS_SIGMA) -> None:
        """PD:3274-3320."""
        if normalization_method is None:
            normalization_method = "iterative" if use_normalization else "none"
        if normalization_method == "iterative":
            self._load_iterative_normalization_vectors(gpu_id=gpu_id)
        elif normalization_method == "global":
            self._iterative_normalization_loaded = False
            self._load_global_normalization_vectors(gpu_id=gpu_id, lowpass_sigma=lowpass_sigma)
        elif normalization_method == "none":
            self._iterative_normalization_loaded = False
            self._global_normalization_loaded = False
        else:
            raise ValueError(
                "normalization_method must be one of 'iterative', 'global', "
                f"'none', or None. Got {normalization_method!r}."
            )

    def _active_vectors(self):
        """Vector pair ``_decode_pixels`` applies (iterative wins, PD:2577-2592)."""
        if self._iterative_normalization_loaded:
            return self._iterative_background_vector, self._iterative_normalization_vector
        if self._global_normalization_loaded:
            return self._global_background_vector, self._global_normalization_vector
        return None, None

    # ================================================================== tile loading (PD:1828-1946)
    def _bit_warp_px(self, tile, bit_id, emission_wavelength_um):
        """Decode-time warp of one bit as scipy ``affine_transform`` arguments, or None (identity).

        utils/decode_warping.py:184-245 (round transform x inverse chromatic transform) and
        utils/multiview_registration.py:857-870 (physical -> pixel matrix / offset), float32 like
        the reference.  Bits whose round carries a SOFIMA flow field return a ``{"kind": "flow", ...}`` dict
        for ``m3d_warp_flow`` instead."""
        ds = self._datastore
        if _is_identity_store(ds):
            return None
        round_index = ds.load_local_round_linker(tile=tile, bit=bit_id) - 1
        round_id = None
        round_xf = np.eye(4, dtype=np.float32)
        if round_index > 0:
            round_id = ds.round_ids[round_index]
            xf = ds.load_local_round_transform_zyx_um(tile=tile, round=round_id)
            if xf is None:
                raise RuntimeError(f"Missing local round transform for tile={tile} round={round_id}.")
            round_xf = np.asarray(xf, dtype=np.float32)
        chroma = ds.load_chromatic_affine_transform_zyx_um(wavelength_um=emission_wavelength_um)
        chroma = np.eye(4, dtype=np.float32) if chroma is None else np.asarray(chroma, dtype=np.float32)
        transform = np.linalg.inv(chroma) @ round_xf  # decode_warping.py:70-72
        flow = None
        if round_id is not None and hasattr(ds, "load_local_sofima_flow_field"):
            flow = ds.load_local_sofima_flow_field(tile=tile, round=round_id, return_future=False)
        if flow is not None:
            field, attrs = flow
            if not str(attrs.get("sofima_status", "")).startswith("identity_fallback"):
                # decode_warping.py:163-174: affine + SOFIMA flow, sampled once on the device (m3d_warp_flow)
                field = np.asarray(field, dtype=np.float32)
                if field.ndim != 4:
                    raise ValueError("sofima_flow_field_xyz_px must have channel plus ZYX axes.")
                if field.shape[0] != 3 and field.shape[-1] == 3:
                    field = np.moveaxis(field, -1, 0)
                if field.shape[0] != 3:
                    raise ValueError("SOFIMA flow field must have three XYZ channels.")
                return {
                    "kind": "flow", "transform": np.asarray(transform, dtype=np.float32),
                    "flow": np.ascontiguousarray(field),
                    "stride_zyx": np.asarray(attrs["map_stride_zyx_px"], dtype=np.float32),
                    "box_start_zyx": np.asarray(attrs["map_box_start_xyz_px"], dtype=np.float32)[[2, 1, 0]],
                    "reference_shape": tuple(int(v) for v in attrs["reference_shape_zyx_px"]),
                    "spacing": np.asarray(ds.voxel_size_zyx_um, dtype=np.float32),
                }
        if np.allclose(transform, np.eye(4, dtype=np.float32)):  # decode_warping.py:152-157
            return None
        spacing = np.asarray(ds.voxel_size_zyx_um, dtype=np.float32)
        origin = np.zeros(3, dtype=np.float32)
        transform = np.asarray(transform, dtype=np.float32)
        linear_um, translation_um = transform[:3, :3], transform[:3, 3]
        matrix_px = (linear_um * spacing[np.newaxis, :]) / spacing[:, np.newaxis]
        offset_px = (linear_um @ origin + translation_um - origin) / spacing
        return np.asarray(matrix_px, dtype=np.float32), np.asarray(offset_px, dtype=np.float32)

    def _warp_volume(self, ctx, r, p, warp, out_z0=None, out_nz=None, out=None):
        """Apply a decode-time warp (tuple = affine matrix/offset in pixels, dict = affine + SOFIMA flow)."""
        import torch

        if isinstance(warp, dict):
            flow = warp.get("flow_dev")
            if flow is None:
                flow = torch.empty(warp["flow"].shape, dtype=torch.float32, device=ctx.device)
                ctx.upload([(warp["flow"], flow)])
            return ctx.warp_flow(r, warp["transform"], warp["spacing"], flow, warp["stride_zyx"], warp["box_start_zyx"],
                                 warp["reference_shape"], predictor=p, out_z0=0 if out_z0 is None else out_z0,
                                 out_nz=out_nz, out=out)
        if out_z0 is None:
            return ctx.warp_affine(r, warp[0], warp[1], predictor=p)
        return ctx.warp_affine(r, warp[0], warp[1], predictor=p, out_z0=out_z0, out_nz=out_nz, out=out)

    @staticmethod
    def _image_dtype(image) -> np.dtype:
        """dtype of a loader's return value without materialising a lazy store image."""
        dt = getattr(image, "dtype", None)
        return np.dtype(dt) if dt is not None else np.asarray(image).dtype

    @staticmethod
    def _is_unit_predictor(predictor) -> bool:
        return predictor is None or type(predictor).__name__ == "UnitPredictor"

    def _weighted_volume_device(self, readout, predictor, ctx, warp=None):
        """float32(readout) * float32(predictor) for one (z, y, x) volume, on the device, warped
        into the round-1 frame when the bit carries a decode-time transform."""
        import torch

        def to_dev(arr, dtype):
            src = _zs.host_piece(arr, 0, int(arr.shape[0]), dtype)  # lazy store images decode straight to HBM
            dst = torch.empty(tuple(src.shape), dtype=torch.float32 if dtype == np.float32 else torch.uint16,
                              device=ctx.device)
            _zs.transfer(ctx, [(src, dst)])
            return dst

        is_float = self._image_dtype(readout).kind == "f"
        r = to_dev(readout, np.float32 if is_float else np.uint16)
        p = None if self._is_unit_predictor(predictor) else to_dev(predictor, np.float32)
        return self._finish_weighted_volume(ctx, r, p, warp)

    def _finish_weighted_volume(self, ctx, r, p, warp):
        """device readout (uint16 / float32) and predictor (float32 / None) -> float32(readout) * predictor, warped"""
        import torch

        if warp is not None:
            return self._warp_volume(ctx, r, p, warp)
        if r.dtype == torch.float32:
            return r if p is None else r * p
        return ctx.weight(r, p)

    def _upload_volume_ahead(self, readout, predictor, ctx, copy_stream):
        """Start the host -> device copies of one bit volume on ``copy_stream`` and return ``(readout, predictor | None,
        event)``: the percentile seed uploads volume i + 1 while volume i is corrected, filtered and ranked."""
        import torch

        main = torch.cuda.current_stream(ctx.device)
        is_float = self._image_dtype(readout).kind == "f"
        out = []
        with torch.cuda.stream(copy_stream):
            for arr, dtype in ((readout, np.float32 if is_float else np.uint16),
                               (None if self._is_unit_predictor(predictor) else predictor, np.float32)):
                if arr is None:
                    out.append(None)
                    continue
                src = _zs.host_piece(arr, 0, int(arr.shape[0]), dtype)
                dst = torch.empty(tuple(src.shape), dtype=torch.float32 if dtype == np.float32 else torch.uint16,
                                  device=ctx.device)
                dst.record_stream(main)  # consumed on the main stream: the allocator must not recycle it under it
                _zs.transfer(ctx, [(src, dst)])
                out.append(dst)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return out[0], out[1], ev

    def _load_bit_data(self, feature_predictor_threshold: float | None = 0.1, gpu_id: int = 0,
                       z_bounds: tuple[int, int] | None = None, lowpass_sigma=None) -> None:
        """PD:1828-1946: the tile's bit volumes as one device stack + coordinate metadata.  Takes the
        tile staged ahead by ``_schedule_prefetch`` when there is one, else stages it now."""
        import torch

        st = self._device_state.setdefault(gpu_id, {})
        key = self._stage_key(self._tile_idx, gpu_id, z_bounds, lowpass_sigma)
        cache = self._tile_cache
        if cache is not None and key in cache:
            # the optimiser's later iterations: the decode input of this tile is still in HBM
            # (already weighted / warped / low-passed) -- no datastore read, no PCIe, no filter
            new_state, meta = cache[key]
            self._tile_cache_stats["hits"] += 1
            st.clear()
            st.update(new_state)
            self._em_wvl, self._full_z = meta["em_wvl"], meta["full_z"]
            self._load_coordinate_metadata()
            return
        staged = self._take_prefetched(self._tile_idx, gpu_id, z_bounds, lowpass_sigma)
        if staged is None:
            st.clear()  # release the previous tile before staging this one
            # same buffer set as the previous tile: this stream's order already protects it
            staged = self._stage_tile(self._tile_idx, gpu_id, z_bounds, lowpass_sigma,
                                      slot=self._slot.get(gpu_id, 0), fresh_final=self._tile_cache_admit())
        new_state, meta = staged
        ready = meta.get("ready")
        if ready is not None:  # staged on the prefetch stream: order it before this stream's kernels
            torch.cuda.current_stream(self._ctx(gpu_id).device).wait_event(ready)
        if meta.get("slot") is not None:
            self._slot[gpu_id] = meta["slot"]
        st.clear()
        st.update(new_state)
        if cache is not None:
            self._tile_cache_stats["misses"] += 1
            if meta.get("fresh_final"):
                self._tile_cache_insert(key, new_state, meta)
        self._em_wvl, self._full_z = meta["em_wvl"], meta["full_z"]
        self._load_coordinate_metadata()

    # ------------------------------------------------------------------ decode inputs kept in HBM (optimiser)
    @staticmethod
    def _stage_key(tile_idx, gpu_id, z_bounds, lowpass_sigma) -> tuple:
        return (tile_idx, gpu_id, None if z_bounds is None else tuple(z_bounds),
                None if lowpass_sigma is None else tuple(float(v) for v in lowpass_sigma))

    def _tile_cache_begin(self, gpu_id: int, budget_bytes: int | None = None) -> None:
        """Start keeping staged decode inputs in HBM.  The optimiser decodes the SAME tiles in every iteration
        and only the two normalisation vectors change (PD:4689-4751; the reference re-spawns its workers and
        re-loads every tile per iteration, PD:4702-4729), so after the first pass an iteration is
        ``set_normalization -> decode + label -> features -> exchange`` with no PCIe traffic and no re-filtering.

        Budget: ``budget_bytes`` / ``$M3D_TILE_CACHE_GB``, else 55 % of the device memory that is free now -- the
        rest stays for the two staging slots of the tiles that stream through, the labelling scratch and the
        low-pass temporaries.  Tiles are admitted in the order they are first decoded until the budget is used and
        are never evicted: the access pattern is cyclic, so keeping a fixed subset resident is what maximises hits
        (LRU would evict every tile just before its next use).  Tiles beyond the budget keep streaming through the
        double-buffered upload path."""
        import os

        if budget_bytes is None:
            env = os.environ.get("M3D_TILE_CACHE_GB")
            if env is not None:
                budget_bytes = int(float(env) * 1e9)
        self._tile_cache = {}
        # None = resolved from the device's free memory when the first tile is about to be staged
        self._tile_cache_budget = None if budget_bytes is None else max(int(budget_bytes), 0)
        self._tile_cache_gpu = gpu_id
        self._tile_cache_used = 0
        self._tile_cache_last_bytes = 0
        self._tile_cache_stats = {"hits": 0, "misses": 0, "resident_tiles": 0, "resident_bytes": 0,
                                  "budget_bytes": self._tile_cache_budget}

    def _tile_cache_resolve_budget(self) -> int:
        if self._tile_cache_budget is None:
            import torch

            dev = self._ctx(self._tile_cache_gpu).device
            free, _total = torch.cuda.mem_get_info(dev)
            reusable = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
            self._tile_cache_budget = int(0.55 * (free + max(reusable, 0)))
            self._tile_cache_stats["budget_bytes"] = self._tile_cache_budget
        return self._tile_cache_budget

    def _tile_cache_end(self) -> None:
        self._tile_cache = None
        self._tile_cache_used = 0

    def _tile_cache_admit(self) -> bool:
        """Would one more tile (of the size last seen) fit the budget?  Decided before staging so that the tile's
        final buffer is allocated outside the recycled staging slots."""
        if self._tile_cache is None:
            return False
        budget = self._tile_cache_resolve_budget()
        return budget > 0 and self._tile_cache_used + self._tile_cache_last_bytes <= budget

    def _tile_cache_insert(self, key, state: dict, meta: dict) -> None:
        nbytes = sum(int(t.numel()) * int(t.element_size()) for t in state.values() if hasattr(t, "element_size"))
        self._tile_cache_last_bytes = nbytes
        if self._tile_cache_used + nbytes > self._tile_cache_resolve_budget():
            return  # the size guess was too small: this tile keeps streaming
        self._tile_cache[key] = (dict(state), {"em_wvl": meta["em_wvl"], "full_z": meta["full_z"]})
        self._tile_cache_used += nbytes
        self._tile_cache_stats["resident_tiles"] = len(self._tile_cache)
        self._tile_cache_stats["resident_bytes"] = self._tile_cache_used

    def _release_staging_buffers(self) -> None:
        """Drop the two staging slots (kept: the persistent decoded image).  Used once every tile of this rank is
        resident in the cache."""
        self._drop_prefetched()
        for k in [k for k in self._buffers if k[1] != "dec"]:
            del self._buffers[k]

    # ------------------------------------------------------------------ next-tile prefetch
    def _schedule_prefetch(self, tile_idx, gpu_id: int, z_bounds, lowpass_sigma, fresh_final: bool = False) -> None:
        """Stage ``tile_idx`` (datastore reads, host -> device copies, per-bit low-pass) on a side stream
        from a helper thread while the current tile is decoded, annotated and saved.  Multi-tile loops
        (``decode_all_tiles``, the optimiser) are transfer-bound: 13.4 GB over PCIe per tile against a few
        ms of kernels, so hiding everything else behind the next transfer is what is left to win."""
        from concurrent.futures import ThreadPoolExecutor

        import torch

        self._drop_prefetched()
        if self._prefetch_pool is None:
            self._prefetch_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="m3d-prefetch")
        ctx = self._ctx(gpu_id)
        stream = self._prefetch_streams.get(gpu_id)
        if stream is None:
            stream = self._prefetch_streams[gpu_id] = torch.cuda.Stream(device=ctx.device)

        slot = 1 - self._slot.get(gpu_id, 0)  # the buffer set the current tile does NOT occupy
        main = torch.cuda.current_stream(ctx.device)
        idle = torch.cuda.Event()
        idle.record(main)  # everything that touched that buffer set has been enqueued before this point

        def job():
            torch.cuda.set_device(ctx.device)
            with torch.cuda.stream(stream):
                stream.wait_event(idle)
                st, meta = self._stage_tile(tile_idx, gpu_id, z_bounds, lowpass_sigma, slot=slot, alloc_stream=main,
                                            fresh_final=fresh_final)
                ev = torch.cuda.Event()
                ev.record(stream)
                meta["ready"] = ev
            return st, meta

        key = self._stage_key(tile_idx, gpu_id, z_bounds, lowpass_sigma)
        self._prefetched = (key, self._prefetch_pool.submit(job))

    def _take_prefetched(self, tile_idx, gpu_id: int, z_bounds, lowpass_sigma):
        if self._prefetched is None:
            return None
        key, fut = self._prefetched
        self._prefetched = None
        want = self._stage_key(tile_idx, gpu_id, z_bounds, lowpass_sigma)
        try:
            staged = fut.result()
        except Exception:  # the synchronous path below reports the error with its real traceback
            return None
        return staged if key == want else None

    def _drop_prefetched(self) -> None:
        if self._prefetched is not None:
            _key, fut = self._prefetched
            self._prefetched = None
            try:
                fut.result()
            except Exception:
                pass

    def _tile_buffer(self, gpu_id: int, slot, name: str, shape, dtype, device, alloc_stream=None):
        """Persistent per-slot device buffers for whole-tile loads: two slots alternate (the tile being decoded
        / the tile being staged ahead), so multi-tile loops never go back to the allocator and a prefetch
        never shares memory with the tile in flight.  ``slot`` None = a fresh tensor (z-slab path)."""
        import torch

        if slot is None:
            return torch.empty(shape, dtype=dtype, device=device)
        bufs = self._buffers.setdefault((gpu_id, slot), {})
        t = bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            bufs.pop(name, None)
            t = None
            if alloc_stream is not None:  # helper thread: take the memory from the main stream's pool
                with torch.cuda.stream(alloc_stream):
                    t = torch.empty(shape, dtype=dtype, device=device)
            else:
                t = torch.empty(shape, dtype=dtype, device=device)
            bufs[name] = t
        return t

    def _stage_tile(self, tile_idx, gpu_id: int = 0, z_bounds: tuple[int, int] | None = None, lowpass_sigma=None,
                    slot=None, alloc_stream=None, fresh_final: bool = False):
        """PD:1828-1946: gather the tile's bit volumes into one device stack.  Returns
        ``(state dict, {"em_wvl", "full_z"})`` and touches no per-tile attribute of ``self``, so it can
        run ahead on the prefetch thread / stream.

        Registered data (every bit's decode-time transform is the identity): device state
        ``readout`` (bits, z, y, x) uint16 (float32 when the store holds float data) + ``predictor``
        float32 / None (None = weight exactly 1, multiply skipped).
        Otherwise the reference's warp (PD:1882-1889) runs on the device per bit --
        ``m3d_warp_affine`` with the predictor multiply fused -- and the state holds the float32
        ``stack`` directly.  ``z_bounds`` (z-slab sharding) selects planes [a, b) of the z-cropped
        volume; warped bits still read their full input volume.  ``fresh_final``: the buffer(s) the decode
        kernels will read are allocated fresh instead of taken from the recycled staging slot, so that the caller
        can keep them (the optimiser's tile cache); intermediates still use the slot."""
        import torch

        ctx = self._ctx(gpu_id)
        bit_ids = list(self._datastore.bit_ids)[0 : self._n_merfish_bits]

        def final_buffer(name, shape_, dtype_):
            return self._tile_buffer(gpu_id, None if fresh_final else slot, name, shape_, dtype_, ctx.device, alloc_stream)
        loaded = []
        em_wvl = []
        # issue every read first, then collect: a datastore that returns real futures (tensorstore
        # reads in the reference's qi2labDataStore) overlaps the 2 x bits chunk reads instead of
        # serialising them bit by bit as PD:1861-1874 does
        pending = [
            (self._datastore.load_local_readout_image(tile=tile_idx, bit=bit_id),
             self._datastore.load_local_feature_predictor_image(tile=tile_idx, bit=bit_id))
            for bit_id in bit_ids
        ]
        for bit_id, (fr, fp) in zip(bit_ids, pending):
            pa = fp.result() if hasattr(fp, "result") else fp
            ra = fr.result() if hasattr(fr, "result") else fr
            _ex, em = self._datastore.load_local_wavelengths_um(tile=tile_idx, bit=bit_id)
            warp = self._bit_warp_px(tile_idx, bit_id, em)
            loaded.append((ra, None if self._is_unit_predictor(pa) else pa, warp))
            em_wvl.append(em)
        z_full = int(loaded[0][0].shape[0])
        zs, ze, _step = self._z_slice.indices(z_full)
        full_z = max(ze - zs, 0)
        if self._decode_mode == "3d" and full_z < 2:
            raise ValueError("decode_mode='3d' requires at least two z planes after applying z_range.")
        a, b = (zs, ze) if z_bounds is None else (zs + int(z_bounds[0]), zs + int(z_bounds[1]))
        shape = (b - a, *loaded[0][0].shape[1:])
        float_input = any(self._image_dtype(r).kind == "f" for r, _p, _w in loaded)
        npdt = np.float32 if float_input else np.uint16
        st = {}

        # every host -> device copy goes through m3d_upload_batch: straight DMA for page-locked arrays,
        # the library's pinned staging ring for the pageable arrays a datastore normally returns
        if any(w is not None for _r, _p, w in loaded):
            # unregistered tile: every bit is uploaded in its native frame (full z: the warp samples the whole
            # volume) and, as soon as it has arrived, weighted / warped into the round-1 frame -- and low-passed
            # when asked -- on the compute stream while the following bits are still crossing PCIe
            if z_bounds is not None:
                slot = None
            n_b = len(bit_ids)
            native = tuple(int(v) for v in loaded[0][0].shape)
            dt = torch.float32 if float_input else torch.uint16
            raw = self._tile_buffer(gpu_id, slot, "native", (n_b, *native), dt, ctx.device, alloc_stream)
            rawp = None
            if any(pa is not None for _r, pa, _w in loaded):
                rawp = self._tile_buffer(gpu_id, slot, "native_pred", (n_b, *native), torch.float32, ctx.device,
                                         alloc_stream)
            lp_out = None
            if lowpass_sigma is not None:
                stack = self._tile_buffer(gpu_id, slot, "warped", (n_b, *shape), torch.float32, ctx.device, alloc_stream)
                lp_out = final_buffer("lowpassed", (n_b, *shape), torch.float32)
            else:
                stack = final_buffer("warped", (n_b, *shape), torch.float32)
            for _ra, _pa, warp in loaded:  # flow fields go up first: the staging ring is not re-entrant
                if isinstance(warp, dict) and "flow_dev" not in warp:
                    warp["flow_dev"] = torch.empty(warp["flow"].shape, dtype=torch.float32, device=ctx.device)
                    ctx.upload([(warp["flow"], warp["flow_dev"])])
            pieces, piece_bit = [], []
            for i, (ra, pa, _w) in enumerate(loaded):
                pieces.append((_zs.host_piece(ra, 0, native[0], npdt), raw[i]))
                piece_bit.append(i)
                if pa is not None:
                    pieces.append((_zs.host_piece(pa, 0, native[0], np.float32), rawp[i]))
                    piece_bit.append(i)

            def bit_ready(i):
                _ra, pa, warp = loaded[i]
                p_i = None if pa is None else rawp[i]
                if warp is None:
                    r = raw[i, a:b]
                    if r.dtype == torch.float32:
                        stack[i].copy_(r if p_i is None else r * p_i[a:b])
                    else:
                        ctx.weight(r.contiguous(), None if p_i is None else p_i[a:b].contiguous(), out=stack[i])
                else:
                    self._warp_volume(ctx, raw[i], p_i, warp, out_z0=a, out_nz=b - a, out=stack[i])
                if lp_out is not None:
                    ctx.lowpass(stack[i : i + 1], lowpass_sigma, not self._is_3D, out=lp_out[i : i + 1])

            self._upload_pipelined(ctx, pieces, piece_bit, bit_ready)
            st["readout"], st["predictor"] = None, None
            st["stack"] = stack if lp_out is None else lp_out
            if lp_out is not None:
                st["lowpass_done"] = True
            return st, {"em_wvl": em_wvl, "full_z": full_z, "slot": slot, "fresh_final": fresh_final}
        else:
            dt = torch.float32 if float_input else torch.uint16
            if z_bounds is not None:
                slot = None  # z-slabs keep several stacks alive at once: fresh tensors
            full = (len(bit_ids), *shape)
            has_pred = any(pa is not None for _r, pa, _w in loaded)
            if lowpass_sigma is None:  # the uploaded volumes ARE the decode input
                stack = final_buffer("stack", full, dt)
                pred = final_buffer("pred", full, torch.float32) if has_pred else None
            else:
                stack = self._tile_buffer(gpu_id, slot, "stack", full, dt, ctx.device, alloc_stream)
                pred = self._tile_buffer(gpu_id, slot, "pred", full, torch.float32, ctx.device, alloc_stream) \
                    if has_pred else None
            pieces, piece_bit = [], []  # per bit: readout, then its predictor weights (if stored)
            for i, (ra, pa, _w) in enumerate(loaded):
                pieces.append((_zs.host_piece(ra, a, b, npdt), stack[i]))
                piece_bit.append(i)
                if pred is not None:
                    if pa is None:
                        pred[i].fill_(1.0)
                    else:
                        pieces.append((_zs.host_piece(pa, a, b, np.float32), pred[i]))
                        piece_bit.append(i)
            if lowpass_sigma is None:
                _zs.transfer(ctx, pieces)
                st["readout"], st["predictor"] = stack, pred
            else:
                st["readout"], st["predictor"] = None, None
                out = final_buffer("lowpassed", full, torch.float32)
                st["stack"] = self._upload_and_lowpass(ctx, pieces, piece_bit, stack, pred, lowpass_sigma, out)
                st["lowpass_done"] = True
            return st, {"em_wvl": em_wvl, "full_z": full_z, "slot": slot, "fresh_final": fresh_final}

    def _upload_pipelined(self, ctx, pieces, piece_bit, on_bit):
        """Upload ``pieces`` on a copy stream and call ``on_bit(b)`` -- with the compute stream current and already
        ordered behind bit b's copies -- as soon as the last piece of bit b is enqueued, while the pieces of the
        later bits are still being staged / crossing PCIe.  ``piece_bit[i]`` = bit of piece i (adjacent)."""
        import threading

        import torch

        compute = torch.cuda.current_stream(ctx.device)
        skey = (ctx.device.index, threading.get_ident())  # the prefetch thread gets its own copy stream
        copy = self._copy_streams.get(skey)
        if copy is None:
            copy = self._copy_streams[skey] = torch.cuda.Stream(device=ctx.device)
        last_piece_of = {b: i for i, b in enumerate(piece_bit)}  # later pieces overwrite: the bit's last one
        bit_done_at = {i: b for b, i in last_piece_of.items()}

        def piece_arrived(piece):
            b = bit_done_at.get(piece)
            if b is None:
                return
            ev = torch.cuda.Event()
            ev.record(copy)
            compute.wait_event(ev)
            with torch.cuda.stream(compute):
                on_bit(b)

        copy.wait_stream(compute)  # buffers just allocated / filled on the compute stream
        with torch.cuda.stream(copy):
            _zs.transfer(ctx, pieces, on_piece=piece_arrived)

    def _upload_and_lowpass(self, ctx, pieces, piece_bit, stack, pred, sigma, out):
        """Registered tiles with the low-pass on: the per-bit Gaussian (PD:1982-2024) of bit b runs on the
        compute stream as soon as its volume has arrived -- the filter (float64-pipe bound, ~3.7 ms per bit)
        hides behind the transfer of the following bits (~15 ms per bit)."""

        def lowpass_bit(b):
            ctx.lowpass(stack[b : b + 1], sigma, not self._is_3D,
                        predictor=None if pred is None else pred[b : b + 1], out=out[b : b + 1])

        self._upload_pipelined(ctx, pieces, piece_bit, lowpass_bit)
        return out

    def _load_coordinate_metadata(self) -> None:
        """PD:1900-1943."""
        voxel = self._datastore.voxel_size_zyx_um
        self._pixel_size = voxel[1]
        self._axial_step = voxel[0]
        stage_metadata = self._datastore.load_local_stage_position_zyx_um(tile=self._tile_idx, round=0)
        stage_origin = None
        cam = np.eye(4, dtype=np.float32)
        if stage_metadata is not None:
            stage_origin, cam = stage_metadata
            stage_origin = np.asarray(stage_origin, dtype=np.float32)
            cam = np.asarray(cam, dtype=np.float32)
        affine, origin, spacing = self._datastore.load_global_coord_xforms_um(tile=self._tile_idx)
        if affine is None or origin is None or spacing is None:
            affine = np.eye(4)
            if self._is_3D:
                origin = stage_origin if stage_origin is not None else np.zeros(3, dtype=np.float32)
            elif stage_origin is None:
                origin = np.zeros(3, dtype=np.float32)
            elif stage_origin.size == 2:
                origin = np.asarray([0, stage_origin[0], stage_origin[1]], dtype=np.float32)
            else:
                origin = stage_origin
            spacing = self._datastore.voxel_size_zyx_um
        self._affine = np.asarray(affine, dtype=np.float32)
        self._origin = np.asarray(origin, dtype=np.float32)
        self._spacing = np.asarray(spacing, dtype=np.float32)
        self._camera_to_stage_affine = cam

    # ================================================================== kernels (PD:1948-2643)
    def _lp_filter(self, gpu_id: int = 0, sigma=DEFAULT_DECODE_LOWPASS_SIGMA) -> None:
        """PD:1982-2024: Gaussian low-pass of every bit volume (predictor multiply fused in)."""
        st = self._device_state[gpu_id]
        ctx = self._ctx(gpu_id)
        if st.get("stack") is not None:  # warped on load: already float32(readout) * predictor
            st["stack"] = ctx.lowpass(st["stack"], sigma, not self._is_3D)
        else:
            st["stack"] = ctx.lowpass(st["readout"], sigma, not self._is_3D, predictor=st["predictor"])
        self._filter_type = "lp"

    def _prepare_decode_stack(self, gpu_id: int = 0) -> None:
        """Raw path: the decode input is float32(readout) * predictor (PD:1879-1881)."""
        st = self._device_state[gpu_id]
        if st.get("stack") is not None:
            return
        if st["predictor"] is None:
            st["stack"] = st["readout"]  # float32(uint16) is exact; kernels convert on load
        else:
            import torch

            ctx = self._ctx(gpu_id)
            if st["readout"].dtype == torch.float32:
                st["stack"] = st["readout"] * st["predictor"]
            else:
                st["stack"] = ctx.weight(st["readout"], st["predictor"])

    def _decode_pixels(self, magnitude_threshold=(1.1, 2.0), gpu_id: int = 0,
                       materialize_images: bool = False) -> None:
        """PD:2523-2643 as one fused pass: scale, clip, L2-normalise, nearest codeword, pixel
        gate, magnitude gates, exclusions.  ``materialize_images`` also writes the float16
        magnitude / distance / scaled images (the reference always does; here only for
        ``return_results=True`` -- the feature kernel recomputes them otherwise)."""
        import torch

        st = self._device_state[gpu_id]
        ctx = self._ctx(gpu_id)
        self._prepare_decode_stack(gpu_id)
        stack = st["stack"]
        bkg, nrm = self._active_vectors()
        ctx.set_normalization(bkg, nrm)
        ctx.set_thresholds(self._pixel_assignment_threshold, magnitude_threshold[0], magnitude_threshold[1])
        shape = tuple(stack.shape[1:])
        # the decoded image lives in one persistent buffer per GPU: between two production decodes only the
        # previous tile's foreground voxels have to go back to -1 (m3d_decode_label_persistent)
        had = self._buffers.get((gpu_id, "dec"), {}).get("decoded")
        st["decoded"] = self._tile_buffer(gpu_id, "dec", "decoded", shape, torch.int16, ctx.device)
        fresh = had is None or had.data_ptr() != st["decoded"].data_ptr()
        mag = dist = scaled = None
        if materialize_images:
            mag = torch.empty(shape, dtype=torch.float16, device=ctx.device)
            dist = torch.empty(shape, dtype=torch.float16, device=ctx.device)
            scaled = torch.empty(tuple(stack.shape), dtype=torch.float16, device=ctx.device)
        st["magnitude"], st["distance"], st["scaled"] = mag, dist, scaled
        st["n_features"] = None
        if materialize_images or self._fuse_label_args is None:
            ctx.decode(stack, st["decoded"], mag, dist, scaled)
        else:
            # production path: the search kernel hands its foreground list to the labelling stage
            min_px, max_px = self._fuse_label_args
            st["labels"] = None
            if self._wants_chromatic_centroids():
                st["labels"] = torch.empty(shape, dtype=torch.int32, device=ctx.device)
            st["n_features"] = ctx.decode_label(stack, st["decoded"], not self._is_3D, float(min_px), int(max_px),
                                                labels=st["labels"], persistent=not fresh)

    @staticmethod
    def _warp_pixel(pixel_space_point, spacing, origin, affine, camera_to_stage_affine=None):
        """PD:2645-2683: pixel -> physical -> stage -> global."""
        physical = pixel_space_point * spacing + origin
        if camera_to_stage_affine is not None:
            physical = (np.asarray(camera_to_stage_affine) @ np.array([*list(physical), 1]))[:-1]
        return (np.array(affine) @ np.array([*list(physical), 1]))[:-1]

    def _decoded_z_to_source_z(self, decoded_z):
        """PD:2685-2699."""
        return float(self._z_range[0]) + decoded_z

    # ================================================================== features (PD:2908-3201)
    def _extract_barcodes(self, minimum_pixels: float = 3, maximum_pixels: int = MAXIMUM_PIXELS,
                          gpu_id: int = 0) -> None:
        """Connected components + size filters + regionprops on the device, annotation on
        the host over the (small) feature table."""
        st = self._device_state[gpu_id]
        ctx = self._ctx(gpu_id)
        n = st.get("n_features")
        labels = st.get("labels")
        chroma = self._wants_chromatic_centroids()
        if n is None or (chroma and labels is None):
            if chroma:
                import torch

                labels = torch.empty(tuple(st["decoded"].shape), dtype=torch.int32, device=ctx.device)
            n = ctx.label(st["decoded"], not self._is_3D, float(minimum_pixels), int(maximum_pixels), labels=labels)
        table = ctx.features(st["stack"], st["decoded"], self._optimize_normalization_weights, n)
        centroid_stats = None
        if chroma and n > 0:
            centroid_stats = self._chromatic_centroid_statistics(ctx, st["stack"], labels, table)
        # the transcript gate (PD:3176: distance_min <= threshold, float64 compare) is applied on the device, so rows
        # that will be dropped never cross PCIe and the host annotation only touches surviving rows (the optimiser's
        # first iteration has several 1e5 rows per tile)
        if n > 0:
            keep = table[:, _COL_DMIN] <= float(self._transcript_distance_threshold)
            if not bool(keep.all()):
                if centroid_stats is not None:
                    kh = keep.cpu().numpy()
                    centroid_stats = tuple(a[kh] for a in centroid_stats)
                table = table[keep]
        if getattr(self, "_defer_annotation", False) and centroid_stats is None:
            # inside the optimiser (3-D, nothing written to disk): the loop's statistic only reads the per-bit means and
            # the codeword of every surviving row, and it reads them ON THE DEVICE (_pooled_iterative_vectors) -- the
            # table is not copied to the host and no data frame is built (several 1e5 rows per tile in iteration 0)
            import torch

            nb = self._n_merfish_bits
            self._opt_rows = (table[:, M3D_TABLE_FIXED_COLS : M3D_TABLE_FIXED_COLS + nb].to(torch.float32).contiguous(),
                              table[:, _COL_DEC].to(torch.int64))
            if hasattr(self, "_df_barcodes"):
                del self._df_barcodes
            return
        self._opt_rows = None
        eigvals = ctx.inertia_eigvals(table).cpu().numpy() if table.shape[0] > 0 else None
        # column-major on the host: the annotation works column by column (transposed on the device, a view here)
        tab = table.t().contiguous().cpu().numpy().T
        self._df_barcodes = self._annotate_table(tab, centroid_stats, eigvals)

    def _wants_chromatic_centroids(self) -> bool:
        """PD:3117-3125: per-on-bit centroids are collected only inside the optimiser."""
        return bool(self._optimize_normalization_weights and self._collect_chromatic_centroids)

    def _centroid_config(self) -> tuple[int, float]:
        """(centroid_z_support, centroid_weight_epsilon) of ChromaticAffineEstimationConfig (PD:67-68)."""
        cfg = self._chromatic_affine_config
        return (int(getattr(cfg, "centroid_z_support", 7)), float(getattr(cfg, "centroid_weight_epsilon", 1e-6)))

    def _chromatic_centroid_statistics(self, ctx, stack, labels, table):
        """PD:2742-2790 on the device, all bits in one pass (``m3d_centroid_statistics``)."""
        import torch

        z_support = min(self._centroid_config()[0], int(labels.shape[0]))
        if z_support % 2 == 0:
            z_support -= 1
        code = torch.full((table.shape[0] + 1,), -1, dtype=torch.int16, device=ctx.device)
        code[1:] = table[:, _COL_DEC].to(torch.int16)
        sums, peak = ctx.centroid_statistics(labels, stack, z_support, code)
        return sums[1:].cpu().numpy(), peak[1:].cpu().numpy()

    def _table_columns(self) -> list[str]:
        nb = self._n_merfish_bits
        return (
            ["area", "z", "y", "x"]
            + [f"bit{i:02d}_mean_intensity" for i in range(1, nb + 1)]
            + [f"inertia_tensor_eigvals-{k}" for k in range(3)]
            + ["distance_min", "magnitude_mean", "barcode_id", "gene_id", "tile_idx"]
            + [f"on_bit_{k}" for k in range(1, 5)]
            + ([f"bit{b:02d}_{sfx}" for b in range(1, nb + 1) for sfx in _CENTROID_SUFFIXES]
               if self._wants_chromatic_centroids() else [])
            + ["tile_z", "tile_y", "tile_x", "global_z", "global_y", "global_x"]
            + ["signal_mean", "bkd_mean", "s-b_mean"]
        )

    def _inertia_eigvals(self, tab: np.ndarray) -> np.ndarray:
        """scikit-image ``inertia_tensor_eigvals`` from the central second moments of a HOST table
        (the z-slab path assembles its merged rows on the host): uploaded, solved by
        ``m3d_inertia_eigvals``.  The per-tile path hands the device table over directly."""
        import torch

        ctx = self._ctx(self._local_gpu())
        dev_tab = torch.from_numpy(np.ascontiguousarray(tab[:, :M3D_TABLE_FIXED_COLS], dtype=np.float64)).to(ctx.device)
        return ctx.inertia_eigvals(dev_tab).cpu().numpy()

    def _annotate_table(self, tab: np.ndarray, centroid_stats=None, eigvals: np.ndarray | None = None) -> pd.DataFrame:
        """PD:3066-3177 vectorised over the feature rows (row order = canonical id order)."""
        nb = self._n_merfish_bits
        cols = self._table_columns()
        if tab.shape[0] == 0:
            return pd.DataFrame({c: [] for c in cols})
        dec = tab[:, _COL_DEC].astype(np.int32)
        if not (dec >= 0).all():  # components of decoded voxels always carry a codeword; kept for foreign tables
            if centroid_stats is not None:
                centroid_stats = tuple(a[dec >= 0] for a in centroid_stats)
            if eigvals is not None:
                eigvals = eigvals[dec >= 0]
            tab = tab[dec >= 0]
            dec = dec[dec >= 0]
        # columns are assembled as plain arrays and the frame is built once (a per-column
        # DataFrame insert costs ~0.5 ms; the table has 40+ columns)
        col: dict[str, np.ndarray] = {"area": tab[:, _COL_AREA]}
        z = tab[:, _COL_CZ]
        if self._z_crop:
            z = self._decoded_z_to_source_z(z)
        col["z"], col["y"], col["x"] = z, tab[:, _COL_CY], tab[:, _COL_CX]
        bit_means = np.ascontiguousarray(tab[:, M3D_TABLE_FIXED_COLS : M3D_TABLE_FIXED_COLS + nb])
        for i in range(nb):
            col[f"bit{i + 1:02d}_mean_intensity"] = bit_means[:, i]
        if eigvals is not None:
            ev = eigvals
        else:
            ev = self._inertia_eigvals(tab) if tab.shape[0] else np.zeros((0, 3))
        for k in range(3):
            col[f"inertia_tensor_eigvals-{k}"] = ev[:, k]
        col["distance_min"] = tab[:, _COL_DMIN]
        col["magnitude_mean"] = tab[:, _COL_MAGMEAN]
        col["barcode_id"] = dec + 1
        col["gene_id"] = np.asarray(self._gene_ids, dtype=object)[dec]
        col["tile_idx"] = np.full(dec.shape[0], self._tile_idx)
        codebook_bool = self._codebook_matrix.astype(bool, copy=False)
        on0 = np.argsort(~codebook_bool, axis=1)[:, :4].astype(np.int32)  # PD:3101, verbatim
        on_sel = (on0 + 1)[dec]
        for k in range(4):
            col[f"on_bit_{k + 1}"] = on_sel[:, k]
        if self._wants_chromatic_centroids():
            col.update(self._on_bit_centroid_columns(tab, on_sel, centroid_stats))
        col["tile_z"] = np.round(col["z"], 0).astype(int)
        col["tile_y"] = np.round(col["y"], 0).astype(int)
        col["tile_x"] = np.round(col["x"], 0).astype(int)
        pts = np.stack([col["z"], col["y"], col["x"]], axis=1).astype(np.float64)
        # _warp_pixel over all rows at once (PD:3134-3141): p*spacing+origin -> camera_to_stage -> global
        # affine, float64 like the reference's per-row `A @ [p, 1]`
        phys = pts * self._spacing + self._origin
        cam = np.asarray(self._camera_to_stage_affine, dtype=np.float64)
        aff = np.asarray(self._affine, dtype=np.float64)
        hom = np.concatenate([phys, np.ones((phys.shape[0], 1))], axis=1)
        phys = (hom @ cam.T)[:, :3]
        hom = np.concatenate([phys, np.ones((phys.shape[0], 1))], axis=1)
        glob = (hom @ aff.T)[:, :3]
        col["global_z"] = np.round(glob[:, 0], 2)
        col["global_y"] = np.round(glob[:, 1], 2)
        col["global_x"] = np.round(glob[:, 2], 2)
        total = bit_means.sum(axis=1)
        sig = np.take_along_axis(bit_means, on_sel - 1, axis=1).sum(axis=1)
        col["signal_mean"] = sig / 4.0
        col["bkd_mean"] = (total - sig) / float(nb - 4)
        col["s-b_mean"] = col["signal_mean"] - col["bkd_mean"]
        keep = col["distance_min"] <= self._transcript_distance_threshold
        if not keep.all():
            col = {k: v[keep] for k, v in col.items()}
        return pd.DataFrame({c: col[c] for c in cols}, copy=False)  # one block per column, no consolidation copy

    def _on_bit_centroid_columns(self, tab: np.ndarray, on_sel: np.ndarray, centroid_stats) -> dict:
        """PD:2728-2831: sparse per-bit ``center_{z,y,x}``, ``intensity_sum``, ``intensity_peak`` and
        ``voxel_count`` columns (NaN where the bit is off in the row's codeword)."""
        nb = self._n_merfish_bits
        n = tab.shape[0]
        out = {f"bit{b:02d}_{sfx}": np.full(n, np.nan) for b in range(1, nb + 1) for sfx in _CENTROID_SUFFIXES}
        if n == 0 or centroid_stats is None:
            return out
        sums, peak = centroid_stats
        eps = np.float64(np.float32(self._centroid_config()[1]))
        fallback = tab[:, _COL_CZ : _COL_CZ + 3]  # decoded-space centroid (before the z-crop offset)
        area32 = tab[:, _COL_AREA].astype(np.float32).astype(np.float64)
        for b in range(1, nb + 1):
            rows = np.flatnonzero(np.any(on_sel == b, axis=1))
            if rows.size == 0:
                continue
            w = sums[rows, b - 1, 0].copy()
            centers = sums[rows, b - 1, 1:4] / np.maximum(w, eps)[:, None]
            invalid = (~np.all(np.isfinite(centers), axis=1)) | (w <= 0)
            centers[invalid] = fallback[rows][invalid]
            if self._z_crop:
                centers[:, 0] = self._decoded_z_to_source_z(centers[:, 0])
            area = area32[rows].copy()
            pk = peak[rows, b - 1].astype(np.float64)
            missing = (~np.isfinite(w)) | (w <= 0)
            area[missing] = 0.0
            pk[~np.isfinite(pk)] = 0.0
            w[missing] = 0.0
            for k, sfx in enumerate(("center_z", "center_y", "center_x")):
                out[f"bit{b:02d}_{sfx}"][rows] = centers[:, k]
            out[f"bit{b:02d}_intensity_sum"][rows] = w
            out[f"bit{b:02d}_intensity_peak"][rows] = pk
            out[f"bit{b:02d}_voxel_count"][rows] = area
        return out

    # ================================================================== results / persistence
    def _save_barcodes(self) -> None:
        """PD:3203-3233."""
        if self._optimize_normalization_weights:
            # the reference hands tiles from its worker processes to the parent through these files; here
            # the tables travel in memory (all_gather), so they are only written when they are part of the
            # datastore layout (a decode run key -> decoded/temporary/iteration_XXX), not into a throw-away
            # mkdtemp directory
            if self._decode_run_key is None and not getattr(self, "_keep_temp_tables", False):
                return
            d = Path(self._temp_dir)
            d.mkdir(parents=True, exist_ok=True)
            self._df_barcodes.to_parquet(d / ("tile" + str(self._tile_idx).zfill(3) + "_temp_decoded.parquet"))
        elif not self._barcodes_filtered:
            self._datastore.save_local_decoded_spots(self._df_barcodes, tile=self._tile_idx,
                                                     decode_run_key=self._decode_run_key)
        else:
            self._datastore.save_global_filtered_decoded_spots(self._df_filtered_barcodes,
                                                               decode_run_key=self._decode_run_key)

    @property
    def decoded_barcodes(self) -> pd.DataFrame:
        """PD:3235-3247."""
        if not hasattr(self, "_df_barcodes"):
            return pd.DataFrame()
        return self._df_barcodes.copy()

    @property
    def decoded_image(self) -> np.ndarray:
        """PD:3249-3261 (device -> host on access)."""
        for st in self._device_state.values():
            if st.get("decoded") is not None:
                return st["decoded"].cpu().numpy()
        return np.empty((0,), dtype=np.int16)

    def save_decoded_barcodes(self) -> None:
        """PD:3263-3272."""
        self._save_barcodes()

    def _load_all_barcodes(self) -> None:
        """PD:3322-3384 (cell restriction is out of scope: no segmentation in the configs)."""
        if self._optimize_normalization_weights:
            files = sorted(Path(self._temp_dir).glob("*.parquet"), key=lambda p: p.name)
            data = [pd.read_parquet(f) for f in files]
            self._df_barcodes_loaded = pd.concat(data) if data else pd.DataFrame()
        elif self._load_tile_decoding:
            data = [
                self._datastore.load_local_decoded_spots(t, decode_run_key=self._decode_run_key)
                for t in self._datastore.tile_ids
            ]
            self._df_barcodes_loaded = pd.concat([d for d in data if d is not None])
        else:
            self._df_filtered_barcodes = self._datastore.load_global_filtered_decoded_spots(
                decode_run_key=self._decode_run_key
            )
            self._df_barcodes_loaded = self._df_filtered_barcodes.copy()
            self._barcodes_filtered = True
        if self._df_barcodes_loaded.empty:
            if "gene_id" not in self._df_barcodes_loaded.columns:
                self._df_barcodes_loaded["gene_id"] = pd.Series(dtype="string")
            for column in ("distance_min", "magnitude_mean", "area"):
                if column not in self._df_barcodes_loaded.columns:
                    self._df_barcodes_loaded[column] = pd.Series(dtype=np.float32)
        g = self._df_barcodes_loaded["gene_id"]
        self._df_barcodes_loaded = self._df_barcodes_loaded[g.notna() & g.astype(str).str.strip().ne("")]
        if "distance_min" not in self._df_barcodes_loaded.columns:
            raise ValueError(
                "Decoded transcripts are missing 'distance_min'. "
                "Re-decode local transcript parquet files with the exact two-threshold caller."
            )

    def _cleanup(self, keep_pipeline: bool = False) -> None:
        """PD:4426-4469: drop device buffers and per-tile results.  ``keep_pipeline`` (between the tiles of a
        multi-tile loop) keeps the two persistent tile buffers and the tile being staged ahead."""
        if not keep_pipeline:
            self._drop_prefetched()
            self._buffers.clear()
        for st in self._device_state.values():
            st.clear()
        for name in ("_df_barcodes", "_df_filtered_barcodes"):
            if hasattr(self, name):
                delattr(self, name)

    # ================================================================== public API
    def decode_one_tile(
        self,
        tile_idx: int = 0,
        gpu_id: int = 0,
        display_results: bool = False,
        return_results: bool = False,
        lowpass_sigma: Sequence[float] | None = DEFAULT_DECODE_LOWPASS_SIGMA,
        magnitude_threshold: Sequence[float] | None = None,
        minimum_pixels: float | None = None,
        use_normalization: bool | None = True,
        normalization_method: Literal["iterative", "global", "none"] | None = None,
        feature_predictor_threshold: float | None = 0.1,
    ):
        """PD:4471-4579.  With ``return_results=True`` returns the reference's tuple
        ``(image[_lp] float32 (bits,z,y,x), scaled float16, magnitude float16, distance
        float16, decoded int16)`` as host arrays."""
        if display_results:
            raise NotImplementedError("napari display is outside the decode hot path")
        if magnitude_threshold is None:
            magnitude_threshold = DEFAULT_DECODE_MAGNITUDE_THRESHOLD
        if minimum_pixels is None:
            minimum_pixels = self._default_minimum_pixels()
        import time as _time

        prof = self._profile
        t0 = _time.perf_counter()
        self._prepare_normalization_state(normalization_method, use_normalization, gpu_id, lowpass_sigma)
        self._tile_idx = tile_idx
        sigma = self._effective_lowpass_sigma(lowpass_sigma)
        lp_active = self._lowpass_active(sigma)
        t1 = _time.perf_counter()
        self._load_bit_data(feature_predictor_threshold=feature_predictor_threshold, gpu_id=gpu_id,
                            lowpass_sigma=sigma if lp_active else None)
        self._prefetch_upcoming(gpu_id, sigma if lp_active else None)
        if prof is not None:
            if prof.get("sync"):  # attribute the staged copies / filters to "stage", not to the decode that waits for them
                import torch

                # this stream only (it is ordered behind the tile just staged): the next tile's prefetch keeps running
                torch.cuda.current_stream(self._ctx(gpu_id).device).synchronize()
            t2 = _time.perf_counter()
            prof["vectors_s"] = prof.get("vectors_s", 0.0) + (t1 - t0)
            prof["stage_s"] = prof.get("stage_s", 0.0) + (t2 - t1)
        self._filter_type = "raw"
        if lp_active:
            if self._device_state[gpu_id].get("lowpass_done"):
                self._filter_type = "lp"  # filtered bit by bit behind the upload
            else:
                self._lp_filter(sigma=sigma, gpu_id=gpu_id)
        self._fuse_label_args = (minimum_pixels, MAXIMUM_PIXELS)
        try:
            self._decode_pixels(magnitude_threshold=magnitude_threshold, gpu_id=gpu_id,
                                materialize_images=return_results)
        finally:
            self._fuse_label_args = None
        self._extract_barcodes(minimum_pixels=minimum_pixels, gpu_id=gpu_id)
        if prof is not None:
            prof["decode_extract_s"] = prof.get("decode_extract_s", 0.0) + (_time.perf_counter() - t2)
            prof["tiles"] = prof.get("tiles", 0) + 1
        if return_results:
            import torch

            st = self._device_state[gpu_id]
            image = st["stack"]
            if image.dtype != torch.float32:
                image = image.to(torch.float32)
            return (
                image.cpu().numpy(),
                st["scaled"].cpu().numpy(),
                st["magnitude"].cpu().numpy(),
                st["distance"].cpu().numpy(),
                st["decoded"].cpu().numpy(),
            )
        return None

    def _prefetch_upcoming(self, gpu_id: int, sigma) -> None:
        """Multi-tile loops: stage the next tile that is not already resident (tile cache) on the side stream while
        this one is decoded, annotated and saved.  Works for registered and unregistered stores alike: the per-bit
        warp / low-pass kernels of the staged tile run on the prefetch stream, and the staging ring, the copy stream
        and the buffers are per thread / per slot."""
        upcoming = list(getattr(self, "_upcoming_tiles", None) or [])
        nxt, self._next_tile_hint = self._next_tile_hint, None
        if nxt is not None and nxt not in upcoming:
            upcoming.insert(0, nxt)
        self._upcoming_tiles = None
        for t in upcoming:
            key = self._stage_key(t, gpu_id, None, sigma)
            if self._tile_cache is not None and key in self._tile_cache:
                continue  # resident: nothing to stage; look further ahead
            if self._prefetched is not None and self._prefetched[0] == key:
                return  # already on its way
            self._schedule_prefetch(t, gpu_id, None, sigma, fresh_final=self._tile_cache_admit_next())
            return

    def _tile_cache_admit_next(self) -> bool:
        """Admission for a tile staged AHEAD: the tile being decoded now may not have been counted yet."""
        return self._tile_cache_admit()

    # ------------------------------------------------------------------ z-slab sharding of one volume
    def decode_one_tile_sharded(
        self,
        tile_idx: int = 0,
        gpu_id: int | None = None,
        n_slabs: int | None = None,
        lowpass_sigma: Sequence[float] | None = DEFAULT_DECODE_LOWPASS_SIGMA,
        magnitude_threshold: Sequence[float] | None = None,
        minimum_pixels: float | None = None,
        use_normalization: bool | None = True,
        normalization_method: Literal["iterative", "global", "none"] | None = None,
        slabs_per_rank: int = 1,
    ):
        """Decode ONE tile split into z-slabs; result identical to ``decode_one_tile``.

        Under ``torch.distributed`` (world > 1, ``n_slabs`` None) every rank owns a contiguous z range,
        processed as ``slabs_per_rank`` slabs one after another (at most two resident, so a rank can
        own more planes than fit in HBM at once: config 5 on 2 GPUs).  The boundary planes between
        ranks travel by send/recv, the small equivalence / area lists by all_gather, and rank 0 ends up
        with the complete transcript table (``decoded_barcodes``); each rank keeps its part of the
        decoded image.  With ``n_slabs`` given, all slabs are processed one after another on this GPU
        (volumes larger than HBM or than 2^32 voxels).  No reference counterpart: SURVEY.md 8e."""
        from . import sharded as sh

        import torch

        if not self._is_3D:
            raise ValueError("z-slab sharding applies to 3-D decoding; 2-D mode shards by plane (tiles)")
        if magnitude_threshold is None:
            magnitude_threshold = DEFAULT_DECODE_MAGNITUDE_THRESHOLD
        if minimum_pixels is None:
            minimum_pixels = self._default_minimum_pixels()
        rank, world, dist = self._dist()
        distributed = n_slabs is None and dist is not None and world > 1
        if gpu_id is None:
            gpu_id = self._local_gpu()
        self._prepare_normalization_state(normalization_method, use_normalization, gpu_id, lowpass_sigma)
        self._tile_idx = tile_idx
        sigma = self._effective_lowpass_sigma(lowpass_sigma)
        lp = self._lowpass_active(sigma)
        halo = sh.lowpass_z_radius(sigma) if lp else 0
        ctx = self._ctx(gpu_id)
        bkg, nrm = self._active_vectors()
        optimize = self._optimize_normalization_weights
        nb = self._n_merfish_bits

        def load_slab(z0, z1, full_z):
            a, b = max(0, z0 - halo), min(full_z, z1 + halo)
            self._load_bit_data(gpu_id=gpu_id, z_bounds=(a, b))
            st = self._device_state[gpu_id]
            if lp:
                self._lp_filter(sigma=sigma, gpu_id=gpu_id)
                stack = st["stack"][:, z0 - a : z0 - a + (z1 - z0)].contiguous()
            else:
                self._prepare_decode_stack(gpu_id)
                stack = st["stack"]
            st.clear()
            ctx.set_normalization(bkg, nrm)
            ctx.set_thresholds(self._pixel_assignment_threshold, magnitude_threshold[0], magnitude_threshold[1])
            return stack

        # the volume's z extent (after z_range cropping) without loading it
        probe = self._datastore.load_local_readout_image(tile=tile_idx, bit=list(self._datastore.bit_ids)[0])
        probe = probe.result() if hasattr(probe, "result") else probe
        zs_, ze_, _ = self._z_slice.indices(int(probe.shape[0]))
        full_z = max(ze_ - zs_, 0)
        k = max(1, int(slabs_per_rank))
        bounds = sh.split_z(full_z, world * k if distributed else int(n_slabs or 1))
        if distributed:
            # contiguous runs of slabs per rank; ranks without planes (more ranks than planes) own nothing
            per = (len(bounds) + world - 1) // world
            mine = list(range(rank * per, min((rank + 1) * per, len(bounds))))
            owner = {r: r // per for r in range(len(bounds))}
        else:
            mine = list(range(len(bounds)))
            owner = {r: 0 for r in mine}
        import time as _time

        timing = self._zslab_timing = {}
        sync = bool(self._profile and self._profile.get("sync"))
        t_last = [_time.perf_counter()]

        def mark(name):
            """host wall-clock split of the call (with ``_profile = {"sync": True}`` the device is drained at
            every mark so that copies / kernels are attributed to the phase that issued them)"""
            if sync:
                torch.cuda.synchronize(ctx.device)
            now = _time.perf_counter()
            timing[name] = timing.get(name, 0.0) + (now - t_last[0])
            t_last[0] = now

        results: dict[int, sh.SlabResult] = {}
        kept: dict[int, tuple] = {}  # slab -> (stack, decoded, labels) still on the device
        records: dict[tuple, dict] = {}
        decoded_slabs = []
        prev_planes = None

        def crossing_ids(r):
            """slab-local ids of r known to touch an interface (superset of what resolve keeps)."""
            ids = [results[r].pairs[:, 0].astype(np.int64)] if len(results[r].pairs) else []
            if r + 1 in results and len(results[r + 1].pairs):
                ids.append(results[r + 1].pairs[:, 1].astype(np.int64))
            return np.unique(np.concatenate(ids)) if ids else np.zeros(0, dtype=np.int64)

        def retire(r):
            """in-process mode: take the records slab r may have to contribute, then free it."""
            stack_r, decoded_r, labels_r = kept.pop(r)
            for cid, rec in sh.slab_records(ctx, stack_r, decoded_r, labels_r, crossing_ids(r), optimize).items():
                records[(r, cid)] = rec
            decoded_slabs.append(decoded_r)

        for r in mine:
            z0, z1 = bounds[r]
            stack = load_slab(z0, z1, full_z)
            mark("load_h2d_lowpass_s")
            decoded, labels, table = sh.decode_slab(ctx, stack, False, MAXIMUM_PIXELS, optimize)
            mark("decode_label_features_s")
            res = sh.SlabResult(z0, z1, tuple(stack.shape[2:]), table)
            if distributed:
                reqs = []
                if r + 1 < len(bounds) and owner[r + 1] != rank:
                    # one int32 (2, Y, X) message: decoded plane (NCCL has no int16) + label plane
                    out_planes = torch.stack([decoded[-1].to(torch.int32), labels[-1]]).contiguous()
                    reqs.append(dist.isend(out_planes, owner[r + 1]))
                if r > 0 and owner[r - 1] != rank:
                    in_planes = torch.empty((2, *stack.shape[2:]), dtype=torch.int32, device=stack.device)
                    dist.recv(in_planes, owner[r - 1])
                    prev_planes = (in_planes[0].to(torch.int16).contiguous(), in_planes[1].contiguous())
                for q in reqs:
                    q.wait()
                mark("boundary_send_recv_s")
            if prev_planes is not None:
                res.pairs, res.poisoned_here, poisoned_prev = sh.interface(ctx, prev_planes, decoded, labels)
            else:
                poisoned_prev = np.zeros(0, dtype=np.int64)
            mark("interface_pairs_s")
            res._poisoned_prev = poisoned_prev
            results[r] = res
            kept[r] = (stack, decoded, labels)
            prev_planes = (decoded[-1].contiguous(), labels[-1].contiguous())
            del stack
            if (r - 1) in kept:
                retire(r - 1)  # both of its interfaces are known now: at most two slabs stay resident
        if mine and (not distributed or mine[-1] == len(bounds) - 1):
            retire(mine[-1])  # no upper neighbour; otherwise the next rank computes that interface (below)
        mark("resolve_records_s")
        # ---- resolve (identical on every rank)
        summary = {
            r: dict(z0=s.z0, z1=s.z1, shape_yx=s.shape_yx, areas=s.table[:, _COL_AREA].copy(), pairs=s.pairs,
                    poisoned_here=s.poisoned_here, poisoned_prev=s._poisoned_prev)
            for r, s in results.items()
        }
        if distributed:
            # tensor collective (sizes + one padded float64 buffer per rank), not pickled objects
            gathered = sh.all_gather_vectors(dist, sh.pack_summary(summary))
            summary = {k: v for g in gathered for k, v in sh.unpack_summary(g).items()}
        mark("equivalence_gather_s")
        order = sorted(summary)
        areas = [summary[r]["areas"] for r in order]
        pairs = [summary[r]["pairs"] for r in order]
        poisoned = [np.asarray(summary[r]["poisoned_here"], dtype=np.int64) for r in order]
        for i, r in enumerate(order):
            if i > 0 and len(summary[r]["poisoned_prev"]):
                poisoned[i - 1] = np.union1d(poisoned[i - 1], summary[r]["poisoned_prev"])
        keep_local, groups = sh.resolve(areas, pairs, poisoned, minimum_pixels, MAXIMUM_PIXELS)
        # ---- records of the crossing components this rank holds
        need: dict[int, list[int]] = {}
        for g in groups:
            for r, cid in g:
                need.setdefault(order[r], []).append(cid)
        for r in sorted(kept):  # a rank's last slab: its upper interface was resolved on the next rank
            stack, decoded, labels = kept[r]
            for cid, rec in sh.slab_records(ctx, stack, decoded, labels, sorted(need.get(r, [])), optimize).items():
                records[(r, cid)] = rec
            decoded_slabs.append(decoded)
        local_tabs = {r: results[r].table[keep_local[order.index(r)]] for r in mine}
        mark("resolve_records_s")
        if distributed:
            gathered = sh.gather_vectors(dist, sh.pack_records(records, local_tabs, nb), dst=0)
            if rank == 0:
                unpacked = [sh.unpack_records(g) for g in gathered]
                records = {k: v for u in unpacked for k, v in u[0].items()}
                local_tabs = {k: v for u in unpacked for k, v in u[1].items()}
        mark("record_gather_s")
        # ---- assemble on rank 0 (or the only process)
        st = self._device_state.setdefault(gpu_id, {})
        st.clear()
        if decoded_slabs:
            st["decoded"] = torch.cat(decoded_slabs, dim=0) if len(decoded_slabs) > 1 else decoded_slabs[0]
        kept.clear()
        self._slab_bounds = [bounds[r] for r in mine]
        if distributed and rank != 0:
            self._df_barcodes = pd.DataFrame({c: [] for c in self._table_columns()})
            mark("assemble_annotate_s")
            return None
        self._load_coordinate_metadata()
        merged = []
        for g in groups:
            parts = [(summary[order[r]]["z0"], records[(order[r], cid)]) for r, cid in g]
            merged.append(sh.merged_row(parts, summary[order[0]]["shape_yx"], nb, optimize))
        slabs = [sh.SlabResult(summary[r]["z0"], summary[r]["z1"], summary[r]["shape_yx"], local_tabs[r]) for r in order]
        tab = sh.assemble(slabs, [np.ones(len(local_tabs[r]), dtype=bool) for r in order], merged, nb)
        self._df_barcodes = self._annotate_table(tab)
        mark("assemble_annotate_s")
        return None

    # ------------------------------------------------------------------ multi-GPU plumbing
    @staticmethod
    def _dist():
        """(rank, world_size, module) when running under an initialised process group."""
        try:
            import torch.distributed as dist
        except Exception:  # pragma: no cover
            return 0, 1, None
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
        return 0, 1, None

    @staticmethod
    def _contiguous_chunks(items, n_parts):
        """PD:4811-4818: ceil-sized contiguous subsets, one per GPU (empty ones dropped)."""
        chunk = (len(items) + n_parts - 1) // n_parts if n_parts > 0 else len(items)
        out = []
        for g in range(n_parts):
            sub = items[g * chunk : min((g + 1) * chunk, len(items))]
            out.append(list(sub))
        return out

    def _worker_clone(self, gpu_id: int) -> "PixelDecoder":
        """Per-device decoder sharing the datastore and the loaded vectors (the reference
        builds a fresh decoder per worker process, PD:249-268)."""
        w = PixelDecoder(
            self._datastore, merfish_bits=self._n_merfish_bits, num_gpus=1, verbose=0,
            z_range=None if not self._z_crop else self._z_range, decode_mode=self._decode_mode,
            excluded_gene_ids=self._optimization_excluded_gene_ids
            if self._optimize_normalization_weights else self._excluded_gene_ids,
        )
        w._optimize_normalization_weights = self._optimize_normalization_weights
        w._temp_dir = getattr(self, "_temp_dir", None)
        return w

    def _run_tiles(self, tiles, per_tile):
        """Run ``per_tile(decoder, tile_idx, gpu_id)`` for this rank's tiles.

        Under torch.distributed: contiguous chunk of ``tiles`` for this rank, on the current
        device.  Otherwise: one host thread per local GPU, contiguous chunks like PD:4811."""
        import torch

        rank, world, dist = self._dist()
        if dist is not None and world > 1:
            mine = self._contiguous_chunks(list(tiles), world)[rank]
            gpu = self._local_gpu()
            for i, t in enumerate(mine):
                self._upcoming_tiles = mine[i + 1 :]
                per_tile(self, t, gpu)
            return mine
        n = max(1, min(int(self._num_gpus), torch.cuda.device_count() or 1))
        chunks = [c for c in self._contiguous_chunks(list(tiles), n) if c]
        if len(chunks) <= 1:
            only = chunks[0] if chunks else []
            for i, t in enumerate(only):
                self._upcoming_tiles = only[i + 1 :]
                per_tile(self, t, 0)
            return only

        keep = getattr(self, "_persistent_workers", None)  # the optimiser keeps its per-GPU decoders (tile cache)

        def work(args):
            gpu, subset = args
            torch.cuda.set_device(gpu)
            dec = keep.get(gpu) if keep is not None else None
            if dec is None:
                dec = self._worker_clone(gpu)
                if keep is not None:
                    keep[gpu] = dec
                    if self._tile_cache is not None:
                        dec._tile_cache_begin(gpu, self._tile_cache_budget_request)
            dec._global_normalization_vector = self._global_normalization_vector
            dec._global_background_vector = self._global_background_vector
            dec._global_normalization_loaded = self._global_normalization_loaded
            dec._iterative_normalization_vector = None  # re-read from the datastore, like a fresh worker
            dec._iterative_background_vector = None
            dec._iterative_normalization_loaded = False
            for i, t in enumerate(subset):
                dec._upcoming_tiles = subset[i + 1 :]
                per_tile(dec, t, gpu)
            if keep is None:
                dec._cleanup()
            else:
                dec._cleanup(keep_pipeline=True)

        with ThreadPoolExecutor(max_workers=len(chunks)) as ex:
            list(ex.map(work, list(enumerate(chunks))))
        return None

    def _barrier(self):
        _r, world, dist = self._dist()
        if dist is not None and world > 1:
            dist.barrier()

    def _gather_tables(self, local: pd.DataFrame) -> pd.DataFrame:
        """FALLBACK exchange of the optimiser (tables holding values that are not float32 numbers, which this build
        never writes; the regular exchange is ``_pooled_iterative_vectors``): all-gather of every rank's
        transcript rows so that all ranks compute identical per-bit MEDIANS (a sum all-reduce would
        change the result).  Only what the pooled statistics read travels -- per-bit mean intensities,
        on-bits, gene code, tile, distance_min, global coordinates -- as one float64 matrix per rank:
        counts first, then a padded ``all_gather`` of device tensors (NCCL over NVLink; gloo on CPU
        in the tests).  Rows come back in rank order = tile order (contiguous chunks)."""
        _r, world, dist = self._dist()
        if dist is None or world == 1:
            return local
        import torch

        nb = self._n_merfish_bits
        bit_cols = [f"bit{i:02d}_mean_intensity" for i in range(1, nb + 1)]
        on_cols = [f"on_bit_{k}" for k in range(1, 5)]
        extra = ["tile_idx", "distance_min", "global_z", "global_y", "global_x", "area", "magnitude_mean"]
        have = [c for c in extra if c in local.columns] if len(local.columns) else []
        flags = torch.zeros(len(extra), dtype=torch.int64)
        for i, c in enumerate(extra):
            flags[i] = int(c in have)
        n_local = len(local) if "gene_id" in local.columns else 0
        mat = np.zeros((n_local, nb + 4 + 1 + len(extra)), dtype=np.float64)
        if n_local:
            mat[:, :nb] = local[bit_cols].to_numpy(dtype=np.float64)
            mat[:, nb : nb + 4] = local[on_cols].to_numpy(dtype=np.float64)
            mat[:, nb + 4] = pd.Categorical(local["gene_id"].astype(str),
                                            categories=[str(g) for g in self._gene_ids]).codes  # unknown id -> -1
            for i, c in enumerate(extra):
                if c in have:
                    mat[:, nb + 5 + i] = local[c].to_numpy(dtype=np.float64)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        head = torch.cat([torch.tensor([n_local], dtype=torch.int64), flags]).to(dev)
        heads = [torch.empty_like(head) for _ in range(world)]
        dist.all_gather(heads, head)
        heads = [h.cpu() for h in heads]
        counts = [int(h[0]) for h in heads]
        present = torch.stack([h[1:] for h in heads])[[i for i, c in enumerate(counts) if c > 0]]
        n_max = max(counts)
        if n_max == 0:
            return local
        send = torch.zeros((n_max, mat.shape[1]), dtype=torch.float64, device=dev)
        if n_local:
            send[:n_local] = torch.from_numpy(mat).to(dev)
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send)
        pooled = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts) if c > 0], axis=0)
        out = pd.DataFrame(pooled[:, :nb], columns=bit_cols)
        for k, c in enumerate(on_cols):
            out[c] = pooled[:, nb + k].astype(np.int32)
        codes = pooled[:, nb + 4].astype(np.int64)
        names = np.asarray(list(self._gene_ids) + [""], dtype=object)
        out["gene_id"] = names[codes]  # -1 (unknown id) -> ""
        for i, c in enumerate(extra):
            if len(present) and bool(present[:, i].all()):
                out[c] = pooled[:, nb + 5 + i].astype(np.int64) if c == "tile_idx" else pooled[:, nb + 5 + i]
        return out

    def _device_row_queries(self, rows, device, nan_padded: bool = False):
        """The 2 x bits multisets of ``_norm.iterative_vector_queries`` built from the feature tables as they sit on the
        device: ``rows`` = [(per-bit means float32 (n, bits), codeword index int64 (n,)), ...] of this rank's tiles.
        Blank-ness and the four 'on' bits are properties of the codeword (PD:3101: ``argsort(~codebook)[:, :4]``).
        ``nan_padded``: every multiset is a dense (n_rows,) column with NaN where the row does not belong to it (what
        ``m3d_select_hist_batch`` skips) -- two ``where`` instead of 2 x bits boolean selections with a host sync each."""
        import torch

        nb = self._n_merfish_bits
        rows = [r for r in rows if r is not None and r[0].shape[0] > 0]
        if rows:
            vals = torch.cat([r[0] for r in rows]).to(device)
            dec = torch.cat([r[1] for r in rows]).to(device)
        else:
            vals = torch.zeros((0, nb), dtype=torch.float32, device=device)
            dec = torch.zeros((0,), dtype=torch.int64, device=device)
        n_rows = int(vals.shape[0])
        blank = torch.tensor([str(g).lower().startswith("blank") for g in self._gene_ids], dtype=torch.bool, device=device)
        on0 = np.argsort(~self._codebook_matrix.astype(bool, copy=False), axis=1)[:, :4]
        on_mask = np.zeros((len(self._gene_ids), nb + 1), dtype=bool)
        np.put_along_axis(on_mask, np.minimum(on0, nb), True, axis=1)  # an 'on' index beyond the bit columns counts nowhere
        on_mask = torch.from_numpy(on_mask[:, :nb]).to(device)
        keep = ~blank[dec]
        vals, dec = vals[keep], dec[keep]
        on = on_mask[dec]
        if nan_padded:
            nanv = torch.full_like(vals, float("nan"))
            q_on = torch.where(on, vals, nanv).t().contiguous()
            q_off = torch.where(~on, vals, nanv).t().contiguous()
            return [q_on[j] for j in range(nb)] + [q_off[j] for j in range(nb)], int(vals.shape[0]), n_rows
        finite = ~torch.isnan(vals)
        queries = [vals[:, j][on[:, j] & finite[:, j]].contiguous() for j in range(nb)]
        queries += [vals[:, j][~on[:, j] & finite[:, j]].contiguous() for j in range(nb)]
        return queries, int(vals.shape[0]), n_rows

    def _pooled_iterative_vectors(self, local: pd.DataFrame | None, device_rows=None):
        """The optimiser's statistic (PD:1290-1368) over EVERY rank's transcripts without gathering them: each rank
        keeps its rows; the 2 x bits medians (on-bit / off-bit mean intensities of non-blank transcripts) come from
        a radix select whose digit histograms are summed with ``all_reduce`` -- three collectives of a
        (<= 4 x bits, 2048) int64 tensor per iteration, whatever the number of transcripts.  A median is not a sum:
        this is exact (same values as ``np.median`` over the pooled table), so every rank ends with the vectors a
        single process computes.  Collective: every rank calls it once per iteration.

        Returns ``(result, n_pooled)``; result = ``(normalization, background)`` float32 vectors, or None when
        there is no non-blank transcript anywhere (the reference then keeps its previous vectors, PD:1272-1287)."""
        import torch

        rank, world, dist = self._dist()
        nb = self._n_merfish_bits
        multi = dist is not None and world > 1
        backend = getattr(self, "_order_stats_backend", None)  # tests: a host restatement of the histogram kernel
        if backend is None:
            ctx = self._ctx(self._local_gpu())
            device = ctx.device

            def hist_fn(data, row, pm, pv, sh):
                ctx.select_hist(data, row, prefix_mask=pm, prefix_value=pv, shift=sh)

            hist_batch_fn = ctx.select_hist_batch  # a whole level of the 2 x bits medians in one launch
        else:
            device, hist_fn, hist_batch_fn = torch.device("cpu"), backend, None
        queries, kept, foreign = None, 0, 0
        if device_rows is not None:  # the tables never left the device (3-D optimiser iterations)
            queries, kept, n_local = self._device_row_queries(device_rows, device, nan_padded=hist_batch_fn is not None)
        else:
            n_local = len(local)
            has_cols = len(local.columns) > 0 and "gene_id" in local.columns and any(
                c.startswith("bit") and c.endswith("_mean_intensity") for c in local.columns)
            if has_cols:
                queries, kept = _norm.iterative_vector_queries(local, nb, device)
                foreign = int(queries is None)
            if queries is None or len(queries) != 2 * nb:
                foreign = max(foreign, int(queries is not None))  # a table with other bit columns: host path
                queries = [torch.zeros(0, dtype=torch.float32, device=device) for _ in range(2 * nb)]
        flags = torch.tensor([n_local, kept, foreign], dtype=torch.int64, device=device)
        if multi:
            dist.all_reduce(flags)
        n_pooled, n_kept, any_foreign = (int(v) for v in flags.cpu())
        if any_foreign:
            # values that are not float32 numbers (tables this build did not write): gather and take the host median
            pooled = self._gather_tables(local if local is not None else pd.DataFrame())
            if len(pooled) == 0 and "gene_id" not in pooled.columns:
                return None, n_pooled
            return _norm.iterative_normalization_vectors(pooled, nb), n_pooled
        if n_kept == 0:
            return None, n_pooled

        def reduce(hist):
            dist.all_reduce(hist)

        med = _norm.pooled_medians(queries, hist_fn,
                                   lambda rows: torch.zeros((rows, 2048), dtype=torch.int64, device=device),
                                   reduce if multi else None, hist_batch_fn=hist_batch_fn)
        return _norm.finish_iterative_vectors(med[:nb], med[nb:]), n_pooled

    def optimize_normalization_by_decoding(
        self,
        n_random_tiles: int = 5,
        n_iterations: int = 10,
        minimum_pixels: float | None = None,
        feature_predictor_threshold: float | None = 0.1,
        lowpass_sigma: Sequence[float] | None = DEFAULT_DECODE_LOWPASS_SIGMA,
        magnitude_threshold: Sequence[float] | None = None,
        tile_indices: Sequence[int] | None = None,
        estimate_chromatic_affines: bool | None = None,
        excluded_gene_ids: Sequence[str] | None = None,
    ) -> None:
        """PD:4581-4757: global seed, then ``n_iterations`` of decode -> per-bit medians.

        Iteration 0 decodes with the global vectors, later ones with the iterative vectors
        (PD:395-405).  Tables are handed over in memory (and all-gathered across ranks);
        the per-tile parquet files of the reference are still written to the temporary
        directory so the on-disk layout matches."""
        if self._num_gpus < 1:
            raise RuntimeError("No GPUs allocated.")
        if magnitude_threshold is None:
            magnitude_threshold = DEFAULT_DECODE_MAGNITUDE_THRESHOLD
        if minimum_pixels is None:
            minimum_pixels = self._default_minimum_pixels()
        self._optimization_excluded_gene_ids, excluded_idx = self._resolve_excluded_gene_ids(excluded_gene_ids)
        blank_excl = [g for g in self._optimization_excluded_gene_ids if g.lower().startswith("blank")]
        if blank_excl:
            warnings.warn(
                "Blank codewords already do not contribute to iterative "
                "normalization; excluding them only suppresses their temporary "
                "assignments: " + ", ".join(blank_excl),
                stacklevel=2,
            )
        run_chromatic = (
            self._estimate_chromatic_affines if estimate_chromatic_affines is None else bool(estimate_chromatic_affines)
        )
        if run_chromatic:
            raise NotImplementedError("RNA-derived chromatic affine estimation is outside the hot path (SURVEY 2 #9)")
        all_tiles = list(range(len(self._datastore.tile_ids)))
        self._iterative_background_vector = None
        self._iterative_normalization_vector = None
        self._global_background_vector = None
        self._optimize_normalization_weights = True
        saved_excluded = (self._excluded_gene_ids, self._excluded_codeword_indices)
        self._excluded_gene_ids = self._optimization_excluded_gene_ids
        self._excluded_codeword_indices = excluded_idx
        rank, world, dist = self._dist()
        import time as _time

        timing = self._optimizer_timing = {"seed_s": 0.0, "iterations": []}
        made_temp_dir = False
        temp_dir = None
        try:
            t_seed = _time.perf_counter()
            # decode inputs stay in HBM from the seed / the first iteration on (budgeted; see _tile_cache_begin)
            self._tile_cache_budget_request = getattr(self, "tile_cache_budget_bytes", None)
            self._tile_cache_begin(self._local_gpu(), self._tile_cache_budget_request)
            # the percentile seed (PD:4662-4667 computes it in the parent process): under a process group its
            # tiles are sharded over the ranks and the pooled medians are all-reduced (see the method)
            self._load_global_normalization_vectors(gpu_id=self._local_gpu(), recalculate=True,
                                                    tile_indices=tile_indices, lowpass_sigma=lowpass_sigma,
                                                    collective=True)
            self._barrier()  # rank 0 has written calibrations/attributes.json
            timing["seed_s"] = _time.perf_counter() - t_seed
            if self._decode_run_key is None:
                # tables travel in memory (all_gather); the scratch directory exists only for `_keep_temp_tables`
                # debugging and is removed by the rank that made it
                temp_dir = Path(tempfile.mkdtemp())
                made_temp_dir = True
            else:
                temp_dir = self._datastore.decoded_temporary_dir(self._decode_run_key)
                temp_dir.mkdir(parents=True, exist_ok=True)
            self._temp_dir = temp_dir
            self._persistent_workers = {}
            if tile_indices is not None:
                random_tiles = list(tile_indices)
            elif len(all_tiles) > n_random_tiles:
                random_tiles = sample(all_tiles, n_random_tiles)
                if dist is not None and world > 1:
                    box = [random_tiles]
                    dist.broadcast_object_list(box, src=0)
                    random_tiles = box[0]
            else:
                random_tiles = all_tiles
            for iteration in range(n_iterations):
                use_norm = iteration > 0
                tables: list[pd.DataFrame] = []
                t_it = _time.perf_counter()
                prof_was = self._profile
                self._profile = {"sync": bool(prof_was and prof_was.get("sync"))} if prof_was is not None else None
                hits0 = self._tile_cache_stats["hits"]

                # 3-D runs that write no per-iteration files keep the feature tables on the device (see _extract_barcodes)
                defer = bool(self._is_3D and self._decode_run_key is None and not getattr(self, "_keep_temp_tables", False)
                             and not self._wants_chromatic_centroids()
                             and getattr(self, "_order_stats_backend", None) is None)
                dev_rows: list = []

                def per_tile(dec, tile_idx, gpu, _use=use_norm, _defer=defer):
                    dec._defer_annotation = _defer
                    try:
                        dec.decode_one_tile(
                            tile_idx=tile_idx, gpu_id=gpu, lowpass_sigma=lowpass_sigma,
                            magnitude_threshold=magnitude_threshold, minimum_pixels=minimum_pixels,
                            feature_predictor_threshold=feature_predictor_threshold,
                            normalization_method="iterative" if _use else "global",
                        )
                    finally:
                        dec._defer_annotation = False
                    if _defer and getattr(dec, "_opt_rows", None) is not None:
                        dev_rows.append((tile_idx, dec._opt_rows))
                        tables.append((tile_idx, None))
                    else:
                        dec._save_barcodes()
                        tables.append((tile_idx, dec._df_barcodes))

                mine = self._run_tiles(random_tiles, per_tile)
                if mine is not None and self._tile_cache is not None and mine and all(
                        any(k[0] == t for k in self._tile_cache) for t in mine):
                    self._release_staging_buffers()  # every tile of this rank is resident: the two slots can go
                t_tiles = _time.perf_counter()
                # pooled row order = tile order of `random_tiles` on one process and across ranks alike (the 2-D
                # within-tile de-duplication breaks ties by row)
                pos = {t: i for i, t in enumerate(random_tiles)}
                tables.sort(key=lambda t: pos.get(t[0], len(pos)))
                dev_rows.sort(key=lambda t: pos.get(t[0], len(pos)))
                frames = [t[1] for t in tables if t[1] is not None]
                local = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
                # every rank keeps ITS rows (the reference pools all tiles in the parent process, PD:4731-4733): the
                # 2-D within-tile collapse only ever joins rows of one tile, and the medians are pooled by
                # all-reduced histograms (_pooled_iterative_vectors) -- no table travels between ranks
                if len(local) == 0 and "gene_id" not in local.columns:
                    local = pd.DataFrame({"gene_id": pd.Series(dtype="string")})
                g = local["gene_id"]
                keep_rows = g.notna() & g.astype(str).str.strip().ne("")
                self._df_barcodes_loaded = local if bool(keep_rows.all()) else local[keep_rows]
                if not self._is_3D:
                    self._remove_duplicates_within_tile(
                        radius_xy=self._datastore.voxel_size_zyx_um[-1],
                        radius_z=self._datastore.voxel_size_zyx_um[0],
                    )
                t_local = _time.perf_counter()
                use_dev = defer and len(frames) == 0  # local multi-GPU threads or mixed paths: data frames
                vectors, n_pooled = self._pooled_iterative_vectors(
                    None if use_dev else self._df_barcodes_loaded,
                    device_rows=[r[1] for r in dev_rows] if use_dev else None)
                t_gather = _time.perf_counter()
                pooled = range(n_pooled)  # only its length is reported below
                self._load_global_normalization_vectors(gpu_id=self._local_gpu(), lowpass_sigma=lowpass_sigma)
                if rank == 0:
                    self._iterative_normalization_vectors(gpu_id=self._local_gpu(), precomputed=vectors)
                self._barrier()
                if rank != 0:
                    self._iterative_normalization_vector = None
                    self._load_iterative_normalization_vectors(gpu_id=self._local_gpu())
                self._global_background_vector = None
                self._global_normalization_vector = None
                t_end = _time.perf_counter()
                it = {"iteration": iteration, "total_s": t_end - t_it, "tiles_s": t_tiles - t_it,
                      "local_table_s": t_local - t_tiles, "exchange_s": t_gather - t_local, "vectors_s": t_end - t_gather,
                      "tiles_this_rank": len(tables), "cache_hits": self._tile_cache_stats["hits"] - hits0,
                      "transcripts_pooled": int(len(pooled))}
                if self._profile is not None:
                    it.update({k: v for k, v in self._profile.items() if k.endswith("_s")})
                self._profile = prof_was
                timing["iterations"].append(it)
            timing["cache"] = dict(self._tile_cache_stats)
        finally:
            self._excluded_gene_ids, self._excluded_codeword_indices = saved_excluded
            for w in (getattr(self, "_persistent_workers", None) or {}).values():
                w._tile_cache_end()
                w._cleanup()
            self._persistent_workers = None
            self._tile_cache_end()
            self._cleanup()
            self._optimize_normalization_weights = False
            if made_temp_dir and temp_dir is not None:  # every rank removes the directory it made
                shutil.rmtree(temp_dir, ignore_errors=True)

    def _local_gpu(self) -> int:
        import torch

        _r, world, dist = self._dist()
        if dist is not None and world > 1 and torch.cuda.is_available():
            return torch.cuda.current_device()
        return 0

    # ================================================================== post-decode table stage (8f-3)
    def _filter_all_barcodes_blank_fraction(self, target_gross_misid_rate: float = 0.05, intensity_bins=None,
                                            voxel_number_bins=None, vector_distance_bins=None) -> None:
        """PD:3386-3846: blank-fraction histogram filter over (magnitude_mean, area, distance_min);
        binning and histograms on the device (``m3d_table_hist3d``)."""
        self._df_filtered_barcodes, self._blank_fraction_filter_results = table_stage.filter_blank_fraction(
            self._ctx(self._local_gpu()), self._df_barcodes_loaded, self._blank_count, self._barcode_count,
            target_gross_misid_rate, intensity_bins, voxel_number_bins, vector_distance_bins,
        )
        self._barcodes_filtered = True
        if self._verbose > 1:
            print("blank fraction filter diagnostics:")
            for key, value in self._blank_fraction_filter_results.items():
                if isinstance(value, (float, int, str, bool)):
                    print(f"{key}: {value}")
                else:
                    print(f"{key}: {type(value)} with shape {getattr(value, 'shape', 'N/A')}")

    @staticmethod
    def _calculate_lr_fdr(df: pd.DataFrame, threshold: float, blank_count: int, barcode_count: int,
                          verbose: bool = False) -> float:
        """PD:3849-3905."""
        return table_stage.calculate_lr_fdr(df, threshold, blank_count, barcode_count)

    def _filter_all_barcodes_LR(self, lr_fdr_target: float = 0.05) -> None:
        """PD:3907-4058 (scikit-learn on the host, like the reference)."""
        self._df_filtered_barcodes = table_stage.filter_lr(
            self._df_barcodes_loaded, self._is_3D, self._blank_count, self._barcode_count, lr_fdr_target,
            self._verbose,
        )
        self._barcodes_filtered = True

    def _apply_filter_method(self, filter_method, target_gross_misid_rate: float, lr_fdr_target: float) -> None:
        """PD:4914-4950."""
        if filter_method == "blank_fraction":
            self._filter_all_barcodes_blank_fraction(target_gross_misid_rate=float(target_gross_misid_rate))
            return
        if filter_method == "lr":
            self._filter_all_barcodes_LR(lr_fdr_target=float(lr_fdr_target))
            return
        raise ValueError("filter_method must be one of 'blank_fraction' or 'lr'.")

    def _remove_duplicates_in_tile_overlap(self, radius: float = 0.75) -> None:
        """PD:4137-4177: of two detections from different tiles within ``radius`` um (3-D), the one
        with the larger (distance_min, row) is dropped; radius search on the device."""
        self._df_filtered_barcodes, drop = table_stage.remove_duplicates_in_tile_overlap(
            self._ctx(self._local_gpu()), self._df_filtered_barcodes, radius)
        if self._verbose > 1:
            print("Dropped points: " + str(int(drop.sum())))

    def _remove_duplicates_within_tile(self, radius_xy: float = 0.1, radius_z: float = 0.50) -> None:
        """PD:4179-4363 (2-D mode): collapse same-gene detections split across adjacent z planes.
        Same tile, same gene, XY distance <= radius_xy, 0 < |dz| <= radius_z; union-find clusters on
        the device; keep the smallest ``distance_min`` (ties: first row).  Acts on the filtered
        table when there is one, else on the loaded table, like the reference."""
        filtered = getattr(self, "_df_filtered_barcodes", None) is not None
        df = self._df_filtered_barcodes if filtered else getattr(self, "_df_barcodes_loaded", None)
        if df is None or df.empty or len(df) < 2:
            return
        out, drop = table_stage.remove_duplicates_within_tile(self._ctx(self._local_gpu()), df, radius_xy, radius_z)
        if self._verbose > 1:
            print("Dropped points: " + str(int(drop.sum())))
        if filtered:
            self._df_filtered_barcodes = out
        else:
            self._df_barcodes_loaded = out

    def _assign_cells(self) -> None:
        """PD:4076-4135: ``cell_id`` from the Cellpose ImageJ ROIs in global coordinates
        (``segmentation/cellpose/imagej_rois/global_coords_rois.zip``): the archive is parsed by ``roi.py``
        (the reference uses roifile), point-in-polygon runs on the device (``m3d_assign_cells``; the reference
        uses an R-tree + shapely per row).  A missing / unreadable archive is reported and skipped."""
        import zipfile

        from . import roi as _roi

        path = Path(self._datastore._datastore_path) / "segmentation" / "cellpose" / "imagej_rois" / \
            "global_coords_rois.zip"
        try:
            rois = _roi.read_roi_zip(path)
        except (OSError, FileNotFoundError, ValueError, zipfile.BadZipFile) as e:
            print(f"Failed to read ROIs: {e}")
            return
        # roi.subpixel_coordinates[:, ::-1] (PD:4074): (x, y) vertices -> (y, x) polygons; cell ids count
        # only the usable polygons, like the reference's shapely_polygons list
        polygons = [r[:, ::-1] for r in rois if r is not None and len(r) >= 3]
        self._df_filtered_barcodes["cell_id"] = table_stage.assign_cells(
            self._ctx(self._local_gpu()), self._df_filtered_barcodes, polygons)

    def decode_all_tiles(
        self,
        assign_to_cells: bool = True,
        lowpass_sigma: Sequence[float] | None = DEFAULT_DECODE_LOWPASS_SIGMA,
        magnitude_threshold: Sequence[float] | None = None,
        minimum_pixels: float | None = None,
        feature_predictor_threshold: float | None = 0.1,
        normalization_method: Literal["iterative", "global", "none"] = "iterative",
        duplicate_radius_xy: float | None = None,
        duplicate_radius_z: float | None = None,
        filter_method: Literal["blank_fraction", "lr"] = "blank_fraction",
        target_gross_misid_rate: float = 0.05,
        lr_fdr_target: float = 0.05,
    ) -> None:
        """PD:4759-4870, decode stage: every tile -> ``decoded/<tile>_decoded_features.parquet``.

        Tiles are sharded in contiguous chunks over GPUs (ranks or local devices) with no
        data-path collective; then the pooled table stage (PD:4849-4870): blank-fraction / LR
        filter, de-duplication, optional cell assignment, ``all_tiles_filtered_decoded_features``."""
        if self._num_gpus < 1:
            raise RuntimeError("No GPUs allocated.")
        if magnitude_threshold is None:
            magnitude_threshold = DEFAULT_DECODE_MAGNITUDE_THRESHOLD
        if minimum_pixels is None:
            minimum_pixels = self._default_minimum_pixels()
        self._validate_filter_configuration(filter_method, float(target_gross_misid_rate), float(lr_fdr_target))
        all_tiles = list(range(len(self._datastore.tile_ids)))
        self._optimize_normalization_weights = False
        # the reference decodes in fresh worker processes (PD:249-268); here the same object decodes, so a
        # filtered table left by an earlier call must not redirect the per-tile saves
        self._barcodes_filtered = False

        def per_tile(dec, tile_idx, gpu):
            dec.decode_one_tile(
                tile_idx=tile_idx, gpu_id=gpu, lowpass_sigma=lowpass_sigma,
                magnitude_threshold=magnitude_threshold, minimum_pixels=minimum_pixels,
                feature_predictor_threshold=feature_predictor_threshold,
                normalization_method=normalization_method,
            )
            dec._save_barcodes()
            dec._cleanup(keep_pipeline=True)

        # Resolve the normalisation vectors ONCE before the tiles are handed out: with
        # normalization_method="global" and nothing cached, every rank's first tile would otherwise compute its
        # own vectors from its own unseeded random sample of tiles (PD:1018) and all ranks would write
        # calibrations/attributes.json at the same time.  Rank 0 computes and saves, the others load.
        rank0, _w, _d = self._dist()
        if rank0 == 0:
            self._prepare_normalization_state(normalization_method, True, self._local_gpu(), lowpass_sigma)
        self._barrier()
        if rank0 != 0:
            self._prepare_normalization_state(normalization_method, True, self._local_gpu(), lowpass_sigma)

        self._run_tiles(all_tiles, per_tile)
        self._cleanup()
        self._barrier()
        self._load_tile_decoding = True
        self._load_all_barcodes()
        if self._verbose >= 1:
            print(f"Number of loaded barcodes: {len(self._df_barcodes_loaded)}")
        rank, _world, _dist = self._dist()
        if rank != 0:  # the pooled table stage runs once (the reference runs it in the parent process)
            self._barrier()
            return
        self._finish_filtering(assign_to_cells, duplicate_radius_xy, duplicate_radius_z, filter_method,
                               target_gross_misid_rate, lr_fdr_target, overlap=len(all_tiles) > 1)
        self._barrier()

    def _finish_filtering(self, assign_to_cells, duplicate_radius_xy, duplicate_radius_z, filter_method,
                          target_gross_misid_rate, lr_fdr_target, overlap: bool) -> None:
        """PD:4849-4870: filter -> (2-D) within-tile collapse -> tile-overlap de-duplication ->
        cell assignment -> ``all_tiles_filtered_decoded_features``."""
        self._apply_filter_method(filter_method, float(target_gross_misid_rate), float(lr_fdr_target))
        if not self._is_3D:
            vs = self._datastore.voxel_size_zyx_um
            radius_xy = vs[-1] if duplicate_radius_xy is None else float(duplicate_radius_xy)
            radius_z = vs[0] if duplicate_radius_z is None else float(duplicate_radius_z)
            self._remove_duplicates_within_tile(radius_xy=radius_xy, radius_z=radius_z)
        if overlap:
            self._remove_duplicates_in_tile_overlap()
        if assign_to_cells:
            self._assign_cells()
        self._save_barcodes()

    @staticmethod
    def _validate_filter_configuration(filter_method, target_gross_misid_rate, lr_fdr_target) -> None:
        """PD:4872-4913."""
        if filter_method == "blank_fraction":
            if lr_fdr_target != 0.05:
                raise ValueError(
                    "lr_fdr_target only applies when filter_method='lr'. "
                    "Use target_gross_misid_rate with filter_method='blank_fraction'."
                )
            return
        if filter_method == "lr":
            if target_gross_misid_rate != 0.05:
                raise ValueError(
                    "target_gross_misid_rate only applies when "
                    "filter_method='blank_fraction'. Use lr_fdr_target with "
                    "filter_method='lr'."
                )
            return
        raise ValueError("filter_method must be one of 'blank_fraction' or 'lr'.")

    def optimize_filtering(
        self,
        assign_to_cells: bool = True,
        duplicate_radius_xy: float | None = None,
        duplicate_radius_z: float | None = None,
        filter_method: Literal["blank_fraction", "lr"] = "blank_fraction",
        target_gross_misid_rate: float = 0.05,
        lr_fdr_target: float = 0.05,
    ) -> None:
        """PD:4952-5029: re-apply the downstream filters to previously decoded tiles
        (``decoded/<tile>_decoded_features.parquet``)."""
        if self._verbose >= 1:
            print("reprocess existing: load decoded transcripts")
        self._load_tile_decoding = True
        self._load_all_barcodes()
        if self._verbose >= 1:
            print(f"Number of loaded barcodes: {len(self._df_barcodes_loaded)}")
        self._load_tile_decoding = False
        self._validate_filter_configuration(filter_method, float(target_gross_misid_rate), float(lr_fdr_target))
        n_tiles = len(self._datastore.tile_ids)
        self._apply_filter_method(filter_method, float(target_gross_misid_rate), float(lr_fdr_target))
        if n_tiles or not self._is_3D:  # PD:5002-5019: 2-D collapses within tiles, 3-D de-duplicates overlaps
            if not self._is_3D:
                vs = self._datastore.voxel_size_zyx_um
                radius_xy = vs[-1] if duplicate_radius_xy is None else float(duplicate_radius_xy)
                radius_z = vs[0] if duplicate_radius_z is None else float(duplicate_radius_z)
                self._remove_duplicates_within_tile(radius_xy=radius_xy, radius_z=radius_z)
            else:
                self._remove_duplicates_in_tile_overlap()
        if assign_to_cells:
            self._assign_cells()
        self._save_barcodes()
        if self._verbose >= 1:
            print(f"Number of retained barcodes: {len(self._df_filtered_barcodes)}")


__all__ = ["PixelDecoder", "ChromaticAffineEstimationConfig", "M3dError"]
