#!/usr/bin/env python
"""Benchmark of the PixelDecoder hot path: decoded Gvoxel/s and % of HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one synthetic tile of BASELINE.json's configs[1]
(16 bits x 100 z x 2048 x 2048 uint16, 140-word MHD4 codebook): fused decode -> connected
components -> regionprops.  At N > 1 every rank decodes its own tile (tile sharding, no
data-path collective): weak scaling.  Prints ONE JSON line on rank 0.

`value`  : device-resident throughput (stack already in HBM), CUDA events, max over ranks.
`e2e`    : the same step through ``PixelDecoder.decode_one_tile`` with the stack in pinned
           HOST memory: H2D of the whole stack and D2H of the feature table inside the
           timed region.
`roofline`: decode_gate_kernel (the streaming kernel that reads the whole stack), timed with
           CUDA events recorded by the library around that launch on its own stream.
`cpu_baseline`: the NumPy/SciPy oracle on a bounded sample, fanned out over the host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = "configs[1]: single 3D tile 16 bits x 100 z x 2048 x 2048 uint16, 140-codeword MHD4 codebook"
SHAPE = (100, 2048, 2048)
N_BITS = 16
SEED = 2002
MAG = (1.5, 10.0)
MIN_PX = 16.0
BKG, NRM = 200.0, 900.0
# algorithmic bytes per voxel of the dominant (gate) kernel: it reads bits x 2 B of uint16 input.
# The 2 B/voxel dense "decoded = -1" fill of SURVEY's 34 B figure is not part of the steady-state step any more: the
# decoded image is persistent and only the previous tile's foreground voxels are reset (m3d_decode_label_persistent).
GATE_BYTES_PER_VOXEL = N_BITS * 2
CPU_SAMPLE_SHAPE = (8, 256, 256)


def bench_config(n_gpus: int) -> dict:
    """The `config` object of the JSON line -- the SAME dict in both arms (GPU and `--impl reference`)."""
    return {
        "workload": WORKLOAD,
        "step": "decode (scale, clip, L2-normalise, nearest codeword, gates) + connected components + size filters + "
                "regionprops of one tile",
        "lowpass": "off (north_star kernel sequence)",
        "thresholds": {"magnitude": list(MAG), "minimum_pixels": MIN_PX, "maximum_pixels": 500},
        "normalization": {"background": BKG, "normalization": NRM},
        "parallelism": f"tile-sharded x{n_gpus}",
        "l2": "every step reads a whole tile (13.4 GB on the GPU, fresh sub-tiles on the CPU arm): far beyond any cache",
    }


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a
    thread (the timed region of the resident-stack bench is only tens of ms, too short for
    nvidia-smi's process start-up), nvidia-smi as the fallback when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.rows = []  # (sm_mhz, sm_max_mhz, [reason names], time)
        self.t_start = self.t_end = None
        self._stop = threading.Event()
        self._t = None
        self.source = "nvml"
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            self.source = "nvidia-smi"

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self._h))
        bits = {
            "hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        self.rows.append((sm, self._max, [k for k, b in bits.items() if r & b], time.perf_counter()))

    def _sample_smi(self):
        out = subprocess.run(
            ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
            capture_output=True, text=True, timeout=5,
        ).stdout.strip()
        if out:
            c = [v.strip() for v in out.splitlines()[0].split(",")]
            self.rows.append((float(c[0]), float(c[1]),
                              [n for n, v in zip(self.NAMES, c[2:6]) if v.lower().startswith("active")],
                              time.perf_counter()))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.002 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def mark_start(self):
        self.t_start = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        rows, window = self.rows, "warm-up + timed steps (same load)"
        if self.t_start is not None and self.t_end is not None:
            inside = [r for r in self.rows if self.t_start <= r[3] <= self.t_end]
            if len(inside) >= 3:
                rows, window = inside, "timed steps"
        reasons = sorted({r for row in rows for r in row[2]})
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_max_mhz": float(max(r[1] for r in rows)),
                "reasons": reasons, "samples": len(rows), "window": window, "source": self.source}


# ---------------------------------------------------------------------------------- CPU arm
_CPU_STATE = {}


def _cpu_init(counter=None, barrier=None):
    """Per-process set-up outside the timed region: imports, codebook and this worker's own sub-tile (every
    worker has its data before the first task arrives, whichever worker a task lands on)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from merfish3d_analysis_b200 import synthetic
    from oracle import decode_oracle as orc

    m = synthetic.mhd4_codebook_matrix(16)
    _CPU_STATE["m"] = m
    _CPU_STATE["cb"] = orc.load_codebook(synthetic.codebook_dataframe(m, n_blank=10), 16)
    idx = 0
    if counter is not None:
        with counter.get_lock():
            idx = counter.value
            counter.value += 1
    _CPU_STATE["stack"] = synthetic.make_stack(m, CPU_SAMPLE_SHAPE, SEED + idx)
    _CPU_STATE["barrier"] = barrier


def _cpu_ready(_):
    """Blocks until every worker of the pool is initialised (one such task per worker)."""
    b = _CPU_STATE.get("barrier")
    if b is not None:
        try:
            b.wait(timeout=120)
        except Exception:
            pass
    return os.getpid()


def _cpu_one(_):
    """One bounded sample through the oracle (decode -> CCL -> regionprops), single process.
    Returns (seconds total, seconds of the per-voxel decode alone, transcripts)."""
    from oracle import decode_oracle as orc

    bkg = np.full(16, BKG, dtype=np.float32)
    nrm = np.full(16, NRM, dtype=np.float32)
    cb = _CPU_STATE["cb"]
    t0 = time.perf_counter()
    stack = orc.weight_readout(_CPU_STATE["stack"], None)
    unit = orc.normalize_codebook(cb["matrix"][:, : stack.shape[0]])
    out = orc.decode_pixels(stack, unit, bkg, nrm, cb["pixel_assignment_threshold"], MAG, ())
    t1 = time.perf_counter()
    df = orc.extract_barcodes(out, out["scaled"], cb["matrix"], cb["gene_ids"], True, MIN_PX,
                              cb["transcript_distance_threshold"])
    return time.perf_counter() - t0, t1 - t0, len(df)


def cpu_reference_throughput(n_rounds: int = 1):
    """Oracle port on all host cores: one sub-tile per process, like the reference's own
    tile-level parallelism (PD:4811-4838).  Data generation and imports are outside the timed
    region.  Returns (Gvoxel/s, cores, seconds, description)."""
    from concurrent.futures import ProcessPoolExecutor

    cores = os.cpu_count() or 1
    vox = int(np.prod(CPU_SAMPLE_SHAPE))
    import multiprocessing as mp

    counter = mp.Value("i", 0)
    barrier = mp.Barrier(cores)
    with ProcessPoolExecutor(max_workers=cores, initializer=_cpu_init, initargs=(counter, barrier)) as ex:
        # one blocking task per worker: all workers exist and have generated their sub-tile before the timer starts
        pids = set(ex.map(_cpu_ready, range(cores)))
        n_tasks = cores * n_rounds
        t0 = time.perf_counter()
        res = list(ex.map(_cpu_one, range(n_tasks)))
        wall = time.perf_counter() - t0
    frac = sum(r[1] for r in res) / max(sum(r[0] for r in res), 1e-9)  # share of the per-voxel decode
    decode_only = n_tasks * vox / (wall * frac) / 1e9 if frac > 0 else 0.0
    sample = (f"{n_tasks} sub-tiles of 16x{CPU_SAMPLE_SHAPE[0]}x{CPU_SAMPLE_SHAPE[1]}x{CPU_SAMPLE_SHAPE[2]} uint16 "
              f"(same value model / codebook / thresholds as the workload) over {cores} worker processes, "
              "NumPy/SciPy oracle: decode + CCL + regionprops, no low-pass; "
              f"per-voxel decode alone = {100 * frac:.0f}% of the time ({decode_only:.4f} Gvoxel/s decode-only)")
    return n_tasks * vox / wall / 1e9, cores, wall, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_throughput(1)
    times, vals = [], []
    for _ in range(args.steps):
        g, cores, wall, sample = cpu_reference_throughput(1)
        vals.append(g)
        times.append(wall)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "decoded Gvoxel/s", "value": value, "unit": "Gvoxel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------- GPU arm
def _zarr_store_extra(sub, df_cb, nrm, bkg, tmp_dir, local, compression="blosc-zstd"):
    """decode_one_tile with the tile read from a reference-layout datastore (chunk files, page-cache warm)."""
    from concurrent.futures import ThreadPoolExecutor

    import torch

    from merfish3d_analysis_b200 import zarr_store as zs
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    root = Path(tmp_dir) / f"zarr_{compression}" / "qi2labdatastore"
    zds = zs.Qi2labZarrDataStore.create(root, df_cb, num_tiles=1)

    def write_bit(b):
        zs.write_ome_image(root / "readouts" / "tile0000" / f"bit{b + 1:03d}" / "corrected_data", sub[b],
                           compression=compression, extra_attributes={"round_linker": 1, "excitation_um": 0.561, "emission_um": 0.58})

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        list(ex.map(write_bit, range(sub.shape[0])))
    zds._save_entity_attributes(root / "fiducial" / "tile0000" / "round001", {
        "stage_zyx_um": [0, 0, 0], "affine_zyx_px": np.eye(4), "local_round_transform_zyx_um": np.eye(4)})
    stored = sum(f.stat().st_size for f in (root / "readouts").rglob("*") if f.is_file())
    zds = zs.Qi2labZarrDataStore(root)
    zds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = PixelDecoder(zds, merfish_bits=16, verbose=0)
    for i in range(3):
        if i == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        dec.decode_one_tile(0, gpu_id=local, lowpass_sigma=None, magnitude_threshold=MAG, minimum_pixels=MIN_PX,
                            normalization_method="global")
    s = (time.perf_counter() - t0) / 2
    n_vox = int(np.prod(sub.shape[1:]))
    out = {
        "ms_per_step": s * 1e3, "gvoxel_per_s": n_vox / s / 1e9, "decoded_gb_s": sub.nbytes / s / 1e9,
        "stored_gb": stored / 1e9, "compression_ratio": sub.nbytes / stored, "host_threads": os.cpu_count(),
        "transcripts": int(len(dec._df_barcodes)),
        "note": f"decode_one_tile on 16 bits x {tuple(sub.shape[1:])} read from <image>.ome.zarr (Zarr v3, {compression} "
                "bitshuffle, (16,512,512) chunks, page-cache warm) through m3d_zarr_read_chunks; " + (
                    "the chunk files cross PCIe compressed and the GPU decodes the LZ4 streams (one warp each), "
                    "un-shuffles and places them" if compression == "blosc-lz4" else
                    "the chunk files cross PCIe compressed and the GPU decodes the zstd frames (a warp per Blosc block: one "
                    "lane per Huffman stream, shared copies), un-shuffles and places them"
                    if os.environ.get("M3D_ZARR_GPU_ZSTD", "2") != "0" else
                    "host threads zstd-decode into pinned slots, the GPU un-shuffles and places the chunks"),
    }
    shutil.rmtree(root.parent, ignore_errors=True)
    del dec
    torch.cuda.empty_cache()
    return out


OPT_SHAPE = (64, 2048, 2048)   # configs[2]: human-olfactory-bulb-shaped tiles
OPT_TILES_PER_RANK = 2
OPT_ITERATIONS = 4
ZSLAB_PLANES_PER_RANK = 50     # configs[4] = 16 x 400 x 4096 x 4096 is exactly 8 ranks x 50 planes
ZSLAB_YX = 4096


def _host_gb_available():
    try:
        import psutil

        return psutil.virtual_memory().available / 1e9
    except Exception:  # noqa: BLE001
        return None


def _max_over_ranks(dist, world, dev, values):
    import torch

    t = torch.tensor(list(values), device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def _optimizer_extra(dist, world, rank, local, dev, matrix, df_cb, shared_root):
    """configs[2]-shaped run of `optimize_normalization_by_decoding` (PD:4581-4757): OPT_TILES_PER_RANK tiles of
    16 x 64 x 2048 x 2048 uint16 per rank in pinned host memory, reference-default low-pass (3,1,1), one shared
    datastore directory.  Iteration 0 uploads, low-passes and decodes (dense-candidate regime: percentile-seeded
    vectors); iterations 1.. decode the HBM-resident low-passed stacks.  Exchanges: all_reduce of the percentile
    seed's digit histograms, one padded all_gather of transcript rows per iteration."""
    import torch

    from merfish3d_analysis_b200 import synthetic
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    T = OPT_TILES_PER_RANK
    n_vox = int(np.prod(OPT_SHAPE))
    need_gb = 16 * n_vox * 2 * T * world / 1e9
    avail = _host_gb_available()
    if avail is not None and need_gb * 1.2 > avail:
        return {"skipped": f"host memory: needs {need_gb:.0f} GB pinned, {avail:.0f} GB available"}
    hosts = []
    for k in range(T):
        blk = synthetic.make_stack_device(matrix, OPT_SHAPE, 3000 + rank * T + k, device=dev)
        h = torch.empty(blk.shape, dtype=torch.uint16, pin_memory=True)
        h.copy_(blk)
        hosts.append(h)
        del blk
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    root = Path(shared_root) / "qi2labdatastore_optimizer"
    if rank == 0:
        ArrayDataStore(root, codebook=df_cb)
    if world > 1:
        dist.barrier()
    ds = ArrayDataStore(root)
    for r in range(world):  # tile metadata goes into ONE calibrations/attributes.json: ranks register in turn
        if r == rank:
            ds._refresh()
            for k in range(T):
                ds.add_tile(hosts[k].numpy(), tile_id=f"tile{r * T + k:04d}")
        if world > 1:
            dist.barrier()
    ds._refresh()
    dec = PixelDecoder(ds, merfish_bits=16, num_gpus=world, verbose=0)
    # warm-up, like every other measurement of this file: one iteration over ONE tile per rank (first use of the kernels,
    # the library's labelling scratch, pinned registrations); the timed call starts with an empty tile cache again
    dec.optimize_normalization_by_decoding(n_iterations=1, lowpass_sigma=(3.0, 1.0, 1.0), magnitude_threshold=MAG,
                                           minimum_pixels=MIN_PX, tile_indices=[r * T for r in range(world)])
    dec._profile = {"sync": True}
    ctx = dec._ctx(local)
    ctx.reset_counters()
    ctx.set_timing(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec.optimize_normalization_by_decoding(n_iterations=OPT_ITERATIONS, lowpass_sigma=(3.0, 1.0, 1.0),
                                           magnitude_threshold=MAG, minimum_pixels=MIN_PX,
                                           tile_indices=list(range(world * T)))
    torch.cuda.synchronize()
    total_s = time.perf_counter() - t0
    tm = dec._optimizer_timing
    kt = ctx.kernel_times_ms()
    keys = ("total_s", "tiles_s", "local_table_s", "exchange_s", "vectors_s", "stage_s", "decode_extract_s")
    its = []
    for it in tm["iterations"]:
        vals = _max_over_ranks(dist, world, dev, [it.get(k, 0.0) for k in keys])
        its.append(dict(zip(keys, vals), cache_hits=it["cache_hits"], transcripts_pooled=it["transcripts_pooled"]))
    seed_s, total_s = _max_over_ranks(dist, world, dev, [tm["seed_s"], total_s])
    steady = its[1:]
    steady_s = float(np.mean([it["total_s"] for it in steady])) if steady else None
    i_n, i_b = ds.load_decode_normalization_vectors(None, "iterative")
    out = {
        "tiles": world * T, "tiles_per_rank": T, "tile": f"16 x {OPT_SHAPE[0]} x {OPT_SHAPE[1]} x {OPT_SHAPE[2]} uint16",
        "iterations": OPT_ITERATIONS, "lowpass_sigma": [3.0, 1.0, 1.0], "total_s": total_s, "seed_s": seed_s,
        "iteration0": its[0] if its else None, "steady_iterations": steady,
        "steady_s_per_iteration": steady_s,
        "steady_gvoxel_per_s": (world * T * n_vox / steady_s / 1e9) if steady_s else None,
        "iteration0_gvoxel_per_s": (world * T * n_vox / its[0]["total_s"] / 1e9) if its else None,
        "cache": tm.get("cache"),
        "kernel_ms_whole_run_rank0": {k: v for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:8]},
        "iterative_normalization_head": [float(v) for v in np.asarray(i_n)[:4]],
        "exchange": "seed: all_reduce(SUM) of 2048-bin int64 digit histograms (radix select over tiles sharded on "
                    "ranks); per iteration: 3 all_reduce(SUM) of a (<= 64, 2048) int64 histogram tensor (pooled per-bit "
                    "medians by radix select; no table leaves its rank)" if world > 1 else "single process: no collective",
        "note": "times are max over ranks; stage_s = datastore read + H2D + per-bit low-pass of tiles not yet resident "
                "(0 once cached), decode_extract_s = decode + label + features + annotation, local_table_s = concatenating this rank's tables, exchange_s = pooled "
                "medians (histogram all_reduce), vectors_s = JSON hand-off; all phases separated by stream syncs",
    }
    ctx.set_timing(False)
    dec._cleanup()
    del dec, hosts
    torch.cuda.empty_cache()
    return out


def _zslab_extra(dist, world, rank, local, dev, matrix, df_cb, tmp_dir):
    """configs[4]-shaped z-slab sharding: ONE volume of 16 x (50 N) x 4096 x 4096 uint16 (N = 8 is configs[4]
    itself), every rank holding only its own 50 planes in pinned host memory, through
    `decode_one_tile_sharded`: H2D, decode + label per slab, boundary-plane send/recv, all_gather of the
    cross-slab equivalences, gather of the crossing components' records, assembly on rank 0."""
    import pandas as pd
    import torch

    from merfish3d_analysis_b200 import synthetic
    from merfish3d_analysis_b200.datastore import ArrayDataStore, ZWindowVolume
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    per, yx = ZSLAB_PLANES_PER_RANK, ZSLAB_YX
    avail = _host_gb_available()
    note = None
    while avail is not None and 16 * per * yx * yx * 2 * world * 1.2 / 1e9 > avail and per > 10:
        per //= 2
        note = f"planes per rank reduced to {per}: host memory ({avail:.0f} GB available)"
    Z = per * world
    nrm = np.full(16, NRM, dtype=np.float32)
    bkg = np.full(16, BKG, dtype=np.float32)
    block = torch.empty((16, per, yx, yx), dtype=torch.uint16, pin_memory=True)
    step = 10
    for z0 in range(0, per, step):  # generated on the device a few planes at a time (data synthesis only)
        z1 = min(per, z0 + step)
        blk = synthetic.make_stack_device(matrix, (z1 - z0, yx, yx), 5005 + rank * 100 + z0, device=dev)
        block[:, z0:z1].copy_(blk)
        del blk
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ds = ArrayDataStore(Path(tmp_dir) / f"qi2labdatastore_zslab_r{rank}", codebook=df_cb)
    ds.add_tile([ZWindowVolume((Z, yx, yx), rank * per, block[b].numpy()) for b in range(16)])
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = PixelDecoder(ds, merfish_bits=16, num_gpus=world, verbose=0)
    dec._profile = {"sync": True}
    times, splits = [], []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec.decode_one_tile_sharded(0, lowpass_sigma=None, magnitude_threshold=MAG, minimum_pixels=MIN_PX,
                                    normalization_method="global")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
        splits.append(dict(dec._zslab_timing))
    n_tr = len(dec._df_barcodes) if rank == 0 else 0
    keys = sorted(splits[-1])
    best = int(np.argmin(times[1:])) + 1
    split = dict(zip(keys, _max_over_ranks(dist, world, dev, [splits[best].get(k, 0.0) for k in keys])))
    t_best = _max_over_ranks(dist, world, dev, [times[best]])[0]
    n_vox = Z * yx * yx
    dec._cleanup()
    del dec, block
    torch.cuda.empty_cache()
    # identical-to-unsharded check on a volume that fits one GPU (every rank holds all of it)
    identical = None
    try:
        zs_, ys_ = 6 * max(world, 2), 512
        small = synthetic.make_stack_device(matrix, (zs_, ys_, ys_), 777, device=dev, density=4e-4).cpu().numpy()
        ds2 = ArrayDataStore(Path(tmp_dir) / f"qi2labdatastore_zcheck_r{rank}", codebook=df_cb)
        ds2.add_tile(small)
        ds2.save_decode_normalization_vectors(None, "global", nrm, bkg)
        d_sh = PixelDecoder(ds2, merfish_bits=16, num_gpus=world, verbose=0)
        kw = dict(lowpass_sigma=None, magnitude_threshold=MAG, minimum_pixels=MIN_PX, normalization_method="global")
        if world > 1:
            d_sh.decode_one_tile_sharded(0, **kw)
        else:
            d_sh.decode_one_tile_sharded(0, n_slabs=3, **kw)
        if rank == 0:
            d_un = PixelDecoder(ds2, merfish_bits=16, num_gpus=1, verbose=0)
            d_un.decode_one_tile(0, gpu_id=local, **kw)
            pd.testing.assert_frame_equal(d_sh.decoded_barcodes, d_un.decoded_barcodes)
            identical = {"ok": True, "volume": f"16 x {zs_} x {ys_} x {ys_}", "transcripts": int(len(d_un.decoded_barcodes)),
                         "slabs": world if world > 1 else 3}
    except AssertionError as e:
        identical = {"ok": False, "error": str(e)[:300]}
    return {
        "volume": f"16 x {Z} x {yx} x {yx} uint16", "planes_per_rank": per, "gb_total": 16 * n_vox * 2 / 1e9,
        "s_per_volume": t_best, "gvoxel_per_s": n_vox / t_best / 1e9, "runs_s": times, "transcripts": int(n_tr),
        "split_s_max_over_ranks": split, "identical_to_unsharded": identical,
        "exchange": "boundary planes: NCCL send/recv of one (2, Y, X) int32 message per interface; equivalences / areas: "
                    "sizes + padded float64 all_gather; crossing components' records + kept rows: padded float64 gather "
                    "to rank 0" if world > 1 else "one slab, no exchange",
        "note": note,
    }


def _h2d_rates(dist, world, dev, host):
    """Concurrent pinned H2D of every rank's tile: per-rank GB/s (the host fabric's ceiling for `e2e` at N > 1)."""
    import torch

    dst = torch.empty(host.shape, dtype=host.dtype, device=dev)
    dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dst.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = host.numel() * host.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    if world > 1:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        rates = [float(p.item()) for p in parts]
    else:
        rates = [gbs]
    del dst
    return rates


def run_b200(args):
    import torch
    import torch.distributed as dist

    from merfish3d_analysis_b200 import synthetic
    from merfish3d_analysis_b200._capi import DecodeContext
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    shape = tuple(args.shape) if args.shape else SHAPE
    n_vox = int(np.prod(shape))
    matrix = synthetic.mhd4_codebook_matrix(16)
    df_cb = synthetic.codebook_dataframe(matrix, n_blank=10)
    bkg = np.full(16, BKG, dtype=np.float32)
    nrm = np.full(16, NRM, dtype=np.float32)

    # synthetic tile built directly in HBM (data synthesis only; not on the timed path)
    stack = synthetic.make_stack_device(matrix, shape, SEED + rank, device=dev)
    torch.cuda.synchronize()

    tmp = tempfile.TemporaryDirectory()
    ds = ArrayDataStore(Path(tmp.name) / f"qi2labdatastore_r{rank}", codebook=df_cb)
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    unit = dec._decoding_matrix.astype(np.float32)
    ctx = DecodeContext(unit, (), device=local)
    ctx.set_normalization(bkg, nrm)
    ctx.set_thresholds(dec._pixel_assignment_threshold, MAG[0], MAG[1])
    decoded = torch.empty(shape, dtype=torch.int16, device=dev)

    def step_resident():
        # `decoded` is the persistent per-GPU image of the production path (PixelDecoder._decode_pixels)
        n = ctx.decode_label(stack, decoded, False, MIN_PX, 500, persistent=True)
        table = ctx.features(stack, decoded, False, n)
        return n, table

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the sampler thread runs from the warm-up on (a timed region of a few tens of ms is shorter than a thread
    # start-up); summary() prefers the samples that fall inside the timed window
    with ClockSampler(local) as clocks:
        for _ in range(max(args.warmup, 3)):
            n_feat, table = step_resident()
        barrier()
        ctx.reset_counters()
        ctx.set_timing(True)
        barrier()
        clocks.mark_start()
        e0.record()
        for _ in range(args.steps):
            n_feat, table = step_resident()
        e1.record()
        barrier()
        clocks.mark_end()
    ms_total = e0.elapsed_time(e1)
    ktimes = ctx.kernel_times_ms()
    klaunch = ctx.launches_by_kernel()
    ctx.set_timing(False)
    n_fg = int((decoded >= 0).sum().item())
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step_serial = float(t.item()) / args.steps
    gpu_launches = int(sum(klaunch.values()))
    ms_step = ms_step_serial

    value = world * n_vox / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (per launch, measured live)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gate_ms = ktimes.get("decode_gate_kernel", 0.0) / max(klaunch.get("decode_gate_kernel", 1), 1)
    achieved = GATE_BYTES_PER_VOXEL * n_vox / (gate_ms * 1e-3) / 1e9 if gate_ms else 0.0
    # measured DRAM traffic of that kernel from the committed ncu --set full capture (same shape only)
    traffic = None
    tf = ROOT / "profiles" / "r2_gate_traffic.json"
    if tf.exists():
        tj = json.loads(tf.read_text())
        if tuple(tj.get("shape_zyx", ())) == tuple(shape) and tj.get("n_bits") == N_BITS:
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "decode_gate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write)",
        "algorithmic_gb_per_launch": GATE_BYTES_PER_VOXEL * n_vox / 1e9,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6.65 TB/s (of fallback)",
        "ms_per_launch": gate_ms, "algorithmic_bytes_per_voxel": GATE_BYTES_PER_VOXEL,
        "frac_of_nominal_8TBs": achieved / 8000.0,
        "kernel_ms_per_step": {k: v / args.steps for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1])},
    }

    # ---- secondary measurements (not the headline): reference-default low-pass on, and the
    # all-foreground worst case where every voxel passes the magnitude gate
    extras = {}
    want = set((args.extras or "all").split(","))

    def wanted(name):
        return not args.no_extras and ("all" in want or name in want)

    if world == 1 and (wanted("lowpass_on") or wanted("all_foreground")):
        ctx.reset_counters()
        ctx.set_timing(True)
        lp = None
        for i in range(3):
            if i == 1:
                torch.cuda.synchronize()
                ctx.reset_counters()
                t0 = time.perf_counter()
            lp = ctx.lowpass(stack, (3.0, 1.0, 1.0), False, out=lp)
            nl = ctx.decode_label(lp, decoded, False, MIN_PX, 500)
            ctx.features(lp, decoded, False, nl)
        torch.cuda.synchronize()
        lp_ms = (time.perf_counter() - t0) * 1e3 / 2
        kt = ctx.kernel_times_ms()
        extras["lowpass_on"] = {
            "ms_per_step": lp_ms, "gvoxel_per_s": n_vox / lp_ms / 1e6, "features": int(nl),
            "note": "m3d_lowpass sigma=(3,1,1) (SciPy-exact fp64 accumulation) + decode_label + features on float32",
            "kernel_ms_per_step": {k: v / 2 for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:6]},
        }
        # the opt-in float32 accumulation (CuPy-style arithmetic, not pinned): HBM-bound instead of float64-pipe-bound
        ctx.set_lowpass_accumulate("float32")
        for i in range(3):
            if i == 1:
                torch.cuda.synchronize()
                ctx.reset_counters()
                t0 = time.perf_counter()
            lp = ctx.lowpass(stack, (3.0, 1.0, 1.0), False, out=lp)
        torch.cuda.synchronize()
        lp32_ms = (time.perf_counter() - t0) * 1e3 / 2
        kt = ctx.kernel_times_ms()
        ctx.set_lowpass_accumulate("float64")
        extras["lowpass_float32_mode"] = {
            "ms_per_tile": lp32_ms, "gb_s_materialised": (2 + 4 + 4 + 4) * N_BITS * n_vox / lp32_ms / 1e6,
            "kernel_ms_per_tile": {k: v / 2 for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:2]},
            "note": "m3d_lowpass alone with m3d_set_lowpass_mode(ctx, 1): float32 weights + FMA accumulation (what cupyx's "
                    "filter is believed to do; opt-in, results differ from the SciPy-exact default by a few float32 ulps); "
                    "traffic = uint16 read + float32 temporary write/read + float32 write per voxel-bit",
        }
        del lp
        torch.cuda.empty_cache()
        # worst case: thresholds that let every voxel through the magnitude gate
        ctx.set_thresholds(dec._pixel_assignment_threshold, 1.0e-3, 10.0)
        ctx.set_normalization(np.full(16, 0.0, dtype=np.float32), np.full(16, 250.0, dtype=np.float32))
        ctx.reset_counters()
        for i in range(3):
            if i == 1:
                torch.cuda.synchronize()
                ctx.reset_counters()
                t0 = time.perf_counter()
            ctx.decode(stack, decoded)
        torch.cuda.synchronize()
        wc_ms = (time.perf_counter() - t0) * 1e3 / 2
        kt = ctx.kernel_times_ms()
        extras["all_foreground"] = {
            "ms_per_decode": wc_ms, "gvoxel_per_s": n_vox / wc_ms / 1e6,
            "decoded_fraction": float((decoded >= 0).float().mean().item()),
            "note": "every voxel is a candidate (magnitude gate open): exact search at all voxels, m3d_decode only",
            "kernel_ms_per_step": {k: v / 2 for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:3]},
        }
        ctx.set_timing(False)
        ctx.set_normalization(bkg, nrm)
        ctx.set_thresholds(dec._pixel_assignment_threshold, MAG[0], MAG[1])

    # ---- e2e: the reference-facing call, stack in pinned host memory
    e2e = None
    if not args.no_e2e:
        host = torch.empty(stack.shape, dtype=torch.uint16, pin_memory=True)
        host.copy_(stack)
        torch.cuda.synchronize()
        ds.add_tile(host.numpy())
        ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
        del stack, decoded
        torch.cuda.empty_cache()

        def step_e2e():
            dec.decode_one_tile(0, gpu_id=local, lowpass_sigma=None, magnitude_threshold=MAG, minimum_pixels=MIN_PX,
                                normalization_method="global")
            return dec._df_barcodes

        for _ in range(2):
            df = step_e2e()
        barrier()
        n_e2e = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            df = step_e2e()
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e_s = float(tw.item()) / n_e2e
        e2e = {
            "value": world * n_vox / e2e_s / 1e9, "unit": "Gvoxel/s",
            "h2d_bytes_per_step": int(host.numel() * 2),
            "d2h_bytes_per_step": int(n_feat * (14 + N_BITS) * 8),
            "ms_per_step": e2e_s * 1e3, "steps": n_e2e, "transcripts": int(len(df)),
            "api": "PixelDecoder.decode_one_tile(lowpass_sigma=None, normalization_method='global')",
        }

        # the reference-default call (low-pass sigma (3,1,1) on): the per-bit filter runs behind the upload
        if world == 1 and wanted("e2e_variants"):
            for i in range(3):
                if i == 1:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                dec.decode_one_tile(0, gpu_id=local, lowpass_sigma=(3.0, 1.0, 1.0), magnitude_threshold=MAG,
                                    minimum_pixels=MIN_PX, normalization_method="global")
            torch.cuda.synchronize()
            lp_s = (time.perf_counter() - t0) / 2
            extras["e2e_lowpass_on"] = {
                "ms_per_step": lp_s * 1e3, "gvoxel_per_s": n_vox / lp_s / 1e9, "transcripts": int(len(dec._df_barcodes)),
                "note": "decode_one_tile(lowpass_sigma=(3,1,1)) from pinned host memory; bit b is low-passed "
                        "(SciPy-exact, float64 pipe) while bits b+1.. are still crossing PCIe",
            }
            dec._cleanup()
            torch.cuda.empty_cache()
        # unregistered tile (the usual case: only round-1 bits are in the reference frame): 12 of 16 bits carry an
        # affine round transform and are resampled on the device, then the reference-default low-pass
        if world == 1 and wanted("e2e_variants"):
            rng_x = np.random.default_rng(7)
            xf = {}
            for r in (2, 3, 4):
                m_ = np.eye(4)
                m_[:3, :3] += rng_x.normal(0, 1e-3, (3, 3))
                m_[:3, 3] = rng_x.normal(0, 1.0, 3) * np.array([0.315, 0.098, 0.098])
                xf[r] = m_
            ds4 = ArrayDataStore(Path(tmp.name) / "qi2labdatastore_unreg", codebook=df_cb)
            ds4.add_tile(host.numpy(), bit_round=[1 + b // 4 for b in range(16)], round_transforms_zyx_um=xf)
            ds4.save_decode_normalization_vectors(None, "global", nrm, bkg)
            dec4 = PixelDecoder(ds4, merfish_bits=16, verbose=0)
            for i in range(3):
                if i == 1:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                dec4.decode_one_tile(0, gpu_id=local, lowpass_sigma=(3.0, 1.0, 1.0), magnitude_threshold=MAG,
                                     minimum_pixels=MIN_PX, normalization_method="global")
            torch.cuda.synchronize()
            un_s = (time.perf_counter() - t0) / 2
            extras["e2e_unregistered_lowpass_on"] = {
                "ms_per_step": un_s * 1e3, "gvoxel_per_s": n_vox / un_s / 1e9, "transcripts": int(len(dec4._df_barcodes)),
                "note": "decode_one_tile on a tile whose bits 5-16 need the decode-time affine warp (m3d_warp_affine, "
                        "SciPy-exact float64 taps) + low-pass (3,1,1): both run per bit behind the upload of the next bits",
            }
            dec4._cleanup()
            del dec4
            torch.cuda.empty_cache()
        # multi-tile public API: decode_all_tiles over 3 tiles (the same pinned array registered three times),
        # per-tile parquet output + the pooled table stage; tile t+1 is staged while tile t is finished
        if world == 1 and wanted("e2e_variants"):
            ds3 = ArrayDataStore(Path(tmp.name) / "qi2labdatastore_multi", codebook=df_cb)
            for k in range(3):
                ds3.add_tile(host.numpy(), stage_origin_zyx_um=(0.0, 0.0, 250.0 * k))
            ds3.save_decode_normalization_vectors(None, "global", nrm, bkg)
            dec3 = PixelDecoder(ds3, merfish_bits=16, verbose=0)
            all_s = None
            for _rep in range(2):  # the first pass allocates both staging slots; the second is the steady state
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                dec3.decode_all_tiles(assign_to_cells=False, lowpass_sigma=None, magnitude_threshold=MAG,
                                      minimum_pixels=MIN_PX, normalization_method="global")
                torch.cuda.synchronize()
                all_s = time.perf_counter() - t0
            extras["decode_all_tiles_3"] = {
                "s_total": all_s, "ms_per_tile": all_s / 3 * 1e3, "gvoxel_per_s": 3 * n_vox / all_s / 1e9,
                "filtered_transcripts": int(len(dec3._df_filtered_barcodes)),
                "note": "decode_all_tiles: 3 tiles from pinned host memory, per-tile parquet files, blank-fraction "
                        "filter + tile-overlap de-duplication + filtered table written; next tile prefetched",
            }
            dec3._cleanup()
            del dec3
            torch.cuda.empty_cache()
        # the same call with the stack in PAGEABLE host memory (what a datastore returning plain NumPy
        # arrays gives the loader): staged through the library's pinned ring by m3d_upload_batch
        if world == 1 and wanted("e2e_variants"):
            pageable = np.empty(tuple(host.shape), dtype=np.uint16)
            np.copyto(pageable, host.numpy())
            ds2 = ArrayDataStore(Path(tmp.name) / "qi2labdatastore_pageable", codebook=df_cb)
            ds2.add_tile(pageable)
            ds2.save_decode_normalization_vectors(None, "global", nrm, bkg)
            dec2 = PixelDecoder(ds2, merfish_bits=16, verbose=0)
            for i in range(3):
                if i == 1:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                dec2.decode_one_tile(0, gpu_id=local, lowpass_sigma=None, magnitude_threshold=MAG,
                                     minimum_pixels=MIN_PX, normalization_method="global")
            pg_s = (time.perf_counter() - t0) / 2
            t0 = time.perf_counter()
            plain = torch.from_numpy(pageable[0]).to(dev)
            torch.cuda.synchronize()
            plain_gbs = pageable[0].nbytes / (time.perf_counter() - t0) / 1e9
            extras["e2e_pageable_host"] = {
                "ms_per_step": pg_s * 1e3, "gvoxel_per_s": n_vox / pg_s / 1e9,
                "staged_h2d_gb_s_incl_decode": host.numel() * 2 / pg_s / 1e9,
                "plain_pageable_cudaMemcpy_gb_s": plain_gbs,
                "note": "decode_one_tile with the tile in pageable NumPy memory; m3d_upload_batch stages through "
                        "8 x 32 MB pinned slots filled by host threads",
            }
            del plain, dec2, pageable
        # the same call with the tile in the reference's on-disk form: `<image>.ome.zarr` Zarr v3 arrays of
        # blosc-zstd bit-shuffled (16, 512, 512) chunks.  32 planes keep the store small; the
        # figure of merit is decoded GB/s.  Guarded: a full disk must not cost the bench line.
        if world == 1 and wanted("zarr"):
            for key, comp, env in (("e2e_from_zarr_store", "blosc-zstd", None),
                                   ("e2e_from_zarr_store_host_zstd", "blosc-zstd", "0"),
                                   ("e2e_from_zarr_store_lz4", "blosc-lz4", None)):
                saved_env = os.environ.get("M3D_ZARR_GPU_ZSTD")
                try:
                    if env is not None:
                        os.environ["M3D_ZARR_GPU_ZSTD"] = env
                    extras[key] = _zarr_store_extra(host.numpy()[:, :32], df_cb, nrm, bkg, tmp.name, local, comp)
                except Exception as e:  # noqa: BLE001
                    extras[key] = {"error": f"{type(e).__name__}: {e}"}
                finally:
                    if saved_env is None:
                        os.environ.pop("M3D_ZARR_GPU_ZSTD", None)
                    else:
                        os.environ["M3D_ZARR_GPU_ZSTD"] = saved_env
            # the same 32 planes from pinned host memory, for scale
            try:
                ds3 = ArrayDataStore(Path(tmp.name) / "qi2labdatastore_32", codebook=df_cb)
                ds3.add_tile(host.numpy()[:, :32])
                ds3.save_decode_normalization_vectors(None, "global", nrm, bkg)
                dec3 = PixelDecoder(ds3, merfish_bits=16, verbose=0)
                for i in range(3):
                    if i == 1:
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                    dec3.decode_one_tile(0, gpu_id=local, lowpass_sigma=None, magnitude_threshold=MAG,
                                         minimum_pixels=MIN_PX, normalization_method="global")
                extras["e2e_32_planes_from_pinned_host"] = {"ms_per_step": (time.perf_counter() - t0) / 2 * 1e3}
                del dec3
            except Exception as e:  # noqa: BLE001
                extras["e2e_32_planes_from_pinned_host"] = {"error": f"{type(e).__name__}: {e}"}

    if e2e is not None:
        try:
            rates = _h2d_rates(dist, world, dev, host)
            e2e["h2d_gb_s_per_rank_concurrent"] = [round(r, 2) for r in rates]
            e2e["h2d_gb_s_total"] = round(float(sum(rates)), 2)
        except Exception as ex:  # noqa: BLE001
            e2e["h2d_gb_s_per_rank_concurrent"] = f"{type(ex).__name__}: {ex}"
    # free everything the first part of the bench held before the two sharded configurations
    try:
        del host
    except NameError:
        pass
    try:
        del stack, decoded
    except NameError:
        pass
    dec._cleanup()
    torch.cuda.empty_cache()
    import gc

    gc.collect()
    shared = [tmp.name]
    if world > 1:
        dist.broadcast_object_list(shared, src=0)  # control plane: the path of the shared datastore directory
    for name, fn, root in (("optimizer", _optimizer_extra, shared[0]), ("zslab", _zslab_extra, tmp.name)):
        if not wanted(name):
            continue
        try:
            extras[name] = fn(dist, world, rank, local, dev, matrix, df_cb, root)
        except Exception as ex:  # noqa: BLE001
            import traceback

            extras[name] = {"error": f"{type(ex).__name__}: {ex}", "trace": traceback.format_exc()[-600:]}
            if world > 1:
                raise  # a rank that left a collective early would hang the others: fail loudly
        torch.cuda.empty_cache()
        gc.collect()

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            g, cores, wall, sample = cpu_reference_throughput(1)
            cpu = {"value": g, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "sample": sample, "seconds": wall}
        line = {
            "metric": "decoded Gvoxel/s", "value": value, "unit": "Gvoxel/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": bench_config(world) if shape == SHAPE else dict(
                bench_config(world), workload=f"reduced tile 16x{shape[0]}x{shape[1]}x{shape[2]} uint16 (debug)"),
            "workload_stats": {"foreground_voxels": n_fg, "features": int(n_feat),
                               "input_gb_per_step": stack_bytes(shape) / 1e9,
                               "device_step": "m3d_decode_label_persistent (gate + search + CCL) + m3d_features on the "
                                              "HBM-resident stack"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches,
            "extras": extras,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    tmp.cleanup()
    return line


def stack_bytes(shape):
    return N_BITS * int(np.prod(shape)) * 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", type=int, nargs=3, default=None, help="z y x override (debug only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", default="all", help="comma list of extras to run (default all): lowpass_on,all_foreground,"
                    "e2e_variants,zarr,optimizer,zslab")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
