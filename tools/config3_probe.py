"""configs[2]-shaped optimiser run at reduced tile count: T tiles of 16 x 64 x 2048 x 2048 uint16, the
self-optimising per-bit normalisation loop with tiles sharded over the ranks and one NCCL all_gather of
transcript rows per iteration.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tools/config3_probe.py
"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200.datastore import ArrayDataStore  # noqa: E402
from merfish3d_analysis_b200.PixelDecoder import PixelDecoder  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tiles_per_rank = int(os.environ.get("TILES_PER_RANK", "4"))
    n_iter = int(os.environ.get("ITERATIONS", "3"))
    shape = (64, 2048, 2048)
    T = tiles_per_rank * world
    matrix = synthetic.mhd4_codebook_matrix(16)
    df_cb = synthetic.codebook_dataframe(matrix, n_blank=10)
    # ONE datastore shared by path (like the reference's workers, PD:249): rank 0 registers the tiles, every
    # rank then attaches the pixel data of the tiles it will read (in pinned host memory)
    root = Path("/tmp/m3d_config3_probe/qi2labdatastore")
    if rank == 0:
        import shutil

        shutil.rmtree(root.parent, ignore_errors=True)
        ds0 = ArrayDataStore(root, codebook=df_cb)
        for t in range(T):
            ds0.add_tile(np.zeros((16, 2, 8, 8), dtype=np.uint16))
    dist.barrier()
    ds = ArrayDataStore(root)
    mine = set(range(rank * tiles_per_rank, (rank + 1) * tiles_per_rank))
    seed_tiles = [0, 1]  # the percentile seed (rank 0 only) reads these
    hosts = []
    for t in range(T):
        if t in mine or (rank == 0 and t in seed_tiles):
            blk = synthetic.make_stack_device(matrix, shape, 3000 + t, device=dev)
            h = torch.empty(blk.shape, dtype=torch.uint16, pin_memory=True)
            h.copy_(blk)
            del blk
            hosts.append(h)
            for b, bit_id in enumerate(ds.bit_ids):
                ds._mem_readout[(ds.tile_ids[t], bit_id)] = h.numpy()[b]
                ds._mem_predictor[(ds.tile_ids[t], bit_id)] = None
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    dec = PixelDecoder(ds, merfish_bits=16, num_gpus=world, verbose=0)
    # the global seed is computed on rank 0 from `seed_tiles` and shared through the datastore attributes,
    # which here are per-rank stores: seed every rank's store with the same vectors instead
    if rank == 0:
        t0 = time.perf_counter()
        dec._load_global_normalization_vectors(gpu_id=local, recalculate=True, tile_indices=seed_tiles, lowpass_sigma=None)
        torch.cuda.synchronize()
        t_seed = time.perf_counter() - t0
        vec = [dec._global_normalization_vector, dec._global_background_vector]
    else:
        t_seed, vec = 0.0, None
    dist.barrier()

    # time the iterations: decode own tiles -> all_gather rows -> medians (identical on every rank)
    import types

    orig = dec._load_global_normalization_vectors

    def cached_global(self, gpu_id=0, recalculate=False, tile_indices=None, lowpass_sigma=None):
        return orig(gpu_id=gpu_id, recalculate=False, tile_indices=None, lowpass_sigma=lowpass_sigma)

    dec._load_global_normalization_vectors = types.MethodType(cached_global, dec)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    phases, calls = {}, {}
    if os.environ.get("C3_PHASES"):
        def timed(name, fn):
            def wrap(*a, **k):
                torch.cuda.synchronize()
                t_ = time.perf_counter()
                r = fn(*a, **k)
                torch.cuda.synchronize()
                phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t_) * 1e3
                calls.setdefault(name, []).append(round((time.perf_counter() - t_) * 1e3, 1))
                return r
            return wrap
        for name in ("_load_bit_data", "_decode_pixels", "_extract_barcodes", "_save_barcodes", "_gather_tables",
                     "_iterative_normalization_vectors", "_load_global_normalization_vectors"):
            setattr(dec, name, timed(name, getattr(dec, name)))
    prof = None
    if os.environ.get("C3_PROFILE") and rank == 0:
        import cProfile

        prof = cProfile.Profile()
        prof.enable()
    dec.optimize_normalization_by_decoding(n_iterations=n_iter, lowpass_sigma=None, magnitude_threshold=(1.5, 10.0),
                                           minimum_pixels=16.0, tile_indices=list(range(T)))
    if prof is not None:
        import pstats

        prof.disable()
        pstats.Stats(prof).sort_stats("cumulative").print_stats(45)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    if phases:
        print(f"[rank {rank}] phases ms over {n_iter} iterations x {tiles_per_rank} tiles:",
              {k: round(v, 1) for k, v in phases.items()}, flush=True)
        print(f"[rank {rank}] per call:", {k: calls[k] for k in ("_decode_pixels", "_gather_tables", "_extract_barcodes")
                                           if k in calls}, flush=True)
    nv = dec._iterative_normalization_vector
    all_nv = [None] * world
    dist.all_gather_object(all_nv, nv.tolist())
    if rank == 0:
        assert all(v == all_nv[0] for v in all_nv), "ranks disagree on the iterative vectors"
        n_vox = T * int(np.prod(shape))
        out = {"world": world, "tiles": T, "tile_shape": [16, *shape], "iterations": n_iter,
               "global_seed_s_rank0": t_seed, "optimiser_wall_s": wall,
               "s_per_iteration": wall / n_iter, "gvoxel_per_s": n_vox * n_iter / wall / 1e9,
               "iterative_normalization_vector_head": [float(v) for v in nv[:4]],
               "note": "tiles in pinned host memory; each iteration = H2D + decode + CCL + regionprops of every tile on "
                       "its rank, padded all_gather of the transcript rows over NCCL, per-bit medians on every rank"}
        Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"config3_probe_n{world}.json").write_text(json.dumps(out, indent=1))
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
