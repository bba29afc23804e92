"""Multi-rank smoke (run under torchrun, NCCL): tile-sharded optimiser loop + decode_all_tiles,
checked against the CPU oracle on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_smoke.py
"""
import os
import shutil
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
from merfish3d_analysis_b200.datastore import ArrayDataStore  # noqa: E402
from merfish3d_analysis_b200.PixelDecoder import PixelDecoder  # noqa: E402
from oracle import decode_oracle as orc  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    root = Path("/tmp/m3d_dist_smoke/qi2labdatastore")
    df_cb, cb = cases.codebook16()
    n_tiles = 5
    stacks = [cases.small_stack(cb["matrix"], shape=(10, 48, 64), seed=71 + i, density=4e-3) for i in range(n_tiles)]
    if rank == 0:
        shutil.rmtree(root.parent, ignore_errors=True)
        ds = ArrayDataStore(root, codebook=df_cb)
        for st in stacks:
            ds.add_tile(st, persist=True)
    dist.barrier()
    ds = ArrayDataStore(root)
    dec = PixelDecoder(ds, merfish_bits=16, num_gpus=world, verbose=0)
    sigma = (1.0, 0.5, 0.5)
    dec.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=sigma,
                                           magnitude_threshold=(0.9, 10.0), tile_indices=list(range(n_tiles)))
    dec.decode_all_tiles(lowpass_sigma=sigma, minimum_pixels=4, magnitude_threshold=(0.9, 10.0))
    n_local = len(dec._df_barcodes_loaded)
    counts = [None] * world
    dist.all_gather_object(counts, n_local)
    if rank == 0:
        ref = orc.optimize_normalization([(s, None) for s in stacks], cb, 3, True, sigma, (0.9, 10.0), 4)
        i_n, i_b = ds.load_decode_normalization_vectors(None, "iterative")
        np.testing.assert_array_equal(i_n, ref["iterative"][0])
        np.testing.assert_array_equal(i_b, ref["iterative"][1])
        total = 0
        for t, st in enumerate(stacks):
            df, _ = orc.decode_tile(st, None, cb, i_b, i_n, True, sigma, (0.9, 10.0), 4, tile_idx=t,
                                    spacing=ds.voxel_size_zyx_um)
            got = ds.load_local_decoded_spots(t)
            assert got["gene_id"].tolist() == df["gene_id"].tolist(), t
            np.testing.assert_allclose(got["distance_min"].to_numpy(float), df["distance_min"].to_numpy(float), rtol=1e-5)
            total += len(df)
        assert all(c == total for c in counts), (counts, total)
        print(f"dist_smoke ok: world={world} tiles={n_tiles} transcripts={total} vectors match the oracle")
    dist.barrier()
    # ---- z-slab sharding of ONE tile across the ranks == unsharded decode
    big = cases.small_stack(cb["matrix"], shape=(8 * world + 3, 48, 64), seed=97, density=4e-3)
    on = np.flatnonzero(cb["matrix"][7])
    big[:, :, 2:8, 2:8] = 210
    big[np.ix_(on, np.arange(big.shape[1]), np.arange(2, 8), np.arange(2, 8))] = 2000  # oversized, crosses all slabs
    on2 = np.flatnonzero(cb["matrix"][21])
    big[:, 2:-2, 30, 40] = 205
    big[np.ix_(on2, np.arange(2, big.shape[1] - 2), [30], [40])] = 1900  # only large enough once merged
    if rank == 0:
        ds0 = ArrayDataStore(root)
        ds0.add_tile(big, persist=True)
    dist.barrier()
    ds = ArrayDataStore(root)
    t_big = len(ds.tile_ids) - 1
    for sig, per_rank in ((None, 1), ((3.0, 1.0, 1.0), 1), (None, 3), ((3.0, 1.0, 1.0), 2)):
        d = PixelDecoder(ds, merfish_bits=16, num_gpus=world, verbose=0)
        d.decode_one_tile_sharded(t_big, lowpass_sigma=sig, minimum_pixels=12, slabs_per_rank=per_rank)
        slab = d.decoded_image
        u = PixelDecoder(ds, merfish_bits=16, num_gpus=1, verbose=0)
        u.decode_one_tile(t_big, gpu_id=local, lowpass_sigma=sig, minimum_pixels=12)
        z0, z1 = d._slab_bounds[0][0], d._slab_bounds[-1][1]
        assert len(d._slab_bounds) == per_rank
        np.testing.assert_array_equal(slab, u.decoded_image[z0:z1])  # every rank: its planes of the decoded image
        if rank == 0:
            import pandas as pd

            pd.testing.assert_frame_equal(d.decoded_barcodes, u.decoded_barcodes)
            print(f"dist_smoke ok: z-slab sharded x{world}, {per_rank} slab(s) per rank, lowpass={sig}: "
                  f"{len(u.decoded_barcodes)} transcripts identical")
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
