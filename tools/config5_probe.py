"""Z-slab sharding of ONE volume at realistic size (configs[4] geometry scaled to what the host RAM of a
2-GPU box holds): 16 bits x (100 * world) z x 2048 x 2048 uint16, one z range per rank over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29521 tools/config5_probe.py

Every rank times ``PixelDecoder.decode_one_tile_sharded`` (host-resident volume -> transcripts on rank 0);
rank 0 then decodes the same volume unsharded and checks that the tables are identical."""
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
    per = int(os.environ.get("PLANES_PER_RANK", "100"))
    Y = X = int(os.environ.get("YX", "2048"))
    slabs_per_rank = int(os.environ.get("SLABS_PER_RANK", "1"))
    matrix = synthetic.mhd4_codebook_matrix(16)
    df_cb = synthetic.codebook_dataframe(matrix, n_blank=10)
    Z = per * world
    host = torch.empty((16, Z, Y, X), dtype=torch.uint16, pin_memory=True)
    for part in range(world):  # the same volume on every rank (each rank only READS its own z range + the probe)
        blk = synthetic.make_stack_device(matrix, (per, Y, X), 5005 + part, device=dev)
        host[:, part * per:(part + 1) * per].copy_(blk)
        del blk
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    tmp = tempfile.TemporaryDirectory()
    ds = ArrayDataStore(Path(tmp.name) / f"store_r{rank}", codebook=df_cb)
    ds.add_tile(host.numpy())
    ds.save_decode_normalization_vectors(None, "global", np.full(16, 900.0, np.float32), np.full(16, 200.0, np.float32))
    dec = PixelDecoder(ds, merfish_bits=16, num_gpus=world, verbose=0)
    times = []
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec.decode_one_tile_sharded(0, lowpass_sigma=None, normalization_method="global", slabs_per_rank=slabs_per_rank)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t0)
    n_vox = Z * Y * X
    if rank == 0:
        got = dec.decoded_barcodes
        t0 = time.perf_counter()
        uns = PixelDecoder(ds, merfish_bits=16, num_gpus=1, verbose=0)
        uns.decode_one_tile(0, gpu_id=local, lowpass_sigma=None, normalization_method="global")
        torch.cuda.synchronize()
        t_uns = time.perf_counter() - t0
        ref = uns.decoded_barcodes
        import pandas as pd

        pd.testing.assert_frame_equal(got, ref)
        out = {"world": world, "shape": [16, Z, Y, X], "slabs_per_rank": slabs_per_rank, "transcripts": int(len(got)),
               "sharded_s": times, "sharded_gvoxel_per_s": n_vox / min(times) / 1e9,
               "unsharded_one_gpu_s": t_uns, "identical_to_unsharded": True,
               "note": "wall time of decode_one_tile_sharded incl. H2D of each rank's planes from pinned host memory, "
                       "boundary-plane send/recv, all_gather of equivalences and assembly on rank 0"}
        Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"config5_probe_n{world}.json").write_text(json.dumps(out, indent=1))
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
