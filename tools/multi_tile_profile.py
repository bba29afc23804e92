"""cProfile of decode_all_tiles over 3 tiles of configs[1] (same pinned array registered three times)."""
import cProfile
import pstats
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200.datastore import ArrayDataStore  # noqa: E402
from merfish3d_analysis_b200.PixelDecoder import PixelDecoder  # noqa: E402

shape = (100, 2048, 2048)
dev = torch.device("cuda", 0)
matrix = synthetic.mhd4_codebook_matrix(16)
df_cb = synthetic.codebook_dataframe(matrix, n_blank=10)
stack = synthetic.make_stack_device(matrix, shape, 2002, device=dev)
host = torch.empty(stack.shape, dtype=torch.uint16, pin_memory=True)
host.copy_(stack)
torch.cuda.synchronize()
del stack
torch.cuda.empty_cache()
tmp = tempfile.TemporaryDirectory()
ds = ArrayDataStore(Path(tmp.name) / "qi2labdatastore", codebook=df_cb)
for k in range(3):
    ds.add_tile(host.numpy(), stage_origin_zyx_um=(0.0, 0.0, 250.0 * k))
ds.save_decode_normalization_vectors(None, "global", np.full(16, 900.0, np.float32), np.full(16, 200.0, np.float32))
dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
kw = dict(assign_to_cells=False, lowpass_sigma=None, magnitude_threshold=(1.5, 10.0), minimum_pixels=16.0,
          normalization_method="global")
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec.decode_all_tiles(**kw)
    torch.cuda.synchronize()
    print(f"decode_all_tiles x3: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
pr = cProfile.Profile()
pr.enable()
dec.decode_all_tiles(**kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
