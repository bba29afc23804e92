"""Host-side profile of one optimiser-iteration-0-like decode_one_tile call (dense-candidate regime: noise-level
vectors, hundreds of thousands of components per tile): where the ~1 s per 16x64x2048x2048 tile goes."""
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

shape = (64, 2048, 2048)
dev = torch.device("cuda", 0)
matrix = synthetic.mhd4_codebook_matrix(16)
df_cb = synthetic.codebook_dataframe(matrix, n_blank=10)
blk = synthetic.make_stack_device(matrix, shape, 3000, device=dev)
host = torch.empty(blk.shape, dtype=torch.uint16, pin_memory=True)
host.copy_(blk)
del blk
torch.cuda.empty_cache()
tmp = tempfile.TemporaryDirectory()
ds = ArrayDataStore(Path(tmp.name) / "qi2labdatastore", codebook=df_cb)
ds.add_tile(host.numpy())
dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
kw = dict(n_iterations=1, lowpass_sigma=(3.0, 1.0, 1.0), magnitude_threshold=(1.5, 10.0), minimum_pixels=16.0, tile_indices=[0])
dec.optimize_normalization_by_decoding(**kw)  # warm-up (allocations)
ctx = dec._ctx(0)
ctx.set_timing(True)
ctx.reset_counters()
pr = cProfile.Profile()
torch.cuda.synchronize()
t0 = time.perf_counter()
pr.enable()
dec.optimize_normalization_by_decoding(**kw)
torch.cuda.synchronize()
pr.disable()
print("one optimiser iteration 0 (1 tile, incl. the percentile seed):", round(time.perf_counter() - t0, 3), "s")
print(dec._optimizer_timing)
print({k: round(v, 2) for k, v in sorted(ctx.kernel_times_ms().items(), key=lambda kv: -kv[1])[:12]})
pstats.Stats(pr).sort_stats("cumtime").print_stats(45)
