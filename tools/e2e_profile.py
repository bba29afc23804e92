"""Where does the e2e step (host-resident stack -> transcripts table) spend its time?

Times the phases of ``PixelDecoder.decode_one_tile`` on configs[1] with a device synchronise
between them (so this is a breakdown, not a bench number)."""
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

shape = (100, 2048, 2048) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1:4])
import os
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(LOCAL)
dev = torch.device("cuda", LOCAL)
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
ds.add_tile(host.numpy())
ds.save_decode_normalization_vectors(None, "global", np.full(16, 900.0, np.float32), np.full(16, 200.0, np.float32))
dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
GPU = LOCAL

phases = {}


def timed(name, fn):
    def wrap(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    return wrap


for name in ("_prepare_normalization_state", "_load_bit_data", "_decode_pixels", "_extract_barcodes", "_annotate_table"):
    setattr(dec, name, timed(name, getattr(dec, name)))

for it in range(4):
    phases.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec.decode_one_tile(0, gpu_id=GPU, lowpass_sigma=None, magnitude_threshold=(1.5, 10.0), minimum_pixels=16.0,
                        normalization_method="global")
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
    print(f"[rank {LOCAL} omp={os.environ.get('OMP_NUM_THREADS')}] iter {it}: total {total:.1f} ms  " + "  ".join(f"{k}={v:.1f}" for k, v in phases.items()),
          f" rows={len(dec._df_barcodes)}", flush=True)

if os.environ.get("E2E_CPROFILE"):
    import cProfile  # noqa: E402
    import pstats  # noqa: E402
    
    pr = cProfile.Profile()
    pr.enable()
    dec.decode_one_tile(0, gpu_id=GPU, lowpass_sigma=None, magnitude_threshold=(1.5, 10.0), minimum_pixels=16.0,
                        normalization_method="global")
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
