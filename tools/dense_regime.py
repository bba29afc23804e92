"""Dense-candidate regime of the search kernel: the optimiser's first iteration decodes with percentile-seeded
(noise-level) vectors, so EVERY voxel passes the magnitude gate.  Small volume so it can run under ncu."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200._capi import DecodeContext  # noqa: E402

import os

shape = (16, 1024, 1024) if len(sys.argv) < 4 else tuple(int(v) for v in sys.argv[1:4])
MODE = os.environ.get("M3D_PROBE_MODE", "seed")  # "seed": percentile-seeded vectors; "allfg": bench.py extras.all_foreground
matrix = synthetic.mhd4_codebook_matrix(16)
unit = (matrix / np.linalg.norm(matrix, axis=1, keepdims=True)).astype(np.float32)
ctx = DecodeContext(unit, (), device=0)
stack = synthetic.make_stack_device(matrix, shape, 3000, device=torch.device("cuda", 0))
# what _global_normalization_vectors yields on this value model: bkg ~ median of the lowest decile,
# nrm ~ median of the top decile above it
if MODE == "smooth":
    # the optimiser's first iteration as bench.py's extras.optimizer runs it: the stack low-passed (sigma 3, 1, 1) and decoded
    # with percentile-seeded vectors computed from the filtered data -> smooth traces, components of millions of voxels
    stack = ctx.lowpass(stack, (3.0, 1.0, 1.0), False)
    bk, nr = [], []
    for b in range(16):
        v = stack[b].flatten()[:: 17].float()
        p10 = torch.quantile(v, 0.10)
        bkg_b = v[v < p10].median()
        q = (v - bkg_b).clamp(min=0)
        p90 = torch.quantile(q, 0.90)
        bk.append(float(bkg_b))
        nr.append(float(q[q > p90].median()))
    ctx.set_normalization(np.asarray(bk, np.float32), np.asarray(nr, np.float32))
    ctx.set_thresholds(0.7653668647, 0.9, 10.0)
elif MODE == "allfg":  # unsaturated traces, magnitude gate open: few exact ties, most voxels settled by the top-w lookup
    ctx.set_normalization(np.full(16, 0.0, np.float32), np.full(16, 250.0, np.float32))
    ctx.set_thresholds(0.7653668647, 1.0e-3, 10.0)
else:
    ctx.set_normalization(np.full(16, 187.0, np.float32), np.full(16, 17.0, np.float32))
    ctx.set_thresholds(0.7653668647, 1.5, 10.0)
decoded = torch.empty(shape, dtype=torch.int16, device="cuda")
ctx.set_timing(True)
for i in range(3):
    if i == 1:
        ctx.reset_counters()
    n = ctx.decode_label(stack, decoded, False, 16.0, 500)
torch.cuda.synchronize()
kt = ctx.kernel_times_ms()
nv = int(np.prod(shape))
print({k: round(v / 2, 3) for k, v in kt.items()}, "features", n, "decoded fraction", float((decoded >= 0).float().mean()),
      "ns/voxel", round(sum(kt.values()) / 2 / nv * 1e6, 3))
