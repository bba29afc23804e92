"""Low-pass kernels alone on one bit volume (argv: z y x, default 100 2048 2048; argv[4] = float32 for the opt-in mode):
prints the per-kernel times; run under ncu for the pipe counters (profiles/r2_lowpass_*.txt)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200._capi import DecodeContext  # noqa: E402

z, y, x = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (100, 2048, 2048)
m = synthetic.mhd4_codebook_matrix(16).astype(np.float32)
ctx = DecodeContext(m / np.linalg.norm(m, axis=1, keepdims=True))
if len(sys.argv) > 4:
    ctx.set_lowpass_accumulate(sys.argv[4])
g = torch.Generator(device=ctx.device)
g.manual_seed(1)
vol = (torch.poisson(torch.full((1, z, y, x), 100.0, device=ctx.device), generator=g) + 100).to(torch.uint16)
out = None
ctx.set_timing(True)
for i in range(3):
    if i == 1:
        torch.cuda.synchronize()
        ctx.reset_counters()
    out = ctx.lowpass(vol, (3.0, 1.0, 1.0), False, out=out)
torch.cuda.synchronize()
print({k: v / 2 for k, v in ctx.kernel_times_ms().items()}, "ms per volume", z * y * x / 1e6, "Mvox")
