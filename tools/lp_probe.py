"""Low-pass kernel probe for ncu: one bit volume through m3d_lowpass."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from merfish3d_analysis_b200._capi import DecodeContext
unit = np.eye(4, 16, dtype=np.float32)
ctx = DecodeContext(unit, (), device=0)
Z, Y, X = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (100, 2048, 2048)
vol = torch.randint(0, 4000, (2, Z, Y, X), device="cuda", dtype=torch.int32).to(torch.int16).view(torch.uint16)
out = None
ctx.set_timing(True)
for i in range(3):
    if i == 1:
        ctx.reset_counters()
    out = ctx.lowpass(vol, (3.0, 1.0, 1.0), False, out=out)
torch.cuda.synchronize()
print({k: v / 4 for k, v in ctx.kernel_times_ms().items()}, "ms per volume", Z * Y * X / 1e6, "Mvox")
