"""Time one digit-histogram pass of the radix select (m3d_select_hist) over a float32 volume of one filtered bit image
(16 x 64 x 2048 x 2048 tile -> 268 M elements, 1.07 GB): first digit (no prefix), first digit under a predicate, second
and third digit (prefix set).  HBM time of a pass = bytes / measured copy bandwidth."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200._capi import DecodeContext  # noqa: E402

matrix = synthetic.mhd4_codebook_matrix(16)
unit = (matrix / np.linalg.norm(matrix, axis=1, keepdims=True)).astype(np.float32)
ctx = DecodeContext(unit, (), device=0)
n = 64 * 2048 * 2048
g = torch.Generator(device="cuda").manual_seed(5)
vol = (200.0 + 30.0 * torch.randn(n, device="cuda", generator=g)).clamp_(min=0).float()  # background-like values
vol[:: 997] += 900.0  # sparse bright spots
hist = torch.zeros(2048, dtype=torch.int64, device="cuda")


def timed(label, **kw):
    for _ in range(2):
        ctx.select_hist(vol, hist, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ctx.select_hist(vol, hist, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label:44s} {ms:7.3f} ms  {n * 4 / ms / 1e6:7.1f} GB/s")


key200 = int(np.array([200.0], np.float32).view(np.uint32)[0]) | 0x80000000
base = dict(sub=0.0, clip0=False, pred=0, cutoff=0.0, prefix_mask=0, prefix_value=0, shift=21)
timed("first digit, all elements", **base)
timed("first digit, v < cutoff (10 % pass)", **{**base, "pred": 1, "cutoff": 161.5})
timed("first digit, clip(v - bkg) > cutoff", **{**base, "pred": 2, "cutoff": 40.0, "sub": 180.0, "clip0": True})
timed("second digit, prefix = bin of 200", **{**base, "prefix_mask": 0xFFE00000, "prefix_value": key200 & 0xFFE00000, "shift": 10})
timed("third digit, prefix = 22 bits of 200", **{**base, "prefix_mask": 0xFFFFFC00, "prefix_value": key200 & 0xFFFFFC00, "shift": 0})
