"""Time the device zstd decoders (M3D_ZARR_GPU_ZSTD=1|2, argv[1]) on ONE chunk of (argv[2] or 16, 512, 512) uint16, ring warm:
one launch, 2 x planes Blosc blocks = warps.  A 256-plane chunk is one launch as wide as 16 ordinary chunks in flight."""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

os.environ["M3D_ZARR_GPU_ZSTD"] = sys.argv[1] if len(sys.argv) > 1 else "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200 import zarr_store as zs  # noqa: E402
from merfish3d_analysis_b200._capi import DecodeContext  # noqa: E402

rng = np.random.default_rng(0)
planes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
a = (rng.poisson(100, (planes, 512, 512)) + 100).astype(np.uint16)
m = synthetic.mhd4_codebook_matrix(16).astype(np.float32)
ctx = DecodeContext(m / np.linalg.norm(m, axis=1, keepdims=True))
with tempfile.TemporaryDirectory() as t:
    zs.write_ome_image(t + "/img", a, chunks=(planes, 512, 512))
    img = zs.ZarrImage(t + "/img.ome.zarr")
    dst = torch.zeros(a.shape, dtype=torch.uint16, device=ctx.device)
    zs.transfer(ctx, [(img, dst)])
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy(), a)
    ctx.set_timing(True)
    ctx.reset_counters()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        zs.transfer(ctx, [(img, dst)])
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("mode", os.environ["M3D_ZARR_GPU_ZSTD"], f"planes {planes} ({a.nbytes / 1e6:.1f} MB) wall ms per chunk:", [round(v, 2) for v in ts], "kernel ms total (3 launches):", ctx.kernel_times_ms())
