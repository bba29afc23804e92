"""Config-4 probe (22-bit HW4 codebook, ~1000 words, 2-D mode): production path timing at
K = 140 / 385 / 1000 to decide whether the voxel x bits x codeword contraction needs tensor cores."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200._capi import DecodeContext  # noqa: E402

shape = (32, 2048, 2048)
n_vox = int(np.prod(shape))
out = {}
for n_bits, K in ((22, 140), (22, 385), (22, 1000)):
    m = synthetic.random_hw4_codebook_matrix(n_bits, K, 4004)
    unit = (m / np.linalg.norm(m, axis=1, keepdims=True)).astype(np.float32)
    ctx = DecodeContext(unit, (), device=0)
    ctx.set_normalization(np.full(n_bits, 200.0, np.float32), np.full(n_bits, 900.0, np.float32))
    ctx.set_thresholds(0.7653668647, 1.5, 10.0)
    stack = synthetic.make_stack_device(m, shape, 4004, device="cuda")
    dec = torch.empty(shape, dtype=torch.int16, device="cuda")
    ctx.set_timing(True)
    for i in range(5):
        if i == 2:
            torch.cuda.synchronize()
            ctx.reset_counters()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        n = ctx.decode_label(stack, dec, True, 7.0, 500)
        tab = ctx.features(stack, dec, False, n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    kt = {k: round(v / 3, 4) for k, v in sorted(ctx.kernel_times_ms().items(), key=lambda kv: -kv[1])[:4]}
    # worst case for the search: every voxel a candidate
    ctx.set_thresholds(0.7653668647, 1.0e-3, 10.0)
    sub = stack[:, :4].contiguous()
    dsub = torch.empty(sub.shape[1:], dtype=torch.int16, device="cuda")
    ctx.decode(sub, dsub)
    torch.cuda.synchronize()
    ctx.reset_counters()
    ctx.decode(sub, dsub)
    torch.cuda.synchronize()
    wc = ctx.kernel_times_ms().get("decode_search_kernel", 0.0)
    out[f"bits{n_bits}_K{K}"] = dict(ms_per_step=round(ms, 3), gvoxel_s=round(n_vox / ms / 1e6, 1), features=int(n),
                                    fg=int((dec >= 0).sum()), kernels_ms=kt,
                                    all_candidates_search_ns_per_voxel=round(wc * 1e6 / sub[0].numel(), 3))
    ctx.close()
    del stack, dec
print(json.dumps(out, indent=1))
