"""Own zstd decoder (csrc/zstd_decode.cuh) vs the system libzstd on the host: equality and single-thread speed on the
kinds of data a Blosc block holds.  CPU only.  Usage: python tools/zstd_builtin_check.py"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))

from merfish3d_analysis_b200 import _capi  # noqa: E402
from test_cpu_zarr_store import _zstd_datasets  # noqa: E402

for name, data in _zstd_datasets():
    if len(data) < 1000:
        continue
    for level in (3, 9):
        frame = _capi.zstd_host(data, True, level=level)
        t0 = time.perf_counter()
        mine = _capi.zstd_decode_builtin(frame, len(data))
        t1 = time.perf_counter()
        ref = _capi.zstd_host(frame, False, len(data))
        t2 = time.perf_counter()
        print(f"{name:18s} level {level}: raw {len(data):8d} B, frame {len(frame):8d} B, own decoder "
              f"{len(data) / (t1 - t0) / 1e6:7.1f} MB/s, libzstd {len(data) / (t2 - t1) / 1e6:7.1f} MB/s, equal={mine == ref == data}")
