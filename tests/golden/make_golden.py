"""Generates tests/golden/decode_small.npz.

The reference module cannot be imported in the build image (CuPy / cuCIM / cuVS / scikit-image
absent) and upstream ships no numeric goldens for this path, so this fixture
pins the ORACLE's output for a fixed seed (drift detector); the upstream known-answer vectors
that do exist are restated directly in tests/test_cpu_oracle_and_host.py.
Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
from oracle import decode_oracle as orc  # noqa: E402

_df, cb = cases.codebook16()
stack = cases.small_stack(cb["matrix"], shape=(6, 24, 32), seed=101, density=4e-3)
bkg, nrm = cases.simple_vectors(16)
unit = orc.normalize_codebook(cb["matrix"])
out = orc.decode_pixels(stack.astype(np.float32), unit, bkg, nrm, cb["pixel_assignment_threshold"], (1.5, 10.0))
labels = cases.canonical_labels(orc.filter_label_sizes(orc.label_decoded(out["decoded"], True), 4))
np.savez_compressed(
    ROOT / "tests" / "golden" / "decode_small.npz", stack=stack, bkg=bkg, nrm=nrm, decoded=out["decoded"],
    magnitude=out["magnitude"], distance=out["distance"], labels=labels,
)
print("features:", labels.max(), "foreground:", int((out["decoded"] >= 0).sum()))
