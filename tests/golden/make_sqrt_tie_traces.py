"""Search for unit traces on which two codewords have DIFFERENT float32 direct-form sums that round to the SAME distance, the
lower sum belonging to the HIGHER codeword index: NumPy's first arg-min (PD:2512) must keep the lower index.  Writes
tests/golden/sqrt_tie_traces.npy (200 x 16 float32), used by tests/test_gpu_kernels.py::test_equal_distance_from_smaller_sum."""
import sys, numpy as np
ROOT = __import__("pathlib").Path(__file__).resolve().parents[2]; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cases
from oracle import decode_oracle as orc
F = np.float32
_df, cb = cases.codebook16()
unit = orc.normalize_codebook(cb["matrix"][:, :16]).astype(F)
m = cb["matrix"][:, :16].astype(bool)
rng = np.random.default_rng(123)
found = []
for trial in range(60):
    n = 200000
    # traces with 5-7 nearly equal large entries (ulps apart) and small others
    base = rng.uniform(0.3, 0.9, (n, 1)).astype(F)
    x = (rng.uniform(0.0, 0.05, (n, 16))).astype(F)
    nbig = rng.integers(5, 8, n)
    for i in range(7):
        idx = rng.integers(0, 16, n)
        sel = i < nbig
        pert = (rng.integers(-3, 4, n).astype(F) * np.spacing(base[:, 0]))
        x[np.arange(n)[sel], idx[sel]] = (base[:, 0] + pert)[sel]
    nn = np.sqrt((x * x).sum(1, dtype=F)).astype(F)  # not the exact sequential order, but we re-evaluate below via the oracle
    # exact pipeline through the oracle's functions on these as a (16, 1, 1, n) stack
    st = np.ascontiguousarray(x.T.reshape(16, 1, 1, n))
    xs = orc.scale_traces(st.reshape(16, -1), np.zeros(16, F), np.ones(16, F))
    xh, mag = orc.normalize_traces(np.clip(xs, 0, 1).astype(F))
    xh = xh.T.astype(F)  # (n, 16)
    # sequential float32 direct-form sums for all codewords
    acc = np.zeros((n, unit.shape[0]), dtype=F)
    for b in range(16):
        t = (xh[:, b:b+1] - unit[None, :, b]).astype(F)
        term = (t * t).astype(F)
        acc = term if b == 0 else (acc + term).astype(F)
    d = np.sqrt(acc).astype(F)
    k_first = d.argmin(1)
    dmin = d.min(1)
    amin = acc.min(1)
    k_acc = acc.argmin(1)
    # interesting: lowest-acc codeword has a HIGHER index than the first argmin of d (equal sqrt, different sums)
    hit = np.flatnonzero((k_acc != k_first) & (k_acc > k_first) & (acc[np.arange(n), k_first] > amin))
    for h in hit[:50]:
        found.append(x[h].copy())
    if len(found) >= 200: break
print(len(found))
np.save(str(__import__("pathlib").Path(__file__).resolve().parent / "sqrt_tie_traces.npy"), np.asarray(found[:200], dtype=F))
