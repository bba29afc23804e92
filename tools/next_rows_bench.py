"""Timings of the SURVEY 8f "next" rows on one B200 (written to gpurun_out/next_rows.json):
decode-time affine warp (8f-1), per-on-bit centroid statistics (8f-4) at configs[1] size, and the
post-decode table stage (8f-3) at 2e6 transcripts, with the CPU oracle timed on a bounded sample."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from merfish3d_analysis_b200 import synthetic  # noqa: E402
from merfish3d_analysis_b200._capi import DecodeContext  # noqa: E402
from oracle import table_oracle as tor  # noqa: E402  (CPU comparison only)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    out = {}
    dev = torch.device("cuda", 0)
    matrix = synthetic.mhd4_codebook_matrix(16)
    unit = (matrix / np.linalg.norm(matrix, axis=1, keepdims=True)).astype(np.float32)
    ctx = DecodeContext(unit, (), device=0)
    shape = (100, 2048, 2048)
    n_vox = int(np.prod(shape))

    # ---- 8f-1: affine warp of one uint16 bit volume (order 1, float64 taps)
    vol = torch.randint(0, 4000, shape, dtype=torch.int32, device=dev).to(torch.uint16)
    m = np.eye(3) + 1e-3 * np.array([[0, 1, -1], [1, 0, 2], [-2, 1, 0]], dtype=np.float64)
    off = np.array([0.3, -1.7, 2.2])
    outv = torch.empty(shape, dtype=torch.float32, device=dev)
    ms = timed(lambda: ctx.warp_affine(vol, m, off, out=outv))
    out["warp_affine"] = {"ms_per_bit_volume": ms, "gvoxel_per_s": n_vox / ms / 1e6,
                          "hbm_gb_s_algorithmic": n_vox * 6 / ms / 1e6, "note": "uint16 in (2 B) + float32 out (4 B) per voxel"}
    del vol, outv

    # ---- 8f-4: centroid statistics, all 16 bits in one pass, ~28k components
    stack = synthetic.make_stack_device(matrix, shape, 2002, device=dev)
    ctx.set_normalization(np.full(16, 200.0, np.float32), np.full(16, 900.0, np.float32))
    ctx.set_thresholds(0.7653668647, 1.5, 10.0)
    decoded = torch.empty(shape, dtype=torch.int16, device=dev)
    labels = torch.empty(shape, dtype=torch.int32, device=dev)
    n = ctx.decode_label(stack, decoded, False, 16.0, 500, labels=labels)
    table = ctx.features(stack, decoded, True, n)
    code = torch.full((n + 1,), -1, dtype=torch.int16, device=dev)
    code[1:] = table[:, 2].to(torch.int16)
    ms = timed(lambda: ctx.centroid_statistics(labels, stack, 7, code))
    out["centroid_statistics"] = {"ms": ms, "components": int(n), "z_support": 7,
                                  "hbm_gb_s_algorithmic": n_vox * 4 / ms / 1e6,
                                  "note": "one pass for all bits; algorithmic traffic = the int32 label image (4 B/voxel); "
                                          "the reference makes 16 full-volume bincount passes"}
    ms = timed(lambda: ctx.inertia_eigvals(table))
    out["inertia_eigvals"] = {"ms": ms, "rows": int(n)}
    del stack, decoded, labels

    # ---- 8f-3: table stage at 2e6 transcripts
    rng = np.random.default_rng(1)
    N = 2_000_000
    zyx = np.round(np.column_stack([rng.uniform(0, 30, N), rng.uniform(0, 3000, N), rng.uniform(0, 3000, N)]), 2)
    tile = ((zyx[:, 1] // 190).astype(np.int32) * 16 + (zyx[:, 2] // 190).astype(np.int32))
    # a fifth of the rows are re-detections by a neighbouring tile
    dup = rng.choice(N, N // 5, replace=False)
    zyx[dup] = np.round(zyx[(dup + 1) % N] + rng.uniform(-0.3, 0.3, (dup.size, 3)), 2)
    dmin = rng.uniform(0.05, 0.6, N)
    gene = rng.integers(0, 140, N).astype(np.int32)
    d_zyx, d_tile, d_dmin, d_gene = (torch.from_numpy(a).to(dev) for a in (zyx, tile, dmin, gene))
    ms_overlap = timed(lambda: ctx.overlap_duplicates(d_zyx, d_tile, d_dmin, 0.75))
    drop = ctx.overlap_duplicates(d_zyx, d_tile, d_dmin, 0.75).cpu().numpy().astype(bool)
    ms_within = timed(lambda: ctx.within_tile_duplicates(d_zyx, d_tile, d_gene, d_dmin, 0.1085, 1.5))
    v = [torch.from_numpy(rng.uniform(0, 1, N).astype(np.float32)).to(dev) for _ in range(3)]
    blank = torch.from_numpy((rng.uniform(0, 1, N) < 0.05).astype(np.uint8)).to(dev)
    edges = [np.linspace(0, 1.0000001, 11, dtype=np.float32)] * 3
    ms_hist = timed(lambda: ctx.table_hist3d(v[0], v[1], v[2], blank, *edges))
    # CPU oracle on a bounded sample (first 200k rows), same semantics
    S = 200_000
    t0 = time.perf_counter()
    ref = tor.overlap_duplicates(zyx[:S], tile[:S], dmin[:S], 0.75)
    cpu_s = time.perf_counter() - t0
    got = ctx.overlap_duplicates(d_zyx[:S].contiguous(), d_tile[:S].contiguous(), d_dmin[:S].contiguous(), 0.75)
    assert np.array_equal(got.cpu().numpy().astype(bool), ref)
    out["table_stage"] = {
        "rows": N, "overlap_duplicates_ms": ms_overlap, "overlap_dropped": int(drop.sum()),
        "within_tile_duplicates_ms": ms_within, "hist3d_ms": ms_hist,
        "rows_per_s_overlap": N / ms_overlap * 1e3,
        "cpu_oracle_overlap": {"rows": S, "seconds": cpu_s, "rows_per_s": S / cpu_s,
                               "note": "cKDTree.query_pairs + Python loop (the reference's algorithm), 1 core; "
                                       "device result on the same rows is identical"},
    }
    ctx.close()
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "next_rows.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
