"""Seeded scenarios behind tests/golden/reference_*.npz (shared by the generator and the tests)."""
from __future__ import annotations

import numpy as np

import cases

# name -> scenario.  Everything a test needs to rebuild the inputs is stored in the fixture.
SCENARIOS = {
    # 3-D, raw (no low-pass), global vectors, stage origin -> global coordinates
    "raw3d": dict(shape=(8, 40, 48), seed=11, density=4e-3, lowpass=None, norm="global", min_px=4,
                  origin=(10.0, 200.0, -300.0)),
    # 3-D, reference-default low-pass, non-unit predictor weights
    "lp3d": dict(shape=(10, 40, 48), seed=12, density=4e-3, lowpass=(3.0, 1.0, 1.0), norm="global", min_px=6,
                 predictor=True, bkg=60.0, nrm=300.0, mag=(0.9, 10.0)),
    # 2-D decode mode (per-plane 8-connectivity, default 7 px), low-pass sigma (3,1,1) -> (1,1)
    "mode2d": dict(shape=(5, 48, 48), seed=13, density=6e-3, lowpass=(3.0, 1.0, 1.0), norm="global", min_px=None,
                   microscope="2D", bkg=150.0, nrm=500.0, mag=(0.9, 10.0)),
    # exclusions + z_range crop + unnormalised traces
    # optimiser mode with per-on-bit weighted centroids (raw float32 intensities, z-crop offset applied)
    "chroma": dict(shape=(12, 40, 48), seed=15, density=5e-3, lowpass=(1.0, 0.5, 0.5), norm="global", min_px=4,
                   z_range=(1, 11), chroma=True, bkg=150.0, nrm=500.0, mag=(0.9, 10.0)),
    # configs[3] flavour: 22 bits, 300 random weight-4 codewords, 2-D decode mode
    "bits22": dict(shape=(4, 40, 48), seed=16, density=8e-3, lowpass=None, norm="global", min_px=4, bits=22,
                   microscope="2D"),
    # configs[3] at the sizes BASELINE.json / SURVEY 8d name: K = 385 (the most a 22-bit weight-4 distance-4 code can
    # hold) and K ~ 1000 weight-4 words -- sparse candidates (production vectors) and dense candidates (noise-level
    # vectors: every voxel passes the magnitude gate, exact ties among codewords are common).  K > 256 takes the
    # wide candidate sets / larger hash of the search kernel.
    "bits22_k385": dict(shape=(4, 40, 48), seed=18, density=8e-3, lowpass=None, norm="global", min_px=4, bits=22,
                        words=385, microscope="2D"),
    "bits22_k1000": dict(shape=(4, 40, 48), seed=19, density=8e-3, lowpass=None, norm="global", min_px=4, bits=22,
                         words=1000, microscope="2D"),
    "bits22_k385_dense": dict(shape=(3, 40, 48), seed=20, density=1.5e-3, lowpass=None, norm="global", min_px=3, bits=22,
                              words=385, microscope="2D", bkg=195.0, nrm=30.0, nrm_jitter=3.0, mag=(0.05, 10.0)),
    "bits22_k1000_dense": dict(shape=(3, 40, 48), seed=21, density=1.5e-3, lowpass=None, norm="global", min_px=3, bits=22,
                               words=1000, microscope="2D", bkg=195.0, nrm=30.0, nrm_jitter=3.0, mag=(0.05, 10.0)),
    # dense candidates with the 16-bit MHD4 code in 3-D (the optimiser's first iteration looks like this: several
    # bits of every trace clip to exactly 0 or 1, so exact distance ties between codewords are the rule)
    "dense16": dict(shape=(6, 40, 48), seed=22, density=1.5e-3, lowpass=None, norm="global", min_px=4,
                    bkg=195.0, nrm=30.0, nrm_jitter=3.0, mag=(0.05, 10.0)),
    # decode-time warp: bits imaged in later rounds carry an affine round transform; rounds 2 and 3 also a SOFIMA
    # flow field (round 4's is an identity fallback -> affine only); z crop on top
    "warp": dict(shape=(8, 40, 48), seed=17, density=6e-3, lowpass=None, norm="global", min_px=4, warp=True,
                 z_range=(1, 7)),
    "excl_crop": dict(shape=(9, 32, 40), seed=14, density=6e-3, lowpass=None, norm="global", min_px=3,
                      z_range=(2, 8), exclude=3),
    # medium volumes (40x the voxels of the others, ~2000 transcripts / 330 k searched voxels): more rare events --
    # near-ties, touching components, size-filter edge cases.  ``slim``: the fixture stores a checksum of the seeded
    # input instead of the input, and no float32 copy / scaled images (the decoded, magnitude and distance images and
    # the table are what pins the path).
    "raw3d_medium": dict(shape=(16, 128, 160), seed=31, density=5e-3, lowpass=None, norm="global", min_px=4,
                         origin=(-5.0, 120.0, 40.0), slim=True),
    "dense16_medium": dict(shape=(16, 128, 160), seed=32, density=1.5e-3, lowpass=None, norm="global", min_px=4,
                           bkg=195.0, nrm=30.0, nrm_jitter=3.0, mag=(0.05, 10.0), slim=True),
}


def stack_digest(stack) -> str:
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(stack).tobytes()).hexdigest()




def scenario_inputs(sc):
    """(codebook df, oracle codebook dict, stack uint16, predictor | None, bkg, nrm, excluded gene ids | None)."""
    n_bits = int(sc.get("bits", 16))
    df_cb, cb = cases.codebook16() if n_bits == 16 else cases.codebook22(n_words=int(sc.get("words", 300)))
    stack = cases.small_stack(cb["matrix"], shape=sc["shape"], seed=sc["seed"], density=sc["density"])
    rng = np.random.default_rng(sc["seed"] + 1000)
    pred = None
    if sc.get("predictor"):
        pred = rng.uniform(0.0, 1.0, size=stack.shape).astype(np.float32)
    bkg, nrm = cases.simple_vectors(n_bits, bkg=sc.get("bkg", 200.0), nrm=sc.get("nrm", 900.0), seed=sc["seed"])
    if "nrm_jitter" in sc:  # noise-level vectors: a +-50 jitter would make them negative
        j = float(sc["nrm_jitter"])
        nrm = (np.float32(sc["nrm"]) + np.random.default_rng(sc["seed"]).uniform(-j, j, n_bits)).astype(np.float32)
    excluded = None
    if sc.get("exclude"):
        genes = [g for g in df_cb["gene_id"] if not str(g).lower().startswith("blank")]
        excluded = genes[: sc["exclude"]]
    return df_cb, cb, stack, pred, bkg, nrm, excluded


def synthetic_transcript_table(df_cb, seed: int, mode: str = "3d", n: int = 6000):
    """Seeded decoded-transcript table for the post-decode stage: coding genes with
    good statistics, blank genes with poor ones, duplicates across overlapping tiles (3-D radius)
    and -- in 2-D mode -- same-gene detections split across adjacent z planes."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    genes = [str(g) for g in df_cb["gene_id"]]
    blanks = [g for g in genes if g.lower().startswith("blank")]
    coding = [g for g in genes if not g.lower().startswith("blank")]
    n_blank = n // 12
    gene = np.array(list(rng.choice(coding, n - n_blank)) + list(rng.choice(blanks, n_blank)), dtype=object)
    blank = np.array([g.lower().startswith("blank") for g in gene])
    area = np.where(blank, rng.integers(4, 14, n), rng.integers(6, 60, n)).astype(np.float64)
    mag = np.where(blank, rng.normal(1.7, 0.15, n), rng.normal(2.4, 0.4, n)).astype(np.float16).astype(np.float64)
    dist = np.where(blank, rng.uniform(0.35, 0.6, n), rng.uniform(0.05, 0.55, n)).astype(np.float16).astype(np.float64)
    tile = rng.integers(0, 4, n)
    vz = 0.315 if mode == "3d" else 1.5
    z = np.round(rng.integers(0, 12, n) * vz if mode == "2d" else rng.uniform(0, 12 * vz, n), 2)
    y = np.round(rng.uniform(0, 60, n) + (tile // 2) * 50.0, 2)
    x = np.round(rng.uniform(0, 60, n) + (tile % 2) * 50.0, 2)
    df = pd.DataFrame({"gene_id": gene, "area": area, "magnitude_mean": mag, "distance_min": dist,
                       "tile_idx": tile, "global_z": z, "global_y": y, "global_x": x})
    # cross-tile duplicates: the same molecule seen by a neighbouring tile, jittered by < 0.75 um
    dup = df.iloc[rng.choice(n, n // 8, replace=False)].copy()
    dup["tile_idx"] = (dup["tile_idx"] + rng.integers(1, 4, len(dup))) % 4
    for c, s in (("global_z", 0.2), ("global_y", 0.3), ("global_x", 0.3)):
        dup[c] = np.round(dup[c] + rng.uniform(-s, s, len(dup)), 2)
    dup["distance_min"] = (dup["distance_min"] + rng.uniform(-0.02, 0.02, len(dup))).astype(np.float16).astype(float)
    parts = [df, dup]
    if mode == "2d":
        # same tile, same gene, next plane(s), XY within one pixel: chains of 2-3 detections
        split = df.iloc[rng.choice(n, n // 6, replace=False)].copy()
        split["global_z"] = np.round(split["global_z"] + vz, 2)
        for c in ("global_y", "global_x"):
            split[c] = np.round(split[c] + rng.uniform(-0.06, 0.06, len(split)), 2)
        split["distance_min"] = (split["distance_min"] + rng.uniform(-0.03, 0.03, len(split))).astype(
            np.float16).astype(float)
        chain = split.iloc[: len(split) // 3].copy()
        chain["global_z"] = np.round(chain["global_z"] + vz, 2)
        parts += [split, chain]
    out = pd.concat(parts, ignore_index=True)
    out = out.iloc[rng.permutation(len(out))].reset_index(drop=True)
    out["gene_id"] = out["gene_id"].astype(str)
    # the columns the LR filter reads (PD:3929-3950), correlated with blank-ness like real data
    blank_all = out["gene_id"].str.lower().str.startswith("blank").to_numpy()
    m = len(out)
    out["signal_mean"] = np.where(blank_all, rng.normal(0.45, 0.1, m), rng.normal(0.8, 0.12, m))
    out["s-b_mean"] = out["signal_mean"] - rng.uniform(0.01, 0.1, m)
    for k, scale in enumerate((2.0, 1.0, 0.5)):
        out[f"inertia_tensor_eigvals-{k}"] = rng.gamma(2.0, scale, m)
    return out


def warp_tile_kwargs(sc):
    """add_tile keywords (bit rounds, round transforms, SOFIMA flow fields) of a ``warp`` scenario, plus the
    per-bit physical transforms / flows in the form the oracle takes."""
    shape = sc["shape"]
    rng = np.random.default_rng(sc["seed"] + 2000)
    bit_round = [1 + (b // 4) for b in range(16)]  # 4 bits per round, rounds 1..4
    xf = {}
    for r in (2, 3, 4):
        m = np.eye(4)
        m[:3, :3] += rng.normal(0, 2e-3, (3, 3))
        m[:3, 3] = rng.normal(0, 1.0, 3) * np.array([0.315, 0.098, 0.098])
        xf[r] = m
    stride = (2.0, 8.0, 8.0)
    box_start_xyz = (4.0, 4.0, 1.0)
    fshape = (3, shape[0] // 2, shape[1] // 8, shape[2] // 8)
    flows = {}
    for r, status in ((2, "ok"), (3, "ok"), (4, "identity_fallback_no_valid_vectors")):
        field = rng.normal(0, 0.6, fshape).astype(np.float32)
        attrs = dict(map_stride_zyx_px=stride, map_box_start_xyz_px=box_start_xyz,
                     reference_shape_zyx_px=tuple(int(v) for v in shape), sofima_status=status)
        flows[r] = (field, attrs)
    kwargs = dict(bit_round=bit_round, round_transforms_zyx_um=xf, sofima_flow_fields=flows)
    bit_xf = [np.eye(4, dtype=np.float32) if r == 1 else np.asarray(xf[r], dtype=np.float32) for r in bit_round]
    bit_flows = [None if r in (1, 4) else (flows[r][0], stride, box_start_xyz) for r in bit_round]
    return kwargs, bit_xf, bit_flows


# configs[0] (BASELINE.json: "simulation-example dataset: one 3D tile, 16-bit MHD4 codebook"): the parameter set of the
# reference's simulation CLI (cli/statphysbio_simulation/pixeldecode.py:18,259,273-276): magnitude (0.9, 10), minimum
# 28 px, reference-default low-pass, ONE tile x THREE optimiser iterations, then decode_all_tiles with the blank-fraction
# filter and no cell assignment.  The simulation volume is 16 x 48 x 512 x 512; the fixture keeps the call sequence and
# the parameters on a volume small enough to commit.
SIM_CFG0 = dict(shape=(14, 72, 80), seed=23, density=2.5e-3, magnitude=(0.9, 10.0), min_px=28, iterations=3,
                lowpass=(3.0, 1.0, 1.0), amplitude=2600.0)


def simcfg0_stack():
    df_cb, cb = cases.codebook16()
    from merfish3d_analysis_b200 import synthetic

    stack = synthetic.make_stack(cb["matrix"], SIM_CFG0["shape"], SIM_CFG0["seed"], density=SIM_CFG0["density"],
                                 amplitude=SIM_CFG0["amplitude"])
    return df_cb, cb, stack
