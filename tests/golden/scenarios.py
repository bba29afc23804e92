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
    "excl_crop": dict(shape=(9, 32, 40), seed=14, density=6e-3, lowpass=None, norm="global", min_px=3,
                      z_range=(2, 8), exclude=3),
}




def scenario_inputs(sc):
    """(codebook df, oracle codebook dict, stack uint16, predictor | None, bkg, nrm, excluded gene ids | None)."""
    df_cb, cb = cases.codebook16()
    stack = cases.small_stack(cb["matrix"], shape=sc["shape"], seed=sc["seed"], density=sc["density"])
    rng = np.random.default_rng(sc["seed"] + 1000)
    pred = None
    if sc.get("predictor"):
        pred = rng.uniform(0.0, 1.0, size=stack.shape).astype(np.float32)
    bkg, nrm = cases.simple_vectors(16, bkg=sc.get("bkg", 200.0), nrm=sc.get("nrm", 900.0), seed=sc["seed"])
    excluded = None
    if sc.get("exclude"):
        genes = [g for g in df_cb["gene_id"] if not str(g).lower().startswith("blank")]
        excluded = genes[: sc["exclude"]]
    return df_cb, cb, stack, pred, bkg, nrm, excluded
