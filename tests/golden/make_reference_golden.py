"""Golden vectors produced by the REFERENCE's own code (``/root/reference``, unmodified).

    python tests/golden/make_reference_golden.py        # writes tests/golden/reference_*.npz

``reference_shims.py`` answers the reference's CuPy / cuVS / cuCIM / scikit-image calls with
NumPy / SciPy so ``merfish3danalysis.PixelDecoder.PixelDecoder`` runs here without a GPU.  Each
fixture stores the seeded inputs and everything ``decode_one_tile(return_results=True)`` returned
plus the transcript table, so the oracle (CPU tests) and the CUDA path (GPU tests) are compared
with what the reference computed, not with each other only.  ``/root/reference`` is needed to
(re)generate, never to run the tests.
"""
from __future__ import annotations

import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import cases  # noqa: E402
import reference_shims as rs  # noqa: E402
from scenarios import SCENARIOS, scenario_inputs, stack_digest, warp_tile_kwargs  # noqa: E402
from merfish3d_analysis_b200.datastore import ArrayDataStore  # noqa: E402

OUT = ROOT / "tests" / "golden"

def table_arrays(df):
    num = [c for c in df.columns if c != "gene_id"]
    return dict(
        table_columns=np.array(num), table=df[num].to_numpy(dtype=np.float64),
        gene_id=np.array([str(g) for g in df["gene_id"]]),
    )


def run_tile_scenario(name, sc, RefPD):
    df_cb, cb, stack, pred, bkg, nrm, excluded = scenario_inputs(sc)
    tmp = Path(tempfile.mkdtemp())
    try:
        ds = ArrayDataStore(tmp / "qi2labdatastore", codebook=df_cb, microscope_type=sc.get("microscope", "3D"))
        extra = warp_tile_kwargs(sc)[0] if sc.get("warp") else {}
        ds.add_tile(stack, predictors=pred, stage_origin_zyx_um=sc.get("origin"), **extra)
        ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
        dec = RefPD(ds, merfish_bits=int(sc.get("bits", 16)), verbose=0, z_range=sc.get("z_range"), excluded_gene_ids=excluded)
        if sc.get("chroma"):
            dec._optimize_normalization_weights = True
            dec._collect_chromatic_centroids = True
        with rs.pandas2_semantics():
            res = dec.decode_one_tile(
                0, return_results=True, lowpass_sigma=sc["lowpass"], minimum_pixels=sc["min_px"],
                magnitude_threshold=sc.get("mag"), normalization_method=sc["norm"],
            )
        df = dec._df_barcodes
        out = dict(
            stack=stack, bkg=bkg, nrm=nrm,
            image=np.asarray(res[0], dtype=np.float32), scaled=np.asarray(res[1]), magnitude=np.asarray(res[2]),
            distance=np.asarray(res[3]), decoded=np.asarray(res[4]),
            pixel_threshold=np.float64(dec._pixel_assignment_threshold),
            transcript_threshold=np.float64(dec._transcript_distance_threshold),
            excluded=np.array([] if excluded is None else excluded, dtype=str),
            **table_arrays(df),
        )
        if sc.get("slim"):  # medium volumes: inputs by seed + checksum, no float32 copy / scaled images
            for k in ("stack", "image", "scaled"):
                del out[k]
            out["stack_sha256"] = np.array(stack_digest(stack))
        np.savez_compressed(OUT / f"reference_{name}.npz", **out)
        print(f"{name}: {len(df)} transcripts, {int((out['decoded'] >= 0).sum())} decoded voxels")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_optimizer_scenario(RefPD):
    """optimize_normalization_by_decoding (PD:4581-4757) over 3 tiles x 3 iterations, in-process.

    The reference spawns one process per GPU (PD:142-204); here the worker function runs
    synchronously in this process (same code, same order) because the stand-in modules live
    in this interpreter.  The datastore is re-opened by path inside the worker as upstream."""
    import torch

    import merfish3danalysis.PixelDecoder as refmod

    df_cb, cb = cases.codebook16()
    tmp = Path(tempfile.mkdtemp())
    try:
        ds = ArrayDataStore(tmp / "qi2labdatastore", codebook=df_cb)
        stacks = []
        for t in range(3):
            st = cases.small_stack(cb["matrix"], shape=(8, 40, 48), seed=300 + t, density=5e-3)
            stacks.append(st)
            ds.add_tile(st, persist=True)

        class _Done:
            exitcode = 0
            pid = 0

            def join(self):
                return None

        def run_inline(*, target, args, physical_gpu_id):
            with rs.pandas2_semantics():
                target(*args)
            return _Done()

        saved = (refmod._start_gpu_worker_process, refmod.qi2labDataStore, torch.cuda.set_device)
        refmod._start_gpu_worker_process = run_inline
        refmod.qi2labDataStore = lambda path, validate=False: ArrayDataStore(path)
        torch.cuda.set_device = lambda *_a, **_k: None
        try:
            dec = RefPD(ds, merfish_bits=16, verbose=0)
            with rs.pandas2_semantics():
                dec.optimize_normalization_by_decoding(
                    n_iterations=3, minimum_pixels=4, lowpass_sigma=None, magnitude_threshold=(0.9, 10.0),
                    tile_indices=[0, 1, 2],
                )
        finally:
            refmod._start_gpu_worker_process, refmod.qi2labDataStore, torch.cuda.set_device = saved
        ds2 = ArrayDataStore(tmp / "qi2labdatastore")
        g = ds2.load_decode_normalization_vectors(None, "global")
        it = ds2.load_decode_normalization_vectors(None, "iterative")
        np.savez_compressed(
            OUT / "reference_optimizer.npz", stacks=np.stack(stacks),
            global_normalization=np.asarray(g[0], dtype=np.float32), global_background=np.asarray(g[1], dtype=np.float32),
            iterative_normalization=np.asarray(it[0], dtype=np.float32),
            iterative_background=np.asarray(it[1], dtype=np.float32),
        )
        print("optimizer: global", np.asarray(g[0])[:4], np.asarray(g[1])[:4], "iterative", np.asarray(it[0])[:4],
              np.asarray(it[1])[:4])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_simcfg0_scenario(RefPD):
    """configs[0]: the call sequence and parameters of the reference's simulation CLI
    (cli/statphysbio_simulation/pixeldecode.py:259-292) on one tile: optimize_normalization_by_decoding(
    n_random_tiles=1, n_iterations=3, magnitude (0.9, 10), minimum 28 px, default low-pass) followed by
    decode_all_tiles(assign_to_cells=False, ...) with the blank-fraction filter.  Workers run in-process as in
    ``run_optimizer_scenario``."""
    import torch

    import merfish3danalysis.PixelDecoder as refmod
    from scenarios import SIM_CFG0, simcfg0_stack

    df_cb, _cb, stack = simcfg0_stack()
    tmp = Path(tempfile.mkdtemp())
    try:
        ds = ArrayDataStore(tmp / "qi2labdatastore", codebook=df_cb)
        ds.add_tile(stack, persist=True)

        class _Done:
            exitcode = 0
            pid = 0

            def join(self):
                return None

        def run_inline(*, target, args, physical_gpu_id):
            with rs.pandas2_semantics():
                target(*args)
            return _Done()

        saved = (refmod._start_gpu_worker_process, refmod.qi2labDataStore, torch.cuda.set_device)
        refmod._start_gpu_worker_process = run_inline
        refmod.qi2labDataStore = lambda path, validate=False: ArrayDataStore(path)
        torch.cuda.set_device = lambda *_a, **_k: None
        try:
            dec = RefPD(datastore=ds, use_mask=False, merfish_bits=16, verbose=0)
            with rs.pandas2_semantics():
                dec.optimize_normalization_by_decoding(
                    n_random_tiles=1, n_iterations=SIM_CFG0["iterations"], lowpass_sigma=SIM_CFG0["lowpass"],
                    magnitude_threshold=SIM_CFG0["magnitude"], minimum_pixels=SIM_CFG0["min_px"],
                    feature_predictor_threshold=0.5, estimate_chromatic_affines=False,
                )
                dec.decode_all_tiles(
                    assign_to_cells=False, lowpass_sigma=SIM_CFG0["lowpass"], magnitude_threshold=SIM_CFG0["magnitude"],
                    minimum_pixels=SIM_CFG0["min_px"], feature_predictor_threshold=0.5, duplicate_radius_xy=None,
                    duplicate_radius_z=None, filter_method="blank_fraction", target_gross_misid_rate=0.05,
                    lr_fdr_target=0.05,
                )
        finally:
            refmod._start_gpu_worker_process, refmod.qi2labDataStore, torch.cuda.set_device = saved
        ds2 = ArrayDataStore(tmp / "qi2labdatastore")
        g = ds2.load_decode_normalization_vectors(None, "global")
        it = ds2.load_decode_normalization_vectors(None, "iterative")
        tile = ds2.load_local_decoded_spots(0)
        filt = ds2.load_global_filtered_decoded_spots()
        t_arr = table_arrays(tile)
        f_arr = table_arrays(filt)
        np.savez_compressed(
            OUT / "reference_simcfg0.npz", stack=stack,
            global_normalization=np.asarray(g[0], dtype=np.float32), global_background=np.asarray(g[1], dtype=np.float32),
            iterative_normalization=np.asarray(it[0], dtype=np.float32),
            iterative_background=np.asarray(it[1], dtype=np.float32),
            tile_table_columns=t_arr["table_columns"], tile_table=t_arr["table"], tile_gene_id=t_arr["gene_id"],
            filtered_table_columns=f_arr["table_columns"], filtered_table=f_arr["table"], filtered_gene_id=f_arr["gene_id"],
        )
        print(f"simcfg0: iterative {np.asarray(it[0])[:4]} {np.asarray(it[1])[:4]}; {len(tile)} transcripts decoded, "
              f"{len(filt)} after the blank-fraction filter; area min {tile['area'].min() if len(tile) else None}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    """``python make_reference_golden.py [name ...]``: all fixtures, or only the named ones
    (scenario names, ``optimizer``, ``simcfg0``)."""
    only = set(sys.argv[1:])
    RefPD = rs.load_reference_pixeldecoder()
    for name, sc in SCENARIOS.items():
        if not only or name in only:
            run_tile_scenario(name, sc, RefPD)
    if not only or "optimizer" in only:
        run_optimizer_scenario(RefPD)
    if not only or "simcfg0" in only:
        run_simcfg0_scenario(RefPD)


if __name__ == "__main__":
    main()
