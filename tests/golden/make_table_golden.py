"""Golden vectors for the post-decode table stage, produced by the REFERENCE's own methods
(``_filter_all_barcodes_blank_fraction`` PD:3386, ``_remove_duplicates_in_tile_overlap`` PD:4137,
``_remove_duplicates_within_tile`` PD:4179) executed from /root/reference through
``reference_shims`` on a seeded synthetic transcript table.

    python tests/golden/make_table_golden.py        # writes tests/golden/reference_table_stage.npz
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
from scenarios import synthetic_transcript_table  # noqa: E402
from merfish3d_analysis_b200.datastore import ArrayDataStore  # noqa: E402

OUT = ROOT / "tests" / "golden"


def main():
    RefPD = rs.load_reference_pixeldecoder()
    df_cb, _cb = cases.codebook16()
    tmp = Path(tempfile.mkdtemp())
    out = {}
    try:
        for mode, micro in (("3d", "3D"), ("2d", "2D")):
            ds = ArrayDataStore(tmp / f"store_{mode}", codebook=df_cb, microscope_type=micro,
                                voxel_size_zyx_um=(0.315, 0.098, 0.098) if mode == "3d" else (1.5, 0.1085, 0.1085))
            for _ in range(4):
                ds.add_tile(np.zeros((16, 2, 4, 4), dtype=np.uint16))
            dec = RefPD(ds, merfish_bits=16, verbose=0)
            table = synthetic_transcript_table(df_cb, seed=5150 if mode == "3d" else 5151, mode=mode)
            table["_row"] = np.arange(len(table))
            with rs.pandas2_semantics():
                dec._df_barcodes_loaded = table.copy()
                dec._filter_all_barcodes_blank_fraction(target_gross_misid_rate=0.05)
                diag = dec._blank_fraction_filter_results
                kept_filter = dec._df_filtered_barcodes["_row"].to_numpy()
                if mode == "2d":
                    vs = ds.voxel_size_zyx_um
                    dec._remove_duplicates_within_tile(radius_xy=float(vs[-1]), radius_z=float(vs[0]))
                    kept_within = dec._df_filtered_barcodes["_row"].to_numpy()
                    out["kept_within_2d"] = kept_within
                    print("within-tile collapse:", len(kept_filter), "->", len(kept_within))
                dec._remove_duplicates_in_tile_overlap()
                kept_overlap = dec._df_filtered_barcodes["_row"].to_numpy()
            # logistic-regression filter (PD:3907-4058, scikit-learn on the host in the reference too)
            with rs.pandas2_semantics():
                dec._df_barcodes_loaded = table.copy()
                dec._barcodes_filtered = False
                dec._filter_all_barcodes_LR(lr_fdr_target=0.05)
            out[f"kept_lr_{mode}"] = dec._df_filtered_barcodes["_row"].to_numpy()
            out[f"lr_probability_{mode}"] = dec._df_filtered_barcodes["predicted_probability"].to_numpy(dtype=float)
            print(mode, "LR filter keeps", len(out[f"kept_lr_{mode}"]), "of", len(table))
            out[f"kept_filter_{mode}"] = kept_filter
            out[f"kept_overlap_{mode}"] = kept_overlap
            out[f"chosen_threshold_{mode}"] = np.float64(diag["chosen_threshold"])
            out[f"achieved_rate_{mode}"] = np.float64(diag["achieved_gross_misid_rate"])
            out[f"target_reached_{mode}"] = np.bool_(diag["target_reached"])
            out[f"all_histogram_{mode}"] = np.asarray(diag["all_histogram"])
            out[f"blank_histogram_{mode}"] = np.asarray(diag["blank_histogram"])
            out[f"blank_fraction_histogram_{mode}"] = np.asarray(diag["blank_fraction_histogram"])
            for k in ("intensity_bins", "voxel_number_bins", "vector_distance_bins"):
                out[f"{k}_{mode}"] = np.asarray(diag[k])
            out[f"sweep_{mode}"] = diag["threshold_sweep"].to_numpy(dtype=float)
            print(mode, "rows", len(table), "after filter", len(kept_filter), "after dedup", len(kept_overlap),
                  "threshold", diag["chosen_threshold"], "rate", diag["achieved_gross_misid_rate"],
                  "reached", diag["target_reached"])
        np.savez_compressed(OUT / "reference_table_stage.npz", **out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
