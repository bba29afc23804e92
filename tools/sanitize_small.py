"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): tiny volumes only."""
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import cases  # noqa: E402
from scenarios import SCENARIOS, scenario_inputs, synthetic_transcript_table, warp_tile_kwargs  # noqa: E402
from merfish3d_analysis_b200.datastore import ArrayDataStore  # noqa: E402
from merfish3d_analysis_b200.PixelDecoder import PixelDecoder  # noqa: E402

import os

# M3D_SANITIZE_DENSE_ONLY=1: only the dense-candidate scenarios (tensor-core marking + warp-shared pair evaluation), small
# enough for `compute-sanitizer --tool racecheck`
DENSE_ONLY = os.environ.get("M3D_SANITIZE_DENSE_ONLY", "0") == "1"
tmp = Path(tempfile.mkdtemp())
n = 0
for name in (() if DENSE_ONLY else ("raw3d", "lp3d", "mode2d", "chroma", "bits22", "warp")):
    sc = SCENARIOS[name]
    df_cb, _cb, stack, pred, bkg, nrm, excluded = scenario_inputs(sc)
    ds = ArrayDataStore(tmp / name, codebook=df_cb, microscope_type=sc.get("microscope", "3D"))
    extra = warp_tile_kwargs(sc)[0] if sc.get("warp") else {}
    ds.add_tile(stack, predictors=pred, **extra)
    ds.add_tile(stack, predictors=pred, **extra)
    ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
    dec = PixelDecoder(ds, merfish_bits=int(sc.get("bits", 16)), verbose=0, z_range=sc.get("z_range"))
    dec._optimize_normalization_weights = dec._collect_chromatic_centroids = bool(sc.get("chroma"))
    kw = dict(lowpass_sigma=sc["lowpass"], minimum_pixels=sc["min_px"], magnitude_threshold=sc.get("mag"),
              normalization_method="global")
    dec.decode_one_tile(0, return_results=True, **kw)   # dense kernel (tensor-core marking) + all images
    dec.decode_one_tile(0, **kw)                        # gate + search + fused labelling
    n += len(dec.decoded_barcodes)
    if not sc.get("chroma"):
        dec.decode_all_tiles(assign_to_cells=False, **kw)  # prefetch + table stage
    dec._cleanup()
# noise-level vectors: dense-candidate regime through the production path
sc = SCENARIOS["raw3d"]
df_cb, cb, stack, *_ = scenario_inputs(sc)
ds = ArrayDataStore(tmp / "dense", codebook=df_cb)
ds.add_tile(stack)
ds.save_decode_normalization_vectors(None, "global", np.full(16, 17.0, np.float32), np.full(16, 187.0, np.float32))
dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
dec.decode_one_tile(0, lowpass_sigma=None, minimum_pixels=4, normalization_method="global")
n += len(dec.decoded_barcodes)
# round 2: K = 385 / 1000 in the dense regime (two fragment k-steps, wide candidate sets), the optimiser with its tile cache
# and device-resident tables, and the image store with both device entropy decoders
for name in ("bits22_k385_dense", "bits22_k1000_dense", "dense16"):
    sc = SCENARIOS[name]
    df_cb2, _cb2, stack2, pred2, bkg2, nrm2, _ex = scenario_inputs(sc)
    ds = ArrayDataStore(tmp / name, codebook=df_cb2, microscope_type=sc.get("microscope", "3D"))
    ds.add_tile(stack2)
    ds.save_decode_normalization_vectors(None, "global", nrm2, bkg2)
    dec = PixelDecoder(ds, merfish_bits=int(sc.get("bits", 16)), verbose=0)
    kw = dict(lowpass_sigma=None, minimum_pixels=sc["min_px"], magnitude_threshold=sc.get("mag"), normalization_method="global")
    dec.decode_one_tile(0, return_results=True, **kw)
    dec.decode_one_tile(0, **kw)
    n += len(dec.decoded_barcodes)
    dec._cleanup()
if DENSE_ONLY:
    print("sanitize_small (dense only): transcripts", n)
    sys.exit(0)
ds = ArrayDataStore(tmp / "opt", codebook=df_cb)
for k in range(3):
    ds.add_tile(cases.small_stack(cb["matrix"], shape=(8, 40, 48), seed=300 + k, density=5e-3))
dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
dec.optimize_normalization_by_decoding(n_iterations=3, minimum_pixels=4, lowpass_sigma=(3.0, 1.0, 1.0),
                                       magnitude_threshold=(0.9, 10.0), tile_indices=[0, 1, 2])
from merfish3d_analysis_b200 import zarr_store as zs  # noqa: E402

rng = np.random.default_rng(3)
img = (rng.poisson(100, (16, 96, 80)) + 100).astype(np.uint16)
for comp in ("blosc-zstd", "blosc-lz4"):
    zs.write_ome_image(tmp / f"img_{comp}", img, chunks=(8, 32, 32), compression=comp)
    dst = torch.zeros(img.shape, dtype=torch.uint16, device="cuda")
    zs.transfer(dec._ctx(0), [(zs.ZarrImage(tmp / f"img_{comp}.ome.zarr"), dst)])
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy(), img), comp
# table stage incl. 2-D within-tile clusters
for mode, micro in (("3d", "3D"), ("2d", "2D")):
    ds = ArrayDataStore(tmp / f"tab{mode}", codebook=df_cb, microscope_type=micro,
                        voxel_size_zyx_um=(0.315, 0.098, 0.098) if mode == "3d" else (1.5, 0.1085, 0.1085))
    for _ in range(4):
        ds.add_tile(np.zeros((16, 2, 4, 4), dtype=np.uint16))
    dec = PixelDecoder(ds, merfish_bits=16, verbose=0)
    dec._df_barcodes_loaded = synthetic_transcript_table(df_cb, seed=1, mode=mode, n=1500)
    dec._filter_all_barcodes_blank_fraction()
    if mode == "2d":
        dec._remove_duplicates_within_tile(0.1085, 1.5)
    dec._remove_duplicates_in_tile_overlap()
    n += len(dec._df_filtered_barcodes)
torch.cuda.synchronize()
print("sanitize_small ok", n)
