"""Image store -> HBM throughput: one synthetic tile written as a reference-layout datastore
(`<image>.ome.zarr`, Zarr v3, blosc-zstd bit-shuffled (16, 512, 512) chunks, page-cache warm) and then

  a. `zarr_store.transfer`      chunk files -> pinned slots (host zstd threads) -> device un-shuffle/placement
  b. host decode + upload        the reference's order of work: decode to a NumPy array (same C decoder, all
                                 host threads), then the pinned-ring upload
  c. `decode_one_tile`           end to end from the store, against the same tile fed from pinned host memory

Prints one JSON object.  Usage: python tools/zarr_io_bench.py [--bits 16] [--z 48] [--yx 2048] [--out FILE]
"""
import argparse
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=16)
    ap.add_argument("--z", type=int, default=32)
    ap.add_argument("--yx", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--compression", default="blosc-zstd")
    ap.add_argument("--skip-host", action="store_true", help="skip leg b (host decode to arrays, then upload)")
    ap.add_argument("--only-transfer", action="store_true", help="leg a only (store -> device)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--sweep", default=None, help="leg a only, repeated under settings 'K=V,K=V;K=V,...' of the M3D_* "
                    "environment switches (read by the library on every call), store written once")
    args = ap.parse_args()

    import torch

    from merfish3d_analysis_b200 import synthetic
    from merfish3d_analysis_b200 import zarr_store as zs
    from merfish3d_analysis_b200.datastore import ArrayDataStore
    from merfish3d_analysis_b200.PixelDecoder import PixelDecoder

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    args.bits = 16  # the MHD4 codebook of the bench workload
    matrix = synthetic.mhd4_codebook_matrix(16)
    df_cb = synthetic.codebook_dataframe(matrix, n_blank=10)
    shape = (args.z, args.yx, args.yx)
    stack_dev = synthetic.make_stack_device(matrix, shape, 20240, device=dev)
    stack = stack_dev.cpu().numpy()
    del stack_dev
    torch.cuda.empty_cache()
    raw_bytes = stack.nbytes
    res = {"tile": f"{args.bits} bits x {shape} uint16", "compression": args.compression, "raw_gb": raw_bytes / 1e9, "host_threads": os.cpu_count()}

    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp) / "qi2labdatastore"
        ds = zs.Qi2labZarrDataStore.create(root, df_cb)
        t0 = time.perf_counter()
        cal = ds._load_calibrations_attributes()
        cal.update(num_tiles=1, num_bits=args.bits)
        ds._save_calibrations_attributes(cal)
        ds._refresh(cal)

        def write_bit(b):
            d = root / "readouts" / "tile0000" / f"bit{b + 1:03d}"
            zs.write_ome_image(d / "corrected_data", stack[b], compression=args.compression, extra_attributes={
                "round_linker": 1, "excitation_um": 0.561, "emission_um": 0.58})

        with ThreadPoolExecutor(max_workers=min(args.bits, os.cpu_count() or 4)) as ex:
            list(ex.map(write_bit, range(args.bits)))
        ds._save_entity_attributes(root / "fiducial" / "tile0000" / "round001", {
            "stage_zyx_um": [0, 0, 0], "affine_zyx_px": np.eye(4), "local_round_transform_zyx_um": np.eye(4)})
        res["write_s"] = time.perf_counter() - t0
        stored = sum(f.stat().st_size for f in root.rglob("*") if f.is_file())
        res["stored_gb"] = stored / 1e9
        res["compression_ratio"] = raw_bytes / stored
        ds = zs.Qi2labZarrDataStore(root)

        dec = PixelDecoder(ds, merfish_bits=args.bits, verbose=0)
        ctx = dec._ctx(0)
        imgs = [ds.load_local_readout_image(0, b).result() for b in range(args.bits)]
        dst = torch.empty((args.bits,) + shape, dtype=torch.uint16, device=dev)

        def run_transfer():
            zs.transfer(ctx, [(imgs[b], dst[b]) for b in range(args.bits)])
            torch.cuda.synchronize()

        run_transfer()
        assert torch.equal(dst.cpu(), torch.from_numpy(stack)), "device reader differs from the written tile"
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            run_transfer()
            ts.append(time.perf_counter() - t0)
        res["a_store_to_device"] = {"ms": 1e3 * min(ts), "decoded_gb_s": raw_bytes / min(ts) / 1e9,
                                    "stored_gb_s": stored / min(ts) / 1e9, "launches": ctx.launches_by_kernel().get(
                                        "zarr_unshuffle_place_kernel", 0)}
        if args.sweep:
            res["sweep"] = []
            for setting in args.sweep.split(";"):
                pairs = [kv.split("=", 1) for kv in setting.split(",") if kv]
                saved = {k: os.environ.get(k) for k, _ in pairs}
                for k, v in pairs:
                    os.environ[k] = v
                try:
                    run_transfer()
                    ts = []
                    for _ in range(args.reps):
                        t0 = time.perf_counter()
                        run_transfer()
                        ts.append(time.perf_counter() - t0)
                    ok = bool(torch.equal(dst.cpu(), torch.from_numpy(stack)))
                    res["sweep"].append({"setting": setting, "ms": 1e3 * min(ts), "decoded_gb_s": raw_bytes / min(ts) / 1e9,
                                         "identical": ok})
                    print("SWEEP", setting, f"{raw_bytes / min(ts) / 1e9:.1f} GB/s", "ok" if ok else "DIFFERENT", flush=True)
                finally:
                    for k, v in saved.items():
                        if v is None:
                            os.environ.pop(k, None)
                        else:
                            os.environ[k] = v
            line = json.dumps(res)
            print(line)
            if args.out:
                Path(args.out).write_text(line + "\n")
            return
        ctx.set_timing(True)
        ctx.reset_counters()
        run_transfer()
        kt = ctx.kernel_times_ms()
        res["a_store_to_device"]["unshuffle_kernel_ms_total"] = kt.get("zarr_unshuffle_place_kernel")
        res["a_store_to_device"]["lz4_kernel_ms_total"] = kt.get("blosc_lz4_decode_kernel")
        res["a_store_to_device"]["zstd_kernel_ms_total"] = kt.get("blosc_zstd_decode_kernel")
        ctx.set_timing(False)

        def run_host_then_upload():
            hosts = [np.asarray(imgs[b]) for b in range(args.bits)]
            t_mid = time.perf_counter()
            ctx.upload([(hosts[b], dst[b]) for b in range(args.bits)])
            torch.cuda.synchronize()
            return t_mid

        if not args.skip_host:
            run_host_then_upload()
            ts, th = [], []
            for _ in range(args.reps):
                t0 = time.perf_counter()
                t_mid = run_host_then_upload()
                t1 = time.perf_counter()
                ts.append(t1 - t0)
                th.append(t_mid - t0)
            res["b_host_decode_then_upload"] = {"ms": 1e3 * min(ts), "host_decode_ms": 1e3 * min(th),
                                                "decoded_gb_s": raw_bytes / min(ts) / 1e9}

        res["M3D_ZARR_GPU_ZSTD"] = os.environ.get("M3D_ZARR_GPU_ZSTD", "0")
        if args.only_transfer:
            line = json.dumps(res)
            print(line)
            if args.out:
                Path(args.out).write_text(line + "\n")
            return
        nrm = np.full(args.bits, 900.0, dtype=np.float32)  # bench.py's vectors / thresholds
        bkg = np.full(args.bits, 200.0, dtype=np.float32)
        ds.save_decode_normalization_vectors(None, "global", nrm, bkg)
        kw = dict(lowpass_sigma=None, minimum_pixels=16, normalization_method="global", magnitude_threshold=(1.5, 10.0))
        dec.decode_one_tile(0, **kw)
        n_store = len(dec.decoded_barcodes)
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            dec.decode_one_tile(0, **kw)
            ts.append(time.perf_counter() - t0)
        res["c_decode_one_tile_from_store"] = {"ms": 1e3 * min(ts), "gvoxel_s": stack[0].size / min(ts) / 1e9,
                                               "transcripts": n_store}
        table_store = dec.decoded_barcodes.copy()
        del dec, dst
        torch.cuda.empty_cache()

        ads = ArrayDataStore(Path(tmp) / "npy" / "qi2labdatastore", codebook=df_cb)
        pinned = torch.from_numpy(stack).pin_memory().numpy()
        ads.add_tile(pinned)
        ads.save_decode_normalization_vectors(None, "global", nrm, bkg)
        dec2 = PixelDecoder(ads, merfish_bits=args.bits, verbose=0)
        dec2.decode_one_tile(0, **kw)
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            dec2.decode_one_tile(0, **kw)
            ts.append(time.perf_counter() - t0)
        res["c_decode_one_tile_from_pinned_host"] = {"ms": 1e3 * min(ts), "gvoxel_s": stack[0].size / min(ts) / 1e9,
                                                     "transcripts": len(dec2.decoded_barcodes)}
        res["tables_identical"] = bool(table_store.equals(dec2.decoded_barcodes))
    line = json.dumps(res)
    print(line)
    if args.out:
        Path(args.out).write_text(line + "\n")


if __name__ == "__main__":
    main()
