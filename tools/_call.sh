set -x
python -m pytest tests/test_gpu_zarr_store.py -x -q -k "zstd or truncated" 2>&1 | tail -12
for m in 1 2; do timeout 300 python tools/zstd_device_probe.py $m 2>&1 | tail -1; done
for m in 2; do M3D_ZARR_GPU_ZSTD=$m timeout 600 python tools/zarr_io_bench.py --z 32 --skip-host --only-transfer 2>&1 | tail -1 | cut -c1-700; done
