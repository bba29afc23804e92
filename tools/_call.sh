cd /root/repo
timeout 600 python -m pytest tests/test_gpu_zarr_store.py -x -q 2>&1 | tail -3
timeout 600 python tools/zarr_io_bench.py --sweep "M3D_ZARR_BATCH=4;M3D_ZARR_BATCH=4,M3D_ZARR_SLOTS=120;M3D_ZARR_BATCH=6,M3D_ZARR_SLOTS=120;M3D_ZARR_BATCH=8,M3D_ZARR_SLOTS=120;M3D_ZARR_BATCH=2,M3D_ZARR_SLOTS=64;M3D_ZARR_BATCH=3,M3D_ZARR_SLOTS=96;M3D_ZARR_BATCH=4,M3D_IO_THREADS=14;M3D_ZARR_BATCH=4,M3D_IO_THREADS=8;M3D_ZARR_BATCH=1,M3D_ZARR_SLOTS=48" --z 96 --out gpurun_out/zstd_batch_sweep3.json 2>&1 | grep -v "^{" | tail -20
timeout 600 python tools/zarr_io_bench.py --sweep "M3D_ZARR_BATCH=4;M3D_ZARR_BATCH=2;M3D_ZARR_BATCH=4,M3D_ZARR_SLOTS=64" --z 32 --out gpurun_out/zstd_batch_sweep4.json 2>&1 | grep -v "^{" | tail -20
