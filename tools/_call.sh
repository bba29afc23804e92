set -x
free -g; nproc; lscpu | grep -E "Model name|Socket|NUMA" 
python -m pytest tests/test_gpu_zarr_store.py -x -q -k "zstd or truncated" 2>&1 | tail -15
for m in 1 2; do timeout 300 python tools/zstd_device_probe.py $m 2>&1 | tail -2; done
for m in 0 2 1; do M3D_ZARR_GPU_ZSTD=$m timeout 600 python tools/zarr_io_bench.py --z 32 --skip-host --only-transfer --out gpurun_out/r2_zstd_mode$m.json 2>&1 | tail -1; done
