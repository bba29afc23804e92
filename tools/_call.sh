set -x
python -m pytest tests/test_gpu_zarr_store.py -x -q -k "zstd or oracle_written" 2>&1 | tail -3
for m in 2; do M3D_ZARR_GPU_ZSTD=$m timeout 600 python tools/zarr_io_bench.py --z 32 --skip-host --only-transfer 2>&1 | tail -1 | cut -c1-700; done
timeout 300 python tools/zstd_device_probe.py 2 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --extras optimizer 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d['extras']['optimizer']
print({k:round(o[k],4) for k in ('total_s','seed_s','steady_s_per_iteration')}, 'it0', {k: round(v,3) for k,v in o['iteration0'].items() if k.endswith('_s')})"
