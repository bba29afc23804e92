set -x
python -m pytest tests/test_gpu_zarr_store.py -x -q 2>&1 | tail -2
for comp in blosc-zstd blosc-lz4; do timeout 600 python tools/zarr_io_bench.py --z 32 --skip-host --only-transfer --reps 4 --compression $comp 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['a_store_to_device']
print('RES $comp default', round(a['decoded_gb_s'],1), 'GB/s')"; done
python bench.py --steps 5 --warmup 3 --no-cpu --extras zarr 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['extras'].items(): print('RES', k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a!='note'})"
