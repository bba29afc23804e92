set -x
python -m pytest tests/test_gpu_reference_golden.py tests/test_gpu_pixeldecoder.py tests/test_gpu_zarr_store.py -x -q -k "optimizer or simulation or unregistered or without_any or multi_gpu" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --extras optimizer 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d['extras']['optimizer']
print({k:o[k] for k in ('total_s','seed_s','steady_s_per_iteration','steady_gvoxel_per_s','iteration0_gvoxel_per_s','cache')}); print(o['iteration0']); print(o['iterative_normalization_head'])"
