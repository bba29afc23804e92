set -x
for v in 1 0 1 0; do M3D_SEED_OVERLAP=$v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --extras optimizer 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
o=d['extras']['optimizer']
print('overlap=$v', {k:round(o[k],4) for k in ('total_s','seed_s','steady_s_per_iteration')}, 'it0', round(o['iteration0']['total_s'],3), round(o['iteration0']['decode_extract_s'],3), round(o['iteration0']['exchange_s'],3))"; done
