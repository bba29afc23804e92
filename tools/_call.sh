cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "weight or select_hist or lowpass" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --extras optimizer 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['extras']['optimizer']
print('total', o['total_s'], 'seed', o['seed_s'], 'it0', o['iteration0']['total_s'], 'steady', o['steady_s_per_iteration'])
print(o['kernel_ms_whole_run_rank0'])"
