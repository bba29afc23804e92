cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_reference_golden.py tests/test_gpu_pixeldecoder.py tests/test_gpu_fullsize.py -x -q -m gpu -k "label or golden or fused or fullsize or sharded or slab" 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --extras optimizer 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['extras']['optimizer']
print('step', d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['kernel_ms_per_step'].items()})
print('total', o['total_s'], 'seed', o['seed_s'], 'it0', o['iteration0']['total_s'], 'steady', o['steady_s_per_iteration'])
print(o['kernel_ms_whole_run_rank0'])"
