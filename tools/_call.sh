cd /root/repo
cp merfish3d-analysis_b200/libm3d_b200.so /tmp/lib_new.so
echo "== old kernel"; cp merfish3d-analysis_b200/build/lib_B.so merfish3d-analysis_b200/libm3d_b200.so; timeout 300 python tools/select_hist_probe.py 2>&1 | tail -5
echo "== new kernel"; cp /tmp/lib_new.so merfish3d-analysis_b200/libm3d_b200.so; timeout 300 python tools/select_hist_probe.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_reference_golden.py tests/test_gpu_pixeldecoder.py -x -q -m gpu -k "select_hist or optimizer or median or normalization or global" 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --extras optimizer 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['extras']['optimizer']
print('total', o['total_s'], 'seed', o['seed_s'], 'it0', o['iteration0']['total_s'], 'steady', o['steady_s_per_iteration'])
print(o['kernel_ms_whole_run_rank0'])"
