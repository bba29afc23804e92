set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_reference_golden.py -x -q 2>&1 | tail -3
python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
python tools/dense_regime.py 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --extras all_foreground 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('A value', d['value'], d['roofline']['kernel_ms_per_step']['decode_search_kernel'], d['extras']['all_foreground']['ms_per_decode'])"
cp merfish3d-analysis_b200/libm3d_b200.so /tmp/libA.so; cp tools/_libm3d_lb5.so merfish3d-analysis_b200/libm3d_b200.so
python tools/dense_regime.py 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --extras all_foreground 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('B(lb5) value', d['value'], d['roofline']['kernel_ms_per_step']['decode_search_kernel'], d['extras']['all_foreground']['ms_per_decode'])"
python -m pytest tests/test_gpu_kernels.py -x -q -k "decode" 2>&1 | tail -2
