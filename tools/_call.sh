cd /root/repo
cp merfish3d-analysis_b200/libm3d_b200.so /tmp/lib_new.so
echo "== parity (new)"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_reference_golden.py -x -q -m gpu 2>&1 | tail -5
for v in B new; do
  if [ $v = B ]; then cp merfish3d-analysis_b200/build/lib_B.so merfish3d-analysis_b200/libm3d_b200.so; else cp /tmp/lib_new.so merfish3d-analysis_b200/libm3d_b200.so; fi
  echo "== variant $v"
  timeout 300 python tools/dense_regime.py 2>&1 | tail -1
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --extras all_foreground 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step ms', d['ms_per_step'], d['roofline']['kernel_ms_per_step'].get('decode_search_kernel'), 'all_fg', d['extras']['all_foreground']['ms_per_decode'])"
done
cp /tmp/lib_new.so merfish3d-analysis_b200/libm3d_b200.so

