set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_reference_golden.py -x -q 2>&1 | tail -3
python tools/dense_regime.py > gpurun_out/dense_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:decode_search -s 1 -c 1 -o gpurun_out/r2_dense_search_v4 python tools/dense_regime.py > gpurun_out/dense_ncu.log 2>&1
tail -2 gpurun_out/dense_plain.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --extras all_foreground 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step']); print(d['extras']['all_foreground'])"
