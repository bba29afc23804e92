set -x
python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_|ccl_|features_|reset_foreground|DeviceRadixSort|DeviceScan" -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu1.log 2>&1
grep -vc "^==" gpurun_out/r2_launches_cfg2.csv
