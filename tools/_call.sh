set -x
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:decode_gate -s 3 -c 2 -o gpurun_out/r2_gate_full python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-300
