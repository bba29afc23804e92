set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python tools/zstd_device_probe.py 2 > gpurun_out/zstd_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blosc_zstd_decode -s 1 -c 1 -o gpurun_out/r2_zstd_v2_fast python tools/zstd_device_probe.py 2 > gpurun_out/zstd_ncu.log 2>&1
tail -1 gpurun_out/zstd_plain.log
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err) 2>&1 | tail -3
tail -3 gpurun_out/r2_bench_n1_b.err
