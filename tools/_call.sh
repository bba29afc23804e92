cd /root/repo
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo bench rc=$?
tail -c 600 gpurun_out/r2_bench_n1_c.err
