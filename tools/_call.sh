cd /root/repo
echo "== full gpu suite"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench N=1 default"
( time timeout 1500 python bench.py > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err ) 2>&1 | tail -4
tail -c 600 gpurun_out/r2_bench_n1_d.json
