cd /root/repo
echo "== full gpu suite"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench N=1 default"
( time timeout 1500 python bench.py > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err ) 2>&1 | tail -4
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n1_e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_|ccl_|features_|reset_foreground|DeviceRadixSort|DeviceScan' -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-200
