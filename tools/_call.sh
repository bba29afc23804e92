cd /root/repo
echo "== N=2 bench"
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err ) 2>&1 | tail -3
python - <<'PY'
import json
txt=open('gpurun_out/r2_bench_n2_b.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])
o=d['extras']['optimizer']; print('opt total', o['total_s'], 'seed', o['seed_s'], 'it0', o['iteration0']['total_s'], 'steady', o['steady_s_per_iteration'], o['exchange'])
z=d['extras']['zslab']; print('zslab', z['s_per_volume'], z['gvoxel_per_s'], z.get('identical_to_unsharded'))
PY
echo "== dist smoke"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_smoke.py 2>&1 | tail -5
echo "== 2-GPU test"
timeout 600 python -m pytest tests -x -q -m gpu -k "two_gpu or 2gpu or multi_gpu or devices" 2>&1 | tail -3
