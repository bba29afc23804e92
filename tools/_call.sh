cd /root/repo
python tools/dense_regime.py 2>&1 | tail -1
M3D_PROBE_MODE=allfg python tools/dense_regime.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:decode_search -s 1 -c 1 -o gpurun_out/r2_dense_search_v7 -f python tools/dense_regime.py > gpurun_out/dense_ncu.log 2>&1
M3D_PROBE_MODE=allfg ncu --set full --clock-control none --import-source on -k regex:decode_search -s 1 -c 1 -o gpurun_out/r2_dense_search_v7_allfg -f python tools/dense_regime.py > gpurun_out/dense_ncu2.log 2>&1
tail -2 gpurun_out/dense_ncu2.log | cut -c1-200
