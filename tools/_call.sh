cd /root/repo
python tools/lowpass_probe.py > gpurun_out/lp_plain.log 2>&1 || { tail -5 gpurun_out/lp_plain.log; exit 1; }
cat gpurun_out/lp_plain.log
ncu --set full --clock-control none --import-source on -k regex:lowpass --launch-skip 2 -c 2 -o gpurun_out/r2_lowpass -f python tools/lowpass_probe.py > gpurun_out/ncu_lp.log 2>&1
tail -3 gpurun_out/ncu_lp.log
