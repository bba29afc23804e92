cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "lowpass or low_pass" 2>&1 | tail -2
timeout 120 python tools/lowpass_probe.py 2>&1 | tail -1
timeout 120 python tools/lowpass_probe.py 64 2048 2048 2>&1 | tail -1
timeout 120 python tools/lowpass_probe.py 110 2048 2048 2>&1 | tail -1
cp merfish3d-analysis_b200/libm3d_b200.so /tmp/lib_A.so
cp merfish3d-analysis_b200/build/lib_B.so merfish3d-analysis_b200/libm3d_b200.so
echo variant B
timeout 120 python tools/lowpass_probe.py 2>&1 | tail -1
timeout 120 python tools/lowpass_probe.py 100 2048 2048 float32 2>&1 | tail -1
cp /tmp/lib_A.so merfish3d-analysis_b200/libm3d_b200.so
