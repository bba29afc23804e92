cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "lowpass or low_pass" 2>&1 | tail -2
timeout 120 python tools/lowpass_probe.py 2>&1 | tail -1
M3D_LOWPASS_NO_ASYNC=1 timeout 120 python tools/lowpass_probe.py 2>&1 | tail -1
timeout 120 python tools/lowpass_probe.py 100 2048 2048 float32 2>&1 | tail -1
timeout 120 python tools/lowpass_probe.py 64 2048 2048 2>&1 | tail -1
