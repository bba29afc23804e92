set -x
python -m pytest tests/test_gpu_reference_golden.py tests/test_gpu_zarr_store.py -x -q -k "golden or simulation" 2>&1 | tail -15
python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -15
python tools/opt_it0_profile.py 2>&1 | tail -75
