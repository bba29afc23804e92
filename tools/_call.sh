set -x
python tools/multi_tile_profile.py 2>&1 | tail -60
python tools/opt_it0_profile.py 2>&1 | tail -50
