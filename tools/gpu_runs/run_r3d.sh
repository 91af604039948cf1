python -m pytest tests -m gpu -q -x 2>&1 | tail -40
for st in 64 16; do echo strip $st; MLM_MAP_STRIP=$st python tools/profile_map.py 8192 4 2>&1 | tail -1; done
python tools/time_unordered.py 2>&1 | head -3
