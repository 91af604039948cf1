# quick GPU check of a kernel change: parity subset + k_map timing (strip 64 and 16) + event counts
python -m pytest tests -m gpu -x -q -k "golden or map_vs_batch or bitwise or flag_semantics or lightcurve_C2 or models_C5 or full_size or uneven" 2>&1 | tail -4
python tools/profile_map.py 8192 4 2>&1 | tail -2
MLM_MAP_STRIP=16 python tools/profile_map.py 8192 3 2>&1 | tail -1
python tools/count_events.py C4 > gpurun_out/quick_events_c4.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/quick_events_c4.json')); print({k: round(v,4) for k,v in d['per_position'].items()})"
