# r3c: cold starts of the list/map kernels through the three-image guesses
python -m pytest tests -m gpu -q 2>&1 | tail -3
for st in 64 16; do echo strip $st; MLM_MAP_STRIP=$st python tools/profile_map.py 8192 4 2>&1 | tail -1; done
python tools/time_unordered.py 2>&1 | head -3
python bench.py --no-cpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('value %.4e'%d['value'])
for k,v in d['other_configs'].items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','segment_epochs')})
"
