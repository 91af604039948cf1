python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/profile_models.py 10000 3 2>&1 | tail -1
python tools/profile_models.py 1250 3 2>&1 | tail -1
python bench.py --no-cpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('value %.4e'%d['value'])
for k,v in d['other_configs'].items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','segment_epochs')})
"
