python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/run_quick.sh 2>&1 | tail -11
python tools/count_events_points.py 2>&1 | grep -A1 "shuffled_plain\|random_plain" | grep lanes
python bench.py --no-cpu --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('value %.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
for k,v in d['other_configs'].items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','segment_epochs')})
"
