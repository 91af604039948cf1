python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/run_quick.sh 2>&1 | tail -12
python tools/unordered_once.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_unordered_launches.csv python tools/unordered_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2d_unordered_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-24:]: print(r[4][:60], r[-1])
PY
python bench.py --no-cpu --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('value %.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
for k,v in d['other_configs'].items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','segment_epochs','sharded_vs_single')})
"
