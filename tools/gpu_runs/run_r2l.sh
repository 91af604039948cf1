# r2l: test-suite, C5 timing with the new segment rule, fused-fit kernel at 6 vs 8 resident blocks
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/profile_models.py 10000 3 2>&1 | tail -1
python tools/profile_models.py 1250 3 2>&1 | tail -1
for lib in "" variants/lib_chi8.so; do MLM_LIB=$lib python bench.py --no-cpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('value %.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
for k,v in d['other_configs'].items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','segment_epochs','e2e_host_api')})
"; done
