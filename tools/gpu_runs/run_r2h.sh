# r2h: FP32 watch of the non-physical pair + joint reciprocals: parity subset, k_map timing, event counts, points/models timings
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for st in 64 32 16; do echo strip $st; MLM_MAP_STRIP=$st python tools/profile_map.py 8192 4 2>&1 | tail -1; done
python tools/count_events.py C4 > gpurun_out/r2h_events_c4.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2h_events_c4.json')); print({k: round(v,4) for k,v in d['per_position'].items()})"
python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null > gpurun_out/r2h_bench_nocpu.json; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench_nocpu.json') if l.startswith('{')][0])
print('value %.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'executed', {k: d['roofline']['executed'].get(k) for k in ('fp64_inst_per_mag','stale','tflops')} if isinstance(d['roofline'].get('executed'), dict) else None, 'frac', d['roofline'].get('frac'))
for k,v in d['other_configs'].items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','segment_epochs')})
"
