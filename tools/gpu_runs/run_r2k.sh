# r2k: k_models at 8 resident blocks (128 registers, no spill), reconvergence before the sums, auto_segment retuned
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/profile_map.py 8192 4 2>&1 | tail -1
for seg in 0 16 24 32 48 64; do echo "seg=$seg"; MLM_SEGMENT=$seg python tools/profile_models.py 10000 3 2>&1 | tail -1; done
python tools/count_events_models.py > gpurun_out/r2k_events_c5.json 2>gpurun_out/r2k_events_c5.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_events_c5.json')); print(d['segment'], d['per_position']); print(d['active_lanes'])"
python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null > gpurun_out/r2k_bench_nocpu.json; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_bench_nocpu.json') if l.startswith('{')][0])
print('value %.4e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
for k,v in d['other_configs'].items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','segment_epochs')})
for k,v in d['parity'].items(): print(k, v.get('nimg_mismatch'), v.get('nimg_mismatch_unaccounted'), v.get('disputes_outside_rule3_mask'), v.get('failed_rules'))
"
