# r2s: bench under torchrun at N GPUs (argument), like the driver launches it
N=$1
s=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2s_bench_n$N.json 2> gpurun_out/r2s_bench_n$N.err; echo bench=$? $(( $(date +%s) - s )) s
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2s_bench_n$N.json') if l.startswith('{')][0])
print('value %.4e'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'ceiling frac', d['e2e'].get('d2h_ceiling',{}).get('frac_of_ceiling'))
print('config', d['config']['parallelism']); print('kernel_only', d.get('kernel_only',{}).get('value'), 'a4', d.get('a4_outputs_only',{}).get('value'), 'a4 e2e', d.get('a4_outputs_only',{}).get('e2e',{}).get('value'))
print('sharded_vs_single', d.get('sharded_vs_single')); print('gather_limits', d.get('gather_limits'))
for k,v in (d.get('other_configs') or {}).items(): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mags_per_s','sharded_vs_single','gather')})
PY
