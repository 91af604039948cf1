# r4b: 4-GPU map value against the root share (final kernels)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
i=0
for extra in "" "--root-share 0.30" "--root-share 0.33"; do
  i=$((i+1))
  $TR --master-port 2957$i bench.py --gpus 4 --no-other --no-cpu --steps 20 --warmup 3 $extra 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$extra', 'value %.4e'%d['value'], 'ms %.3f'%d['ms_per_step'], d['config']['parallelism'][40:110], 'kernel_only %.3e'%d['kernel_only']['value'], 'ingress', round(d['gather_limits']['ingress_GBs_achieved']), 'mismatch', d['sharded_vs_single']['mismatched_bytes'])"
done
