# r4e: 4-GPU map value with the float64 arrays leaving the peers as TMA bulk stores (default) at two root shares
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
i=0
for extra in "" "--root-share 0.25"; do
  i=$((i+1))
  $TR --master-port 2958$i bench.py --gpus 4 --no-other --no-cpu --steps 20 --warmup 3 $extra 2>gpurun_out/r4e_$i.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$extra', 'value %.4e'%d['value'], 'ms %.3f'%d['ms_per_step'], d['config']['parallelism'][40:110], 'kernel_only %.3e'%d['kernel_only']['value'], 'ingress', round(d['gather_limits']['ingress_GBs_achieved']), 'mismatch', d['sharded_vs_single']['mismatched_bytes'])" || tail -5 gpurun_out/r4e_$i.err
done
