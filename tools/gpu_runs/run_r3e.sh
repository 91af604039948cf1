# r3e: 8-GPU map value against the number of timed steps and the root share
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for extra in "--steps 20 --warmup 3" "--steps 20 --warmup 3 --root-share 0.27" "--steps 20 --warmup 3 --root-share 0.33" "--steps 5 --warmup 3"; do
  i=$((i+1))
  $TR --master-port 2956$i bench.py --gpus 8 --no-other --no-cpu $extra 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$extra', 'value %.4e'%d['value'], 'ms %.3f'%d['ms_per_step'], d['config']['parallelism'][40:110], 'kernel_only %.3e'%d['kernel_only']['value'], 'ingress', round(d['gather_limits']['ingress_GBs_achieved']))"
done
