TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for u in 0 1; do echo "U8ROWS=$u"; MLM_MAP_U8ROWS=$u timeout 150 $TR --nproc-per-node 8 --master-port 2951$u tools/time_gather.py --gathers p2p --shares 0.27,0.30,0.33 --out gpurun_out/r2g_gather_n8_u$u.json 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['world'], 'share %.3f (%d/%d)'%(d['root_share'], d['root_positions'], d['period']), 'ms %.3f kernel %.3f'%(d['ms'], d['ms_kernel_only']), 'mags/s %.3e'%d['mags_per_s'], d['check']['mismatched_bytes'])"; done
for u in 0 1; do echo "U8ROWS=$u"; MLM_MAP_U8ROWS=$u timeout 150 $TR --nproc-per-node 4 --master-port 2952$u tools/time_gather.py --gathers p2p --shares 0.25,0.29,0.32 --out gpurun_out/r2g_gather_n4_u$u.json 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['world'], 'share %.3f (%d/%d)'%(d['root_share'], d['root_positions'], d['period']), 'ms %.3f kernel %.3f'%(d['ms'], d['ms_kernel_only']), 'mags/s %.3e'%d['mags_per_s'], d['check']['mismatched_bytes'])"; done
timeout 150 $TR --nproc-per-node 2 --master-port 29531 tools/time_gather.py --gathers p2p --shares 0.5 --out gpurun_out/r2g_gather_n2.json 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['world'], 'share %.3f'%d['root_share'], 'ms %.3f kernel %.3f'%(d['ms'], d['ms_kernel_only']), 'mags/s %.3e'%d['mags_per_s'], d['check']['mismatched_bytes'])"
