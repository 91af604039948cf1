python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/unordered_once.py > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_unordered_launches.csv python tools/unordered_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2e_unordered_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-22:]: print(r[4][:50], r[-1])
PY
bash tools/run_quick.sh 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_store_probe.py 2>&1 | grep -E "^(peer|local)"
