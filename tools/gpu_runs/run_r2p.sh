python tools/unordered_once.py && ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2p_unordered_launches.csv python tools/unordered_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2p_unordered_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-24:]: print(r[4][:40], r[-3], r[-1])
PY
