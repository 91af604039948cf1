# r2m: unordered lists through the two-pass direct path
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/time_unordered.py
MLM_UNORDERED=sorted python tools/time_unordered.py
python tools/unordered_once.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_unordered_launches.csv python tools/unordered_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2m_unordered_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:60], r[-1])
PY
