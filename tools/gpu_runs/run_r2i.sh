# r2i: profiles with per-phase attribution: k_map<4> (FP32 watch), k_models<4> on C5, k_points<4> shuffled; event counters of C5
python -m pytest tests -m gpu -x -q -k "golden or bitwise or flag_semantics or models_C5 or lightcurve_C2 or full_size" 2>&1 | tail -2
python tools/profile_map.py 8192 4 2>&1 | tail -1
python tools/profile_models.py 10000 3 2>&1 | tail -1
python tools/profile_points.py 3 2>&1 | tail -4
python tools/count_events_models.py > gpurun_out/r2i_events_c5.json 2>gpurun_out/r2i_events_c5.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_events_c5.json')); print(d['segment'], d['per_position']); print(d['active_lanes'])"
ncu --set full --clock-control none --import-source on -k regex:k_map -s 2 -c 1 -f -o gpurun_out/prof_map_r2i python tools/profile_map.py 8192 3 > gpurun_out/r2i_ncu_map.log 2>&1; echo ncu_map=$?
ncu --set full --clock-control none --import-source on -k regex:k_models -s 1 -c 1 -f -o gpurun_out/prof_models_r2i python tools/profile_models.py 10000 2 > gpurun_out/r2i_ncu_models.log 2>&1; echo ncu_models=$?
ncu --set full --clock-control none --import-source on -k regex:k_points -s 4 -c 1 -f -o gpurun_out/prof_points_shuffled_r2i python tools/profile_points.py 2 > gpurun_out/r2i_ncu_points.log 2>&1; echo ncu_points=$?
