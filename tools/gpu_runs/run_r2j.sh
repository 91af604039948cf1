# r2j: flat Newton loop, uniform-grid specialisation of k_models; resident blocks 6/7/8 and segment lengths
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/profile_map.py 8192 4 2>&1 | tail -1
for lib in "" variants/lib_mb7.so variants/lib_mb8.so; do for seg in 0 64 32; do echo "lib=$lib seg=$seg"; MLM_LIB=$lib MLM_SEGMENT=$seg python tools/profile_models.py 10000 3 2>&1 | tail -1; done; done
python tools/profile_points.py 3 2>&1 | tail -4 | grep "rep 2"
python tools/count_events_models.py > gpurun_out/r2j_events_c5.json 2>gpurun_out/r2j_events_c5.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_events_c5.json')); print(d['segment'], d['per_position']); print(d['active_lanes'])"
