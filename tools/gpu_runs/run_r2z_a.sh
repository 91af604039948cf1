# r2z (a): evidence pass -- tests, event counters, ncu --set full of k_map<4>, k_models<4>, k_points_direct<4>
python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/count_events.py C4 > gpurun_out/r2z_events_c4.json 2>gpurun_out/r2z_events.err || tail -3 gpurun_out/r2z_events.err
python tools/count_events_models.py > gpurun_out/r2z_events_c5.json 2>>gpurun_out/r2z_events.err
python tools/profile_map.py 8192 4 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_map -s 2 -c 1 -f -o gpurun_out/prof_map_r2z python tools/profile_map.py 8192 3 > gpurun_out/r2z_ncu_map.log 2>&1; echo ncu_map=$?
ncu --set full --clock-control none --import-source on -k regex:k_models -s 1 -c 1 -f -o gpurun_out/prof_models_r2z python tools/profile_models.py 10000 2 > gpurun_out/r2z_ncu_models.log 2>&1; echo ncu_models=$?
ncu --set full --clock-control none --import-source on -k regex:k_points_direct -s 1 -c 1 -f -o gpurun_out/prof_points_direct_r2z python tools/unordered_once.py > gpurun_out/r2z_ncu_points.log 2>&1; echo ncu_points=$?
ncu --set full --clock-control none --import-source on -k regex:k_points_list -s 1 -c 1 -f -o gpurun_out/prof_points_list_r2z python tools/unordered_once.py > gpurun_out/r2z_ncu_points2.log 2>&1; echo ncu_points_list=$?
