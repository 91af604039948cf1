set -x
python tools/count_events.py C4 > gpurun_out/r2a_events_c4.json 2> gpurun_out/r2a_events.err || tail -5 gpurun_out/r2a_events.err
python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo bench=$?; tail -3 gpurun_out/r2a_bench_n1.err
python tools/profile_map.py 8192 3 > gpurun_out/r2a_profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_map -s 2 -c 1 -f -o gpurun_out/prof_map_r2a python tools/profile_map.py 8192 3 > gpurun_out/r2a_ncu_full.log 2>&1; echo ncu=$?
python bench.py --steps 2 --warmup 1 --no-other --no-cpu > gpurun_out/r2a_bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-other --no-cpu > gpurun_out/r2a_ncu_launches.log 2>&1; echo launches=$?
