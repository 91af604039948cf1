# r2z (b): the default bench line, the reference arm, and the ncu launch list of the same command (short form)
s=$(date +%s); python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo bench=$? $(( $(date +%s) - s )) s
s=$(date +%s); python bench.py --impl reference > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo ref=$? $(( $(date +%s) - s )) s
python bench.py --steps 2 --warmup 1 --no-other --no-cpu > gpurun_out/r2z_bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-other --no-cpu > gpurun_out/r2z_ncu_launches.log 2>&1; echo launches=$?
python -c "import __graft_entry__ as g; g.smoke()"
