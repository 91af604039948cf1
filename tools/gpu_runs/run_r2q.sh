# r2q: full default bench line at N=1 (with CPU baseline and parity blocks) and the reference arm
s=$(date +%s); python bench.py > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; echo bench=$? $(( $(date +%s) - s )) s
s=$(date +%s); python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err; echo ref=$? $(( $(date +%s) - s )) s
