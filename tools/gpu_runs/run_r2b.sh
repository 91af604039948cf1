TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $TR --nproc-per-node 8 --master-port 29518 tools/time_gather.py --gathers p2p --shares 0.24,0.27 --strips 16,32,64 --out gpurun_out/r2b_gather_n8.json 2>&1 | grep '^{' 
timeout 150 $TR --nproc-per-node 4 --master-port 29514 tools/time_gather.py --gathers p2p --shares 0.25,0.28 --strips 16,32 --out gpurun_out/r2b_gather_n4.json 2>&1 | grep '^{'
timeout 300 $TR --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err; echo bench8=$?; tail -3 gpurun_out/r2b_bench_n8.err
timeout 300 $TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo bench2=$?; tail -3 gpurun_out/r2b_bench_n2.err
