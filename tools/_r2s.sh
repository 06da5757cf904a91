set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_s_bench_n8.json 2> gpurun_out/r2_s_bench_n8.err
tail -3 gpurun_out/r2_s_bench_n8.err
