set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "group or sharded or clone or abi" 2>&1 | tail -5 > gpurun_out/r2_o_tests_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_o_bench_n2.json 2> gpurun_out/r2_o_bench_n2.err
tail -2 gpurun_out/r2_o_bench_n2.err
