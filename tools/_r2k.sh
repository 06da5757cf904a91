set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_k_tests.log
python tools/tune_dmma.py c4 200000 0,9,14,12,0 > gpurun_out/r2_k_tune_c4.jsonl 2> gpurun_out/r2_k_tune.err
python tools/tune_dmma.py c4 25000 0,9,12,0 >> gpurun_out/r2_k_tune_c4.jsonl 2>> gpurun_out/r2_k_tune.err
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_k_bench_c4.json 2> gpurun_out/r2_k_bench_c4.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_k_traffic_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_k_c4_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dwalk_p --launch-skip 6 -c 2 -o gpurun_out/r2_k_dwalk_c4 python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_k_c4_full.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_dwalk.py -x -q -k "oracle or extreme" > gpurun_out/r2_k_memcheck.log 2>&1
compute-sanitizer --tool racecheck python -m pytest tests/test_dwalk.py -x -q -k "T24-P63 or T3-P17 or balanced" > gpurun_out/r2_k_racecheck.log 2>&1
ls -la gpurun_out
