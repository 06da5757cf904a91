set -x
timeout 300 python -m pytest tests/test_dwalk.py -x -q 2>&1 | tail -3 > gpurun_out/r2_v_tests.log
timeout 300 python tools/tune_dmma.py c4 200000 0,19,9,0 > gpurun_out/r2_v_tune_c4.jsonl 2> gpurun_out/r2_v_tune.err
timeout 300 python tools/tune_dmma.py c4 25000 0,19,9 >> gpurun_out/r2_v_tune_c4.jsonl 2>> gpurun_out/r2_v_tune.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_v_traffic_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/r2_v_c4_under_ncu.log 2>&1
