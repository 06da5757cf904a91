set -x
timeout 300 python -m pytest tests/test_dwalk.py -x -q 2>&1 | tail -3 > gpurun_out/r2_w_tests.log
timeout 300 python tools/tune_dmma.py c4 200000 0,0 > gpurun_out/r2_w_tune_c4.jsonl 2> gpurun_out/r2_w_tune.err
timeout 300 python tools/tune_dmma.py c4 25000 0 >> gpurun_out/r2_w_tune_c4.jsonl 2>> gpurun_out/r2_w_tune.err
