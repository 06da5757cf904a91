set -x
timeout 300 python -m pytest tests/test_dwalk.py -x -q 2>&1 | tail -3 > gpurun_out/r2_u_tests.log
timeout 300 python tools/tune_dmma.py c4 200000 0,9,0 > gpurun_out/r2_u_tune_c4.jsonl 2> gpurun_out/r2_u_tune.err
timeout 300 python tools/tune_dmma.py c4 25000 0,9 >> gpurun_out/r2_u_tune_c4.jsonl 2>> gpurun_out/r2_u_tune.err
