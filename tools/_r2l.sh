set -x
timeout 600 python tools/tune_dmma.py c5 1000000 0,3,4,0 > gpurun_out/r2_y_tune_c5.jsonl 2> gpurun_out/r2_y_tune.err
timeout 300 python tools/tune_dmma.py c5 125000 0,3,4 >> gpurun_out/r2_y_tune_c5.jsonl 2>> gpurun_out/r2_y_tune.err
