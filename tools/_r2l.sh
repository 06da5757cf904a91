set -x
timeout 300 python tools/tune_dmma.py c4 200000 0,18,0 > gpurun_out/r2_r_tune_c4.jsonl 2> gpurun_out/r2_r_tune.err
timeout 300 python tools/tune_dmma.py c4 25000 0,18 >> gpurun_out/r2_r_tune_c4.jsonl 2>> gpurun_out/r2_r_tune.err
