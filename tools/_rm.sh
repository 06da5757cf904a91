set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/rm_tests.log
timeout 300 python tools/bench_scaled_aa.py 400 50000 > gpurun_out/rm_scaled_aa.log 2>&1
timeout 300 python tools/bench_scaled_aa.py 400 50000 c5 > gpurun_out/rm_scaled_c5.log 2>&1
for f in gpurun_out/rm_tests.log gpurun_out/rm_scaled_aa.log gpurun_out/rm_scaled_c5.log; do tail -n 3 $f; done
