set -x
timeout 300 python -m pytest tests/test_dwalk.py -x -q 2>&1 | tail -5 > gpurun_out/r2_m_tests.log
timeout 200 python tools/tune_dmma.py c4 200000 0,9,0 > gpurun_out/r2_m_tune_c4.jsonl 2> gpurun_out/r2_m_tune.err
timeout 240 compute-sanitizer --tool memcheck python -m pytest tests/test_dwalk.py -x -q -k "T24-P63-C4 or T3-P17-C4 or balanced" > gpurun_out/r2_m_memcheck.log 2>&1
timeout 240 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_dwalk.py -x -q -k "T24-P63-C4-auto or T3-P17-C4-spill8" > gpurun_out/r2_m_racecheck.log 2>&1
echo finished
