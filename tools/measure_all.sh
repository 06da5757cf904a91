#!/bin/bash
# The measurement pass behind profiles/r2_x_*: run on a B200 box as
#     gpurun --timeout 2400 -- 'bash tools/measure_all.sh'
# then, back in the container, `python tools/ncu_traffic.py gpurun_out/r2_x_traffic_<cfg>.csv ... --kernel-rev <tag of phb_version()>`
# for every config (profiles/traffic.json) and `python tools/ncu_summary.py gpurun_out/r2_x_dwalk_c4.ncu-rep --out profiles/...`.
# Kernel revision tags (phb_version()): the walks' captures are r2w, the level-batched tensor-core kernels' (c5) r2z, e.g.
#     python tools/ncu_traffic.py profiles/r2_z_traffic_c5.csv --kernels 'k_dmma_(lower|upper|cherry)' --per k_dmma_pack --key c5:auto --patterns 1000000 --kernel-rev r2z
# Numbers printed by the runs under ncu are never bench values; every step is bounded by `timeout`.
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for cfg in c2 c3 c4 c5 c2_1m; do
  timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_x_traffic_$cfg.csv python bench.py --config $cfg --steps 1 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/r2_x_${cfg}_under_ncu.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dwalk_p --launch-skip 6 -c 2 -o gpurun_out/r2_x_dwalk_c4 python bench.py --config c4 --steps 1 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/r2_x_c4_full.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_x_tests.log
timeout 600 python bench.py > gpurun_out/r2_x_bench.json 2> gpurun_out/r2_x_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_x_ref.json 2> gpurun_out/r2_x_ref.err
