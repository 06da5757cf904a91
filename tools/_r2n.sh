set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for cfg in c2 c3 c4 c5 c2_1m; do
  timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_n_traffic_$cfg.csv python bench.py --config $cfg --steps 1 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/r2_n_${cfg}_under_ncu.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dwalk_p --launch-skip 6 -c 2 -o gpurun_out/r2_n_dwalk_c4 python bench.py --config c4 --steps 1 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/r2_n_c4_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dmma_ --launch-skip 150 -c 12 -o gpurun_out/r2_n_dmma_c5 python bench.py --config c5 --steps 1 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/r2_n_c5_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nuc4_walk --launch-skip 3 -c 1 -o gpurun_out/r2_n_nuc4_c2 python bench.py --config c2 --steps 1 --warmup 3 --sub none --no-cpu-baseline > gpurun_out/r2_n_c2_full.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_n_bench.json 2> gpurun_out/r2_n_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_n_ref.json 2> gpurun_out/r2_n_ref.err
ls -la gpurun_out | head -40
