mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/rm_launches.csv python tools/bench_scaled_aa.py 400 50000 > gpurun_out/rm_ncu.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/rm_launches.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e6 if u in('ns','nsecond') else v/1e3 if u in ('us','usecond') else v
    k=r[ki].split('(')[0][:60]; agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print(f"{t:10.2f} ms {n:6d} {k}")
P
