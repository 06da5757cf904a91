#!/usr/bin/env python
"""DRAM traffic per evaluation of a level-batched kernel sequence (tensor-core / node-at-a-time paths) from an ncu CSV log:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic_c4.csv python bench.py --config c4 --steps 1 --warmup 2 --no-cpu-baseline
    python tools/ncu_traffic.py gpurun_out/traffic_c4.csv --kernels 'k_dwalk_p' --per k_dmma_pack --key c4:auto --patterns 200000 --kernel-rev r2m

sums the bytes of every launch whose name matches --kernels, divides by the number of evaluations (= launches of the --per
kernel, which runs once per evaluation) and records it in profiles/traffic.json under --key (bench.py's roofline.traffic).
"""
import argparse
import csv
import json
import os
import re

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--kernels", required=True)
    ap.add_argument("--per", required=True)
    ap.add_argument("--key", required=True)
    ap.add_argument("--patterns", type=int, required=True, help="patterns per launch of the captured run (bench.py only trusts a capture of the same size)")
    ap.add_argument("--kernel-rev", required=True, help="kernel revision tag of phb_version() the capture was taken on")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv, errors="replace")))
    h = next(r for r in rows if "Kernel Name" in r and "Metric Name" in r)
    ik, im, iu, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
    pat, per = re.compile(a.kernels), re.compile(a.per)
    total = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    evals, launches = set(), set()
    for r in rows:
        if len(r) < len(h) or r is h or r[iid] == "ID":
            continue
        name, metric = r[ik], r[im]
        if per.search(name):
            evals.add(r[iid])
        if pat.search(name) and metric in total:
            total[metric] += float(r[iv].replace(",", "")) * UNITS.get(r[iu], 1.0)
            launches.add(r[iid])
    n = max(len(evals), 1)
    entry = {"dram_bytes_per_launch": (total["dram__bytes_read.sum"] + total["dram__bytes_write.sum"]) / n,
             "dram_read_bytes": total["dram__bytes_read.sum"] / n, "dram_write_bytes": total["dram__bytes_write.sum"] / n,
             "kernel": a.kernels, "launches": len(launches) / n, "evaluations": n, "ncu_ms_per_launch": total["gpu__time_duration.sum"] / n,
             "source": os.path.basename(a.csv), "note": "launch = the kernel sequence of one evaluation", "patterns": a.patterns, "kernel_rev": a.kernel_rev}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[a.key] = entry
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry))


if __name__ == "__main__":
    main()
