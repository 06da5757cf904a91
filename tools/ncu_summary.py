#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into the few numbers the roofline argument needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--out profiles/name.md] [--title "..."]
"""
import argparse
import csv
import io
import subprocess
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--out")
    ap.add_argument("--title", default="")
    ap.add_argument("--traffic", default="", help="KEY: record dram bytes per launch (mean over the captured launches) in profiles/traffic.json "
                                                    "under KEY, e.g. c2:auto -- bench.py reads it for roofline.traffic")
    a = ap.parse_args()
    lines = [f"# {a.title or a.rep}", "", f"source: `{a.rep}` (ncu --set full --clock-control none --import-source on)", ""]
    raw = ncu_csv(a.rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {row[i]} | {units[i]} |")
        stalls = [(h, float(row[i] or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        stalls.sort(key=lambda x: -x[1])
        lines += ["", "warp stall reasons (warps stalled per issue-active cycle): " +
                  ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for h, v in stalls[:8]), ""]
    src = ncu_csv(a.rep, "source")
    # the source page repeats a (kernel title, header, rows...) block per profiled launch
    c, tot, h, blocks = Counter(), 0, None, []
    for r in src:
        if "Instructions Executed" in r and "Source" in r:
            if h is not None and tot:
                blocks.append((c, tot))
            h, c, tot = r, Counter(), 0
            continue
        if h is None or len(r) < len(h):
            continue
        ia, isrc = h.index("Instructions Executed"), h.index("Source")
        t = r[isrc].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        try:
            n = int(r[ia] or 0)
        except ValueError:
            continue
        c[op] += n
        tot += n
    if h is not None and tot:
        blocks.append((c, tot))
    for k, (c, tot) in enumerate(blocks):
        lines += [f"SASS instruction mix, launch {k} (executed warp instructions): " + ", ".join(f"{op} {n/tot*100:.1f}%" for op, n in c.most_common(14)), ""]
    if a.traffic:
        import json
        import os

        def num(row, key):
            i = hdr.index(key)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
            return float(row[i].replace(",", "")) * scale

        per = [num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum") for r in raw[2:]]
        ms = [float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
              for r in raw[2:]]
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[a.traffic] = {"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per), "kernel": raw[2][hdr.index("Kernel Name")][:80],
                        "ncu_ms_per_launch": sum(ms) / len(ms), "source": os.path.basename(a.rep)}
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
    text = "\n".join(lines)
    if a.out:
        open(a.out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
