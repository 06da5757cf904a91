#!/usr/bin/env python
"""Hottest SASS instructions (stall samples) and the instruction mix of each profiled launch of an .ncu-rep, any kernel:

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [--top 40] [--launch 1]
"""
import argparse
import csv
import io
import subprocess
from collections import Counter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--listing", action="store_true", help="print the whole SASS with samples and executions")
    ap.add_argument("--kernel", default="", help="regex on the kernel name (ncu --kernel-name regex:...); --launch then counts within the matches")
    a = ap.parse_args()
    cmd = ["ncu", "-i", a.rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + a.kernel] if a.kernel else [])
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, h = [], None
    for r in rows:
        if "Source" in r and "Instructions Executed" in r:
            h = r
            blocks.append([])
            continue
        if h is not None and len(r) >= len(h):
            blocks[-1].append(r)
    ix = {k: h.index(k) for k in h}
    stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    data = []
    for r in blocks[a.launch]:
        try:
            data.append((r[ix["Source"]], int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), r))
        except ValueError:
            pass
    tot, ninst = sum(d[1] for d in data), sum(d[2] for d in data)
    print(f"launch {a.launch}: {tot} samples, {ninst} executed warp instructions, {len(data)} SASS lines")
    st = Counter()
    for d in data:
        for k in stall_cols:
            v = d[3][ix[k]]
            if v:
                st[k[6:]] += int(v)
    print("stalls:", ", ".join(f"{k} {100 * v / max(tot, 1):.0f}%" for k, v in st.most_common(9)))
    if a.listing:
        for i, d in enumerate(data):
            tops = sorted(((int(d[3][ix[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
            print(f"{i:5d} {d[1]:7d} {d[2]:10d}  {d[0][:90]:90s} {' '.join(f'{n}:{v}' for v, n in tops if v)}")
        return
    for i in sorted(range(len(data)), key=lambda i: -data[i][1])[: a.top]:
        d = data[i]
        tops = sorted(((int(d[3][ix[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:3]
        print(f"{i:5d} {100 * d[1] / max(tot, 1):5.1f}% {d[2]:10d}  {d[0][:80]:80s} {' '.join(f'{n}:{v}' for v, n in tops if v)}")


if __name__ == "__main__":
    main()
