#!/usr/bin/env python
"""Per-region and per-instruction stall samples of the fused 4-state walk from an .ncu-rep (source page, first profiled launch):

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [--out profiles/name.md] [--title "..."]

Regions are found from landmarks in the SASS: the post-order loop ends at the first hot CTA barrier, the pre-order loop at the
last one; the loops start at their first hot mbarrier wait.
"""
import argparse
import csv
import io
import subprocess
from collections import Counter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--out")
    ap.add_argument("--title", default="")
    ap.add_argument("--ops", type=float, default=0.0, help="op executions per warp summed over all warps (for per-op instruction counts)")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
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
    data = []
    for r in blocks[0]:
        try:
            data.append((r[ix["Source"]], int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), r))
        except ValueError:
            pass
    tot, ninst = sum(d[1] for d in data), sum(d[2] for d in data)
    stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    hot = max(d[2] for d in data) / 8
    bars = [i for i, d in enumerate(data) if "BAR.SYNC" in d[0] and d[2] > hot]
    waits = [i for i, d in enumerate(data) if "SYNCS.PHASECHK" in d[0]]
    post = (min(i for i in waits if data[i][2] > hot), bars[0] + 5)
    pre = (min(i for i in waits if i > post[1]), bars[-1] + 5)
    lines = [f"# {a.title or a.rep}", "", f"source: `{a.rep}`, first profiled launch: {tot} stall samples, {ninst} executed warp instructions", ""]

    def region(lo, hi, name):
        s, n = sum(d[1] for d in data[lo:hi]), sum(d[2] for d in data[lo:hi])
        st, mix = Counter(), Counter()
        for d in data[lo:hi]:
            for k in stall_cols:
                v = d[3][ix[k]]
                if v:
                    st[k[6:]] += int(v)
            t = d[0].split()
            if t:
                mix[(t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]] += d[2]
        per_op = f", {n / a.ops:.0f} per op" if a.ops else ""
        lines.append(f"* **{name}**: {100 * s / tot:.1f} % of the samples, {100 * n / ninst:.1f} % of the instructions{per_op}; stalls: "
                     + ", ".join(f"{k} {100 * v / max(s, 1):.0f} %" for k, v in st.most_common(6)) + "; mix: "
                     + ", ".join(f"{k} {100 * v / max(n, 1):.0f} %" for k, v in mix.most_common(8)))

    region(*post, "post-order loop")
    region(*pre, "pre-order loop")
    region(post[1], pre[0], "root integration and set-up between the loops")
    lines += ["", "hottest instructions (share of all samples, executions, SASS, top stall reasons):", ""]
    for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][1])[:12]):
        d = data[i]
        s2 = sorted(((int(d[3][ix[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
        lines.append(f"* `{d[0].strip()[:64]}` — {100 * d[1] / tot:.1f} %, {d[2]} executions; " + ", ".join(f"{k} {v}" for v, k in s2))
    text = "\n".join(lines) + "\n"
    if a.out:
        open(a.out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
