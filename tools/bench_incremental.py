#!/usr/bin/env python
"""Timing of the resident-partials entry points (PHB_OPT_INCREMENTAL) beside a full evaluation, on a bench.py workload:

    python tools/bench_incremental.py --config c2 [--patterns N]

full        one lnL + gradient evaluation with every node dirty (the bench.py step)
inc_lnl     set_branch_length(random branch) + calculate(): only the ancestors of the branch are recomputed
inc_grad    set_branch_length(random branch) + gradient(): ancestors' lowers, the upper partials the change reaches, all reductions
branch_1/8  calculate_branch at 1 / 8 candidate lengths of a random branch (lnL, d/dt, d2/dt2) with uppers resident
walk        the serial-Brent access pattern: for every branch in post-order, 5 candidates through calculate_branch, keep the best
Wall-clock around synchronous C-ABI calls (each returns its result to the host); one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import physher_b200 as phb  # noqa: E402
from physher_b200.treelikelihood import OPT_INCREMENTAL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--patterns", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    cfg = dict(bench.CONFIGS[a.config])
    if a.patterns:
        cfg["patterns"] = a.patterns
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    T, P, S, C = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["cats"]
    N = 2 * T - 1
    tlk = phb.SingleTreeLikelihood(topo.left, topo.right, topo.root, S, C, P, use_tip_states=True, device=0)
    tlk.set_tip_states(patterns)
    tlk.set_pattern_weights(weights)
    tlk.set_eigen(m.evec, m.eval, m.ivec)
    tlk.set_frequencies(m.freqs)
    tlk.set_site_model(rates, props)
    tlk.set_branch_lengths(bl)
    rng = np.random.default_rng(3)
    branches = [n for n in range(N) if n != topo.root and n != topo.right[topo.root]]
    out = {"config": bench.workload_name(cfg), "nodes": N}

    def timed(fn, reps):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) * 1e3 / reps

    def full():
        tlk.update_all_nodes()
        tlk.gradient()

    out["full_ms"] = timed(full, 5)
    tlk.set_option(OPT_INCREMENTAL, 1)
    out["resident_full_ms"] = timed(full, 3)
    cur = bl.copy()

    def inc(kind):
        n = branches[int(rng.integers(len(branches)))]
        cur[n] *= float(rng.uniform(0.9, 1.1))
        tlk.set_branch_length(n, cur[n])
        return tlk.calculate() if kind == "lnl" else tlk.gradient()

    l0 = tlk.launch_count()
    out["inc_lnl_ms"] = timed(lambda: inc("lnl"), a.reps)
    out["inc_lnl_launches"] = (tlk.launch_count() - l0) / (a.reps + 1)
    out["inc_grad_ms"] = timed(lambda: inc("grad"), max(3, a.reps // 4))
    tlk.update_uppers()
    out["update_uppers_ms"] = timed(lambda: (tlk.update_all_nodes(), tlk.update_uppers()), 3)

    def branch(k):
        n = branches[int(rng.integers(len(branches)))]
        return tlk.calculate_branch(n, cur[n] * np.linspace(0.5, 2.0, k))

    out["branch_1_ms"] = timed(lambda: branch(1), a.reps)
    out["branch_8_ms"] = timed(lambda: branch(8), a.reps)
    # serial-Brent access pattern over a subset of branches (post-order = node id order for tips then internals here)
    sub = branches[:: max(1, len(branches) // 100)]
    t0 = time.perf_counter()
    for n in sub:
        lnl, d1, d2 = tlk.calculate_branch(n, cur[n] * np.array([0.5, 0.8, 1.0, 1.25, 2.0]))
        cur[n] *= [0.5, 0.8, 1.0, 1.25, 2.0][int(np.argmax(lnl))]
        tlk.set_branch_length(n, cur[n])
    final = tlk.calculate()
    out["walk_ms_per_branch"] = (time.perf_counter() - t0) * 1e3 / len(sub)
    out["walk_branches"] = len(sub)
    # consistency: the incrementally maintained lnL equals a full evaluation of the same lengths
    tlk.update_all_nodes()
    out["lnl_incremental"], out["lnl_full"] = final, tlk.calculate()
    out["lnl_rel_diff"] = abs(final - out["lnl_full"]) / abs(out["lnl_full"])
    print(json.dumps(out))
    tlk.close()


if __name__ == "__main__":
    main()
