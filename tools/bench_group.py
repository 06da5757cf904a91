#!/usr/bin/env python
"""Single-process multi-GPU evaluation through phb_group (csrc/phb_group.c) on a bench.py workload, weak scaling:

    python tools/bench_group.py --config c2 [--patterns-per-gpu N] [--gpus G]

every device gets `patterns-per-gpu` patterns; one host thread launches the evaluation on all shards, collects and sums.  Wall
clock around the synchronous C-ABI call (new branch lengths in, lnL + gradient out), best-effort comparison with G = 1.
One JSON line.
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


def run(cfg, ndev, reps):
    T, S, C = cfg["taxa"], cfg["states"], cfg["cats"]
    per = cfg["patterns"]
    parts = [bench.make_inputs(cfg, r) for r in range(ndev)]
    topo, bl, m, rates, props = parts[0][:5]
    patterns = np.concatenate([p[5] for p in parts], axis=1)
    weights = np.concatenate([p[6] for p in parts])
    grp = phb.TreeLikelihoodGroup(list(range(ndev)), topo.left, topo.right, topo.root, S, C, per * ndev, use_tip_states=True)
    lib, h = grp.lib, grp.h
    _bp, _dp = phb.treelikelihood._bp, phb.treelikelihood._dp
    a = np.ascontiguousarray(patterns, dtype=np.uint8)
    grp._check(lib.phb_group_set_tip_states(h, a.ctypes.data_as(_bp)))
    grp._check(lib.phb_group_set_pattern_weights(h, np.ascontiguousarray(weights).ctypes.data_as(_dp)))
    grp._check(lib.phb_group_set_eigen(h, m.evec.ctypes.data_as(_dp), m.eval.ctypes.data_as(_dp), m.ivec.ctypes.data_as(_dp)))
    grp._check(lib.phb_group_set_frequencies(h, m.freqs.ctypes.data_as(_dp)))
    grp._check(lib.phb_group_set_site_model(h, rates.ctypes.data_as(_dp), props.ctypes.data_as(_dp)))
    rng = np.random.default_rng(7)
    lnl = None
    for k in range(3 + reps):
        if k == 3:
            t0 = time.perf_counter()
        grp.set_branch_lengths(bl * rng.uniform(0.95, 1.05, bl.shape))
        lnl, g = grp.gradient()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    grp.close()
    N = 2 * T - 1
    return {"devices": ndev, "ms_per_eval": ms, "pattern_node_per_s": per * ndev * N / (ms * 1e-3), "lnl": lnl}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--patterns-per-gpu", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    cfg = dict(bench.CONFIGS[a.config])
    if a.patterns_per_gpu:
        cfg["patterns"] = a.patterns_per_gpu
    ndev = a.gpus or phb.device_count()
    out = {"config": bench.workload_name(cfg), "one": run(cfg, 1, a.reps)}
    if ndev > 1:
        out["all"] = run(cfg, ndev, a.reps)
        out["weak_scaling_efficiency"] = out["all"]["pattern_node_per_s"] / (ndev * out["one"]["pattern_node_per_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
