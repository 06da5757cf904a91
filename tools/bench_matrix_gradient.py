#!/usr/bin/env python
"""Timing of the substitution-model parameter gradient request (phb_tlk_matrix_gradient, the node sweep of calculate_dlnl_dQ,
treelikelihood.c:2337-2583) on a bench.py workload:

    python tools/bench_matrix_gradient.py --config c2 [--patterns N] [--sets 8]

fused      4 states, unscaled: one launch of the fused walk with the per-branch transition statistics + one contraction per set
generic    the same request on the node-at-a-time kernels (materialised upper partials, one sweep per set)
gradient   one lnL + branch-gradient evaluation (the bench.py step), for scale
PHB_OPT_TIMING brackets the dominant kernel(s) with CUDA events; wall-clock is around the synchronous C-ABI call (it uploads
the `sets` matrix sets and returns the results to the host).  One JSON line.
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
from physher_b200.treelikelihood import OPT_TIMING  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(bench.CONFIGS))
    ap.add_argument("--patterns", type=int, default=0)
    ap.add_argument("--sets", type=int, default=8, help="matrix sets per request (GTR: 5 rates + 3 frequencies)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-generic", action="store_true")
    a = ap.parse_args()
    cfg = dict(bench.CONFIGS[a.config])
    if a.patterns:
        cfg["patterns"] = a.patterns
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    T, P, S, C = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["cats"]
    N = 2 * T - 1
    M = np.random.default_rng(5).normal(size=(a.sets, N, C, S, S))
    out = {"config": bench.workload_name(cfg), "nodes": N, "sets": a.sets}

    def make(kernels):
        tlk = phb.SingleTreeLikelihood(topo.left, topo.right, topo.root, S, C, P, use_tip_states=True, device=0)
        tlk.set_option(phb.treelikelihood.OPT_KERNELS, kernels)
        tlk.set_tip_states(patterns)
        tlk.set_pattern_weights(weights)
        tlk.set_eigen(m.evec, m.eval, m.ivec)
        tlk.set_frequencies(m.freqs)
        tlk.set_site_model(rates, props)
        tlk.set_branch_lengths(bl)
        return tlk

    def timed(tlk, fn, reps):
        fn()
        tlk.set_option(OPT_TIMING, 1)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        wall = (time.perf_counter() - t0) * 1e3 / reps
        ms, launches = tlk.kernel_time()
        tlk.set_option(OPT_TIMING, 0)
        return {"wall_ms": wall, "kernel_ms": ms / reps, "timed_launches_per_call": launches / reps}

    results = {}
    for name, kernels in (("fused", phb.KERNELS_AUTO), ("generic", phb.KERNELS_GENERIC)):
        if name == "generic" and a.skip_generic:
            continue
        tlk = make(kernels)
        holder = {}

        def req():
            tlk.update_all_nodes()
            holder["g"] = tlk.matrix_gradient(M)

        out[name] = timed(tlk, req, a.reps)
        results[name] = holder["g"]
        if name == "fused":
            def grad():
                tlk.update_all_nodes()
                tlk.gradient()

            out["gradient"] = timed(tlk, grad, a.reps)
        tlk.close()
    if len(results) == 2:
        ref = results["generic"]
        out["max_rel_diff_fused_vs_generic"] = float(np.max(np.abs(results["fused"] - ref) / np.maximum(np.abs(ref), 1e-6 * np.abs(ref).max())))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
