#!/usr/bin/env python
"""Wall-clock cost of every reference-facing entry point of the hot path on one workload, next to the plain lnL + gradient evaluation:

    python tools/api_sweep.py <config> [patterns] [taxa]

One JSON line {entry point: ms}.  The point is to find cliffs: an entry point that costs a large multiple of an evaluation runs on kernels
nobody tuned (that is how the 13 x slower rescaled path and the 30 x slower substitution-model sweep at 20 / 61 states were found)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import physher_b200 as phb  # noqa: E402
from physher_b200.treelikelihood import OPT_INCREMENTAL  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    cfg = dict(bench.CONFIGS[name])
    if len(sys.argv) > 2:
        cfg["patterns"] = int(sys.argv[2])
    if len(sys.argv) > 3:
        cfg["taxa"] = int(sys.argv[3])
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    T, P, S, C = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["cats"]
    N = 2 * T - 1

    def make(tip_states=True):
        tlk = phb.SingleTreeLikelihood(topo.left, topo.right, topo.root, S, C, P, use_tip_states=tip_states, device=0)
        if tip_states:
            tlk.set_tip_states(patterns)
        else:
            tlk.set_tip_partials(np.eye(S + 1)[np.minimum(patterns, S)][:, :, :S] + (patterns >= S)[:, :, None])
        tlk.set_pattern_weights(weights)
        tlk.set_eigen(m.evec, m.eval, m.ivec)
        tlk.set_frequencies(m.freqs)
        tlk.set_site_model(rates, props)
        tlk.set_branch_lengths(bl)
        return tlk

    out = {"config": bench.workload_name(cfg)}
    k = [0]

    def timed(tlk, fn, reps=3, fresh=True):
        fn()
        tlk.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if fresh:
                k[0] += 1
                tlk.set_branch_lengths(bl * (1.0 + 1e-4 * k[0]))
            fn()
        tlk.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps

    tlk = make()
    out["calculate"] = timed(tlk, tlk.calculate)
    out["gradient"] = timed(tlk, tlk.gradient)
    out["gradient + cat_branch_gradient"] = timed(tlk, lambda: (tlk.gradient(), tlk.cat_branch_gradient()))
    out["gradient + root_frequency_gradient"] = timed(tlk, lambda: (tlk.gradient(), tlk.root_frequency_gradient()))
    out["gradient + category_gradient"] = timed(tlk, lambda: (tlk.gradient(), tlk.category_gradient()))
    out["gradient_batch(4) / 4"] = timed(tlk, lambda: tlk.gradient_batch(np.stack([bl * (1 + 0.001 * i) for i in range(4)])), reps=2, fresh=False) / 4
    M = np.random.default_rng(1).normal(size=(2, N, C, S, S))
    out["matrix_gradient(2 sets)"] = timed(tlk, lambda: tlk.matrix_gradient(M), reps=2)
    out["gradient + get_partials(1 node)"] = timed(tlk, lambda: (tlk.gradient(), tlk.get_partials(T + 1)), reps=2)
    tlk.use_rescaling(True)
    out["gradient, rescaled"] = timed(tlk, lambda: (tlk.use_rescaling(True), tlk.gradient()))
    tlk.use_rescaling(False)
    tlk.set_option(OPT_INCREMENTAL, 1)
    out["incremental: update_uppers"] = timed(tlk, tlk.update_uppers, reps=2)
    node = int(topo.left[topo.root])
    out["incremental: calculate_branch(8 lengths)"] = timed(tlk, lambda: tlk.calculate_branch(node, np.linspace(0.01, 0.2, 8)), fresh=False)

    def one_branch():
        k[0] += 1
        tlk.set_branch_length(node, float(bl[node] * (1.0 + 1e-4 * k[0])))
        return tlk.calculate()

    out["incremental: set_branch_length + calculate"] = timed(tlk, one_branch, fresh=False)

    def one_branch_gradient():
        k[0] += 1
        tlk.set_branch_length(node, float(bl[node] * (1.0 + 1e-4 * k[0])))
        return tlk.gradient()

    out["incremental: set_branch_length + gradient"] = timed(tlk, one_branch_gradient, reps=2, fresh=False)
    tlk.close()
    tlk = make(tip_states=False)
    out["gradient, 0/1 tip partials"] = timed(tlk, tlk.gradient)
    tlk.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
