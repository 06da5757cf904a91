#!/usr/bin/env python
"""lnL + gradient time of an amino-acid workload that needs rescaling (more taxa than FP64 can hold unscaled, ~230 at 20 states):

    python tools/bench_scaled_aa.py [taxa] [patterns] [config, default c4]

prints one JSON line: kernel ms per evaluation with rescaling on (what the library switches to by itself after a -inf) and, for
scale, the unscaled time of the same shape (its lnL is -inf: only the time means anything)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import physher_b200 as phb  # noqa: E402
from physher_b200.treelikelihood import OPT_TIMING  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    base = sys.argv[3] if len(sys.argv) > 3 else "c4"
    cfg = dict(bench.CONFIGS[base], taxa=T, patterns=P)
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    tlk = phb.SingleTreeLikelihood(topo.left, topo.right, topo.root, cfg["states"], cfg["cats"], P, use_tip_states=True, device=0)
    tlk.set_tip_states(patterns)
    tlk.set_pattern_weights(weights)
    tlk.set_eigen(m.evec, m.eval, m.ivec)
    tlk.set_frequencies(m.freqs)
    tlk.set_site_model(rates, props)
    out = {"config": base, "taxa": T, "patterns": P}
    for scaled in (False, True):
        tlk.use_rescaling(scaled)
        tlk.set_branch_lengths(bl)
        lnl = tlk.calculate()
        if not scaled and np.isfinite(lnl):
            out["note"] = "the unscaled evaluation did not underflow"
        tlk.use_rescaling(scaled)  # a -inf switches rescaling on by itself: force the mode being timed
        for _ in range(2):
            tlk.set_branch_lengths(bl * 1.001)
            tlk.gradient()
            tlk.use_rescaling(scaled)
        tlk.set_option(OPT_TIMING, 1)
        K = 3
        t0 = time.perf_counter()
        for i in range(K):
            tlk.set_branch_lengths(bl * (1.0 + 1e-3 * i))
            tlk.gradient()
            tlk.use_rescaling(scaled)
        wall = (time.perf_counter() - t0) * 1e3 / K
        ms, n = tlk.kernel_time()
        tlk.set_option(OPT_TIMING, 0)
        out["scaled" if scaled else "unscaled"] = {"kernel_ms": ms / max(n, 1), "wall_ms": wall, "lnl": float(tlk.calculate()), "rescaling": bool(tlk.rescaling())}
    print(json.dumps(out))
    tlk.close()


if __name__ == "__main__":
    main()
