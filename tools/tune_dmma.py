#!/usr/bin/env python
"""Geometry sweep of the tensor-core message kernels (phb_dmma.cu MsgCfg, PHB_OPT_TUNE) on the BASELINE workloads:

    python tools/tune_dmma.py c4 [patterns] [variants, e.g. 0,9,14] > profiles/r2_tune_c4.jsonl

20 states: variant 0 is the whole-tree walk (phb_dwalk.cu), 9 the level-batched message kernels it replaced (their own geometry
variants 1-6 are reachable as 9 only through phb_dmma.cu's table), 11-14 the walk's test geometries (one slot / 8 warps / both / 4 warps).

One JSON line per variant: ms per lnL + gradient evaluation (kernel sequence timed with CUDA events inside the library) and the
largest deviation of its gradient from variant 0's (the variants must agree to rounding)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import physher_b200 as phb  # noqa: E402
from physher_b200.treelikelihood import OPT_TIMING, OPT_TUNE  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    cfg = dict(bench.CONFIGS[name])
    if len(sys.argv) > 2:
        cfg["patterns"] = int(sys.argv[2])
    nvar = {20: 7, 61: 3}[cfg["states"]]
    variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(nvar)) + [0]
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    tlk = phb.SingleTreeLikelihood(topo.left, topo.right, topo.root, cfg["states"], cfg["cats"], cfg["patterns"], use_tip_states=True, device=0)
    tlk.set_tip_states(patterns)
    tlk.set_pattern_weights(weights)
    tlk.set_eigen(m.evec, m.eval, m.ivec)
    tlk.set_frequencies(m.freqs)
    tlk.set_site_model(rates, props)
    base = None
    for v in variants:
        tlk.set_option(OPT_TUNE, v)
        tlk.set_branch_lengths(bl)
        g = tlk.gradient().copy()
        for _ in range(2):
            tlk.set_branch_lengths(bl * 1.001)
            tlk.gradient()
        tlk.set_option(OPT_TIMING, 1)
        t0 = time.perf_counter()
        K = 5
        for i in range(K):
            tlk.set_branch_lengths(bl * (1.0 + 1e-3 * i))
            tlk.gradient()
        wall = (time.perf_counter() - t0) * 1e3 / K
        ms, n = tlk.kernel_time()
        tlk.set_option(OPT_TIMING, 0)
        if base is None:
            base = g
        dev = float(np.max(np.abs(g - base) / np.maximum(np.abs(base), 1e-6 * np.abs(base).max())))
        print(json.dumps({"config": name, "patterns": cfg["patterns"], "variant": v, "kernel_ms": ms / max(n, 1), "wall_ms": wall, "grad_dev_vs_variant0": dev}), flush=True)
    tlk.close()


if __name__ == "__main__":
    main()
