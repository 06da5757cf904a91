"""Shared helpers for the parity tests: golden fixtures -> oracle.Problem, tolerances."""
from __future__ import annotations

import glob
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: lnL and every branch gradient within 1e-10 relative in FP64.
# Gradient entries are compared relative to max(|g|, |g|_inf * 1e-6).
RTOL = 1e-10


AUX_FIXTURES = {"c1_time_tree", "sitepatterns", "dlnl_dq_gtr_g4", "branch_derivatives"}  # inputs / outputs of other rows (tests/test_time_tree.py), not tree-likelihood problems


def golden_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if n not in AUX_FIXTURES]


def load_golden(name: str):
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    pb = O.Problem(
        left=z["left"], right=z["right"], parent=z["parent"], root=int(z["root"]), nstate=int(z["nstate"]),
        tip_states=z["tip_states"], weights=z["weights"], freqs=z["freqs"], rates=z["rates"], props=z["props"], bl=z["bl"],
        use_tip_states=bool(z["use_tip_states"]), unrooted=bool(z["unrooted"]),
    )
    if "tip_partials" in z:
        pb.tip_partials = z["tip_partials"].astype(np.float64)
    if "evec" in z:
        pb.evec, pb.eval, pb.ivec = z["evec"], z["eval"], z["ivec"]
    else:
        pb.P_override, pb.dP_override = z["P_override"], z["dP_override"]
    pb.meta["time_elapsed"] = z["time_elapsed"]
    return pb, z


def rel_err(a: float, b: float) -> float:
    return abs(a - b) / max(abs(b), 1e-300)


# Full-size shards (hundreds of taxa x thousands of patterns): a branch gradient is a sum over patterns of terms of both signs,
# and the reference forms its per-pattern weights as w_k / exp(pattern_lk[k]) (treelikelihood.c:3207-3210) -- a round trip through
# log and exp that costs |lnL_k| ulps (~3e-14 relative at lnL_k = -250).  A branch whose gradient nearly cancels (|g| below 1e-4 of the
# largest entry) carries that noise, of the reference's own making, as more than 1e-10 of ITS value: measured here as up to 8e-11
# between two pattern orders of the oracle itself (C3 shard, 1500 patterns).  Such entries are compared relative to FLOOR_LARGE * |g|_inf.
FLOOR_LARGE = 1e-4


def grad_err(g: np.ndarray, ref: np.ndarray, floor: float = 1e-6) -> float:
    scale = np.maximum(np.abs(ref), np.abs(ref).max() * floor)
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(g - ref) / scale))
