"""Deterministic synthetic trees and alignments for the parity tests and the benchmark.

Generator spec: topology by repeatedly joining two uniformly chosen live
lineages until one remains; branch lengths U(0.01, 0.1); sequences by an inherit-and-mutate
chain (taxon t copies a uniformly chosen earlier taxon, each site redrawn uniformly with
probability mu) so that columns are ~all unique and per-pattern lnL stays far above the
underflow range.  Everything is seeded; no file or network access.

Node ids follow physher's convention (tree.c:183-199): tips 0..T-1 in post-order encounter
order, internal nodes T..2T-2 in post-order, root = 2T-2.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

NUCLEOTIDES = "ACGT"
AMINO_ACIDS = "ACDEFGHIKLMNPQRSTVWY"


@dataclass
class Topology:
    """Rooted binary tree in physher id order."""

    left: np.ndarray  # [N] int32, -1 for tips
    right: np.ndarray  # [N] int32
    parent: np.ndarray  # [N] int32, -1 for the root
    root: int
    ntips: int

    @property
    def nnodes(self) -> int:
        return int(self.left.shape[0])


def random_topology(ntips: int, seed: int) -> Topology:
    rng = np.random.default_rng(seed)
    # build by joins on temporary labels, then relabel in post-order
    children = {}
    live = list(range(ntips))
    nxt = ntips
    while len(live) > 1:
        i, j = rng.choice(len(live), size=2, replace=False)
        a, b = live[i], live[j]
        for idx in sorted((i, j), reverse=True):
            live.pop(idx)
        children[nxt] = (a, b)
        live.append(nxt)
        nxt += 1
    root_tmp = live[0]
    return _relabel_postorder(children, root_tmp, ntips)


def caterpillar_topology(ntips: int) -> Topology:
    """Maximally unbalanced tree (depth T-1): worst case for level scheduling."""
    children = {}
    cur = 0
    nxt = ntips
    for t in range(1, ntips):
        children[nxt] = (cur, t)
        cur = nxt
        nxt += 1
    return _relabel_postorder(children, cur, ntips)


def balanced_topology(ntips: int) -> Topology:
    """Perfectly balanced tree (ntips must be a power of two): worst case for stack depth."""
    assert ntips & (ntips - 1) == 0
    children = {}
    layer = list(range(ntips))
    nxt = ntips
    while len(layer) > 1:
        new = []
        for i in range(0, len(layer), 2):
            children[nxt] = (layer[i], layer[i + 1])
            new.append(nxt)
            nxt += 1
        layer = new
    return _relabel_postorder(children, layer[0], ntips)


def _relabel_postorder(children: dict, root_tmp: int, ntips: int) -> Topology:
    n = 2 * ntips - 1
    left = np.full(n, -1, np.int32)
    right = np.full(n, -1, np.int32)
    parent = np.full(n, -1, np.int32)
    new_id = {}
    tip_counter = 0
    int_counter = ntips
    # iterative post-order
    stack = [(root_tmp, False)]
    while stack:
        node, done = stack.pop()
        if node not in children:
            new_id[node] = tip_counter
            tip_counter += 1
            continue
        if not done:
            stack.append((node, True))
            a, b = children[node]
            stack.append((b, False))
            stack.append((a, False))
        else:
            nid = int_counter
            int_counter += 1
            new_id[node] = nid
            a, b = children[node]
            left[nid], right[nid] = new_id[a], new_id[b]
            parent[new_id[a]] = nid
            parent[new_id[b]] = nid
    return Topology(left, right, parent, new_id[root_tmp], ntips)


def random_branch_lengths(topo: Topology, seed: int, lo: float = 0.01, hi: float = 0.1, unrooted: bool = True) -> np.ndarray:
    rng = np.random.default_rng(seed)
    bl = rng.uniform(lo, hi, size=topo.nnodes)
    bl[topo.root] = 0.0
    if unrooted:
        # physher's unrooted convention (treelikelihood.c:3249-3255): root's right child has length 0
        bl[topo.right[topo.root]] = 0.0
    return bl


def random_patterns(ntips: int, npatterns: int, nstate: int, mu: float, seed: int, unknown_frac: float = 0.0) -> np.ndarray:
    """uint8 [T][P] tip states by the inherit-and-mutate chain; `unknown_frac` of the entries are
    replaced by the unknown code `nstate` (gaps)."""
    rng = np.random.default_rng(seed)
    out = np.empty((ntips, npatterns), np.uint8)
    out[0] = rng.integers(0, nstate, size=npatterns, dtype=np.uint8)
    for t in range(1, ntips):
        src = out[rng.integers(0, t)]
        mut = rng.random(npatterns) < mu
        out[t] = np.where(mut, rng.integers(0, nstate, size=npatterns, dtype=np.uint8), src)
    if unknown_frac > 0:
        out[rng.random(out.shape) < unknown_frac] = nstate
    return out


def simulate_patterns(topo: Topology, bl: np.ndarray, npatterns: int, nstate: int, seed: int, unknown_frac: float = 0.0) -> np.ndarray:
    """uint8 [T][P] tip states evolved DOWN the given tree (so the data fit the tree and per-pattern lnL stays
    well above the double-precision underflow range): root state uniform; along a branch of length t a site keeps
    its parent's state with probability exp(-t S/(S-1)) and is redrawn uniformly otherwise (a Jukes-Cantor-like
    process on S states)."""
    rng = np.random.default_rng(seed)
    out = np.empty((topo.ntips, npatterns), np.uint8)
    state = {topo.root: rng.integers(0, nstate, size=npatterns, dtype=np.uint8)}
    stack = [topo.root]
    while stack:
        n = stack.pop()
        cur = state.pop(n)
        if topo.left[n] < 0:
            out[n] = cur
            continue
        for ch in (int(topo.left[n]), int(topo.right[n])):
            keep = rng.random(npatterns) < np.exp(-bl[ch] * nstate / (nstate - 1.0))
            state[ch] = np.where(keep, cur, rng.integers(0, nstate, size=npatterns, dtype=np.uint8))
            stack.append(ch)
    if unknown_frac > 0:
        out[rng.random(out.shape) < unknown_frac] = nstate
    return out


def to_newick(topo: Topology, bl: np.ndarray, names: list[str]) -> str:
    """Newick string whose physher parse reproduces `topo`'s ids (left child first)."""
    out = {}
    for n in range(topo.nnodes):  # ids are already a post-order
        if topo.left[n] < 0:
            out[n] = f"{names[n]}:{float(bl[n])!r}"
    order = [n for n in range(topo.ntips, topo.nnodes)]
    for n in order:
        s = f"({out.pop(int(topo.left[n]))},{out.pop(int(topo.right[n]))})"
        out[n] = s if n == topo.root else f"{s}:{float(bl[n])!r}"
    return out[topo.root] + ";"


def sequences_from_patterns(patterns: np.ndarray, alphabet: str, unknown: str = "-") -> list[str]:
    table = np.frombuffer((alphabet + unknown * (256 - len(alphabet))).encode(), dtype=np.uint8)
    return [table[row].tobytes().decode() for row in patterns]
