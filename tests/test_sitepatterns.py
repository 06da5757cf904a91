"""Site-pattern compression: integer work, BIT-EXACT including the reference's pattern order.

Golden: tests/golden/sitepatterns.npz holds the SitePattern the unmodified reference built (new_SitePattern2,
sitepattern.c:186-251) for its own fluA alignment and for a synthetic alignment with ambiguity codes that takes its hash table
through five growth steps.  CPU: the oracle restatement against it.  GPU: phb_compress_patterns against both.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import GOLDEN


def _golden():
    return dict(np.load(os.path.join(GOLDEN, "sitepatterns.npz")))


@pytest.mark.parametrize("case", ["fluA", "synth"])
def test_oracle_reproduces_reference_sitepattern(case):
    z = _golden()
    pat, w, smap = O.compress_patterns(z[case + "_alignment"])
    assert np.array_equal(pat, z[case + "_patterns"]) and np.array_equal(w, z[case + "_weights"])
    aln = z[case + "_alignment"]
    assert np.array_equal(pat[:, smap], aln) and w.sum() == aln.shape[1]


def test_oracle_edge_cases():
    one = np.array([[2], [0], [3]], np.uint8)  # a single site
    pat, w, smap = O.compress_patterns(one)
    assert np.array_equal(pat, one) and w.tolist() == [1.0] and smap.tolist() == [0]
    const = np.zeros((4, 1000), np.uint8)  # every column identical
    pat, w, smap = O.compress_patterns(const)
    assert pat.shape == (4, 1) and w.tolist() == [1000.0] and not smap.any()
    rng = np.random.default_rng(1)
    uniq = rng.integers(0, 4, size=(40, 3000)).astype(np.uint8)  # (almost surely) all unique
    pat, w, smap = O.compress_patterns(uniq)
    assert pat.shape[1] == len({c.tobytes() for c in uniq.T}) and np.array_equal(pat[:, smap], uniq)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["fluA", "synth"])
def test_gpu_matches_reference_sitepattern_bit_exactly(case):
    import physher_b200 as phb

    z = _golden()
    pat, w, smap = phb.compress_patterns(z[case + "_alignment"])
    assert pat.dtype == np.uint8 and np.array_equal(pat, z[case + "_patterns"]), "patterns, in the reference's order"
    assert np.array_equal(w, z[case + "_weights"])
    assert np.array_equal(pat[:, smap], z[case + "_alignment"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1, 0), (3, 1, 0), (4, 1000, 0), (40, 3000, 1), (1000, 20000, 2), (64, 400000, 3), (5, 1000003, 4)])
def test_gpu_matches_oracle(shape):
    """Single site, constant alignment, all-unique columns, BASELINE-shaped tall alignments, heavy duplication with > 10 table growths."""
    import physher_b200 as phb

    T, n, mode = shape
    rng = np.random.default_rng(100 + T + n)
    if mode == 0:
        aln = np.zeros((T, n), np.uint8) + (2 if n == 1 else 0)
    elif mode in (1, 2):
        aln = rng.integers(0, 4, size=(T, n)).astype(np.uint8)
    else:  # few taxa, few states: most columns are repeats
        aln = rng.choice(np.array([0, 1, 2, 3, 15, 17], np.uint8), p=[0.4, 0.3, 0.15, 0.1, 0.03, 0.02], size=(T, n))
        if mode == 3:
            aln[8:] = aln[:1]  # 64 taxa but only 8 independent rows
    want = O.compress_patterns(aln)
    got = phb.compress_patterns(aln)
    for g, x in zip(got, want):
        assert np.array_equal(g, x)
    assert got[1].sum() == n
