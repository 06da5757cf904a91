"""phb_group (csrc/phb_group.c): the pattern-sharded scheme of SURVEY.md 8e driven from ONE host thread through the C ABI -- what a
single-process C host like physher itself uses.  With one visible GPU the shards share the device; with more they spread out."""
import numpy as np
import pytest

import physher_b200 as phb
from oracle import oracle as O
from tests.test_gpu_parity import _synthetic_problem
from tests.util import RTOL, grad_err, rel_err

pytestmark = pytest.mark.gpu


def _devices(n):
    have = max(phb.device_count(), 1)
    return [i % have for i in range(n)]


@pytest.mark.parametrize("shape", [(40, 1001, 4, 4, 3), (12, 333, 20, 2, 2), (8, 130, 61, 1, 4), (10, 77, 5, 2, 2)], ids=lambda s: "S%d-G%d" % (s[2], s[4]))
def test_group_matches_oracle_and_single_device(shape):
    T, P, S, C, G = shape
    pb = _synthetic_problem(T, P, S, C, seed=9100 + S)
    want = O.evaluate(pb)
    one = phb.SingleTreeLikelihood.from_problem(pb)
    lnl_one, g_one = one.calculate(), one.gradient()
    one.close()
    grp = phb.TreeLikelihoodGroup.from_problem(pb, _devices(G))
    assert grp.size() == G
    edges = [grp.shard_range(s) for s in range(G)]
    assert edges[0][0] == 0 and edges[-1][1] == P and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert rel_err(grp.calculate(), want["lnl"]) < RTOL
    lnl, g = grp.gradient()
    assert rel_err(lnl, want["lnl"]) < RTOL and rel_err(lnl, lnl_one) < 1e-12
    assert grad_err(g, want["grad"]) < RTOL and grad_err(g, g_one) < 1e-11
    assert g[pb.root] == 0.0 and g[pb.right[pb.root]] == 0.0
    # new branch lengths reach every shard
    pb.bl = pb.bl * 0.7
    grp.set_branch_lengths(pb.bl)
    want = O.evaluate(pb)
    lnl, g = grp.gradient()
    assert rel_err(lnl, want["lnl"]) < RTOL and grad_err(g, want["grad"]) < RTOL
    # rooted request: the root's right child keeps its gradient
    grp.set_option(phb.treelikelihood.OPT_UNROOTED, 0)
    pb.unrooted = False
    _, g = grp.gradient()
    assert grad_err(g, O.evaluate(pb)["grad"]) < RTOL
    grp.close()


def test_group_tip_partials_are_sliced_like_states():
    pb = _synthetic_problem(15, 400, 4, 4, seed=9200)
    tp = np.zeros((pb.ntips, pb.npatterns, 4))
    known = pb.tip_states < 4
    tp[known, pb.tip_states[known]] = 1.0
    tp[~known] = 1.0
    want = O.evaluate(pb)
    pb.tip_partials, pb.use_tip_states = tp, False
    grp = phb.TreeLikelihoodGroup.from_problem(pb, _devices(3))
    lnl, g = grp.gradient()
    assert rel_err(lnl, want["lnl"]) < RTOL and grad_err(g, want["grad"]) < RTOL
    grp.close()


def test_group_underflow_switches_rescaling_on_every_shard():
    """a deep tree underflows without rescaling: the REDUCED lnL is -inf, every shard switches (treelikelihood.c:1496-1519)"""
    from tests.test_sharded import _problem

    pb = _problem(T=600, P=24, C=1, seed=3, deep=True)
    assert np.isinf(O.evaluate(pb, gradient=False)["lnl"])
    pb.scale = True
    want = O.evaluate(pb)
    pb.scale = False
    grp = phb.TreeLikelihoodGroup.from_problem(pb, _devices(2))
    assert not grp.rescaling()
    lnl, g = grp.gradient()
    assert grp.rescaling()
    assert np.isfinite(lnl) and rel_err(lnl, want["lnl"]) < RTOL and grad_err(g, want["grad"]) < 1e-9
    grp.close()


def test_group_nan_fills_the_gradient():
    pb = _synthetic_problem(9, 64, 4, 2, seed=9300)
    pb.bl = pb.bl.copy()
    grp = phb.TreeLikelihoodGroup.from_problem(pb, _devices(2))
    bad = pb.bl.copy()
    bad[3] = np.nan
    grp.set_branch_lengths(bad)
    lnl, g = grp.gradient()
    assert np.isnan(lnl) and np.isnan(g).all()
    grp.set_branch_lengths(pb.bl)
    lnl, g = grp.gradient()
    assert rel_err(lnl, O.evaluate(pb)["lnl"]) < RTOL and np.isfinite(g).all()
    grp.close()
