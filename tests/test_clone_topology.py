"""clone_SingleTreeLikelihood (treelikelihood.c:1241-1395), SingleTreeLikelihood_update_three_nodes (:1754-1771) and topology moves
(NNI, nniopt.c:301-334): the boundary entry points the reference's parallel users and tree-search drivers rely on.

A clone must reproduce the source's results bit for bit and then live its own life; after a topology move the object must agree
with the oracle on the NEW tree (every kernel family, resident partials included) and refuse a broken tree without losing the old one.
"""
import copy

import numpy as np
import pytest

import physher_b200 as phb
from oracle import oracle as O
from tests.test_gpu_parity import _synthetic_problem
from tests.util import RTOL, grad_err, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _nni(pb, node):
    """swap child `left[node]` with the sibling of `node` (node-id convention untouched); returns a new Problem"""
    left, right, parent = pb.left.copy(), pb.right.copy(), pb.parent.copy()
    p = int(parent[node])
    assert p >= 0 and left[node] >= 0
    sib_is_left = left[p] != node
    s = int(left[p] if sib_is_left else right[p])
    a = int(left[node])
    left[node] = s
    if sib_is_left:
        left[p] = a
    else:
        right[p] = a
    parent[s], parent[a] = node, p
    q = copy.copy(pb)
    q.left, q.right, q.parent = left, right, parent
    return q


@pytest.mark.parametrize("shape", [(30, 700, 4, 4), (12, 260, 20, 2), (9, 120, 61, 1), (10, 90, 5, 2)], ids=lambda s: "S%d" % s[2])
def test_clone_reproduces_and_is_independent(shape):
    T, P, S, C = shape
    pb = _synthetic_problem(T, P, S, C, seed=8100 + S)
    want = O.evaluate(pb)
    src = phb.SingleTreeLikelihood.from_problem(pb)
    lnl, g = src.calculate(), src.gradient().copy()
    twin = src.clone()
    assert twin.calculate() == lnl and np.array_equal(twin.gradient(), g), "same kernels, same inputs: bit-identical"
    # the clone moves on; the source must not notice
    pb2 = copy.copy(pb)
    pb2.bl = pb.bl * 1.3
    twin.set_branch_lengths(pb2.bl)
    want2 = O.evaluate(pb2)
    assert rel_err(twin.calculate(), want2["lnl"]) < RTOL and grad_err(twin.gradient(), want2["grad"]) < RTOL
    src.update_all_nodes()
    assert src.calculate() == lnl and grad_err(src.gradient(), want["grad"]) < RTOL
    # a clone of a clone, after the source is gone
    src.close()
    third = twin.clone()
    twin.close()
    assert rel_err(third.calculate(), want2["lnl"]) < RTOL
    third.close()


def test_clone_keeps_options_rescaling_tip_partials_and_explicit_matrices():
    pb, _ = load_golden("synth_gtr_g4_deep_scaled")
    pb.scale = True  # rescaling on in the source: the clone must inherit it
    want = O.evaluate(pb)
    src = phb.SingleTreeLikelihood.from_problem(pb)
    lnl = src.calculate()
    assert src.rescaling()
    twin = src.clone()
    assert twin.rescaling()
    assert twin.calculate() == lnl and grad_err(twin.gradient(), want["grad"]) < 1e-9
    src.close(), twin.close()
    for name in ("c1_jc69_fluA_tippartials", "c1_jc69_fluA_tipstates"):  # closed-form JC69 matrices uploaded explicitly
        pb, _ = load_golden(name)
        want = O.evaluate(pb)
        src = phb.SingleTreeLikelihood.from_problem(pb)
        twin = src.clone()
        src.close()
        assert rel_err(twin.calculate(), want["lnl"]) < RTOL and grad_err(twin.gradient(), want["grad"]) < RTOL
        twin.close()


def test_clone_carries_the_time_tree():
    from tests.test_time_tree import _c1

    pb, z, _ = _c1()
    src = phb.SingleTreeLikelihood.from_problem(pb)
    src.set_time_tree(z["tip_heights"])
    a = src.gradient_batch_time(z["ratios"][:3], z["rates"][:3, None] if z["rates"].ndim == 1 else z["rates"][:3], include_jacobian=True)
    twin = src.clone()
    src.close()
    b = twin.gradient_batch_time(z["ratios"][:3], z["rates"][:3, None] if z["rates"].ndim == 1 else z["rates"][:3], include_jacobian=True)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    twin.close()


@pytest.mark.parametrize("kernels", [phb.KERNELS_AUTO, phb.KERNELS_GENERIC], ids=["auto", "generic"])
@pytest.mark.parametrize("shape", [(40, 900, 4, 4), (14, 300, 20, 2), (9, 100, 61, 1)], ids=lambda s: "S%d" % s[2])
def test_topology_moves_against_oracle(shape, kernels):
    T, P, S, C = shape
    pb = _synthetic_problem(T, P, S, C, seed=8200 + S)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    assert rel_err(tlk.calculate(), O.evaluate(pb, gradient=False)["lnl"]) < RTOL
    rng = np.random.default_rng(8201)
    cur = pb
    for _ in range(4):  # a short NNI walk
        cands = [n for n in range(T, 2 * T - 1) if n != cur.root]
        cur = _nni(cur, int(rng.choice(cands)))
        tlk.set_topology(cur.left, cur.right, cur.root)
        want = O.evaluate(cur)
        assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
        assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    # a broken tree (a cycle: an internal node made its own grandchild) is refused and the current tree keeps working
    bad_l, bad_r = cur.left.copy(), cur.right.copy()
    n = next(n for n in range(T, 2 * T - 1) if n != cur.root and cur.left[n] >= T)
    bad_l[int(cur.left[n])] = n
    with pytest.raises(phb.PhysherB200Error):
        tlk.set_topology(bad_l, bad_r, cur.root)
    tlk.set_branch_lengths(cur.bl)
    assert rel_err(tlk.calculate(), O.evaluate(cur, gradient=False)["lnl"]) < RTOL
    tlk.close()


def test_topology_move_with_resident_partials_and_three_node_updates():
    """serial-Brent / NNI access pattern: uppers resident, a move, update_three_nodes, single-branch evaluations on the new tree"""
    T, P, S, C = 24, 500, 4, 4
    pb = _synthetic_problem(T, P, S, C, seed=8300)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.update_uppers()
    node = next(n for n in range(T, 2 * T - 1) if n != pb.root)
    cur = _nni(pb, node)
    tlk.set_topology(cur.left, cur.right, cur.root)
    tlk.update_three_nodes(node)
    tlk.update_uppers()
    want = O.evaluate(cur)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    # lnL along one branch of the new tree agrees with full re-evaluations
    target = int(cur.left[node])
    cand = np.array([0.5, 1.0, 2.0]) * cur.bl[target]
    lnl, d1, _ = tlk.calculate_branch(target, cand)
    for k, b in enumerate(cand):
        q = copy.copy(cur)
        q.bl = cur.bl.copy()
        q.bl[target] = b
        ref = O.evaluate(q)
        assert rel_err(lnl[k], ref["lnl"]) < RTOL
        assert abs(d1[k] - ref["grad"][target]) <= 1e-9 * max(abs(ref["grad"][target]), np.abs(ref["grad"]).max() * 1e-6)
    # update_three_nodes on an unchanged tree only marks nodes dirty: same lnL
    tlk.update_three_nodes(node)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    tlk.close()


def test_clone_to_another_device():
    """with two visible GPUs: the clone lives on the other device (data travels device to device) and agrees bit for bit"""
    if phb.device_count() < 2:
        pytest.skip("one GPU visible")
    pb = _synthetic_problem(30, 900, 4, 4, seed=8400)
    src = phb.SingleTreeLikelihood.from_problem(pb, device=0)
    lnl, g = src.calculate(), src.gradient().copy()
    twin = src.clone(device=1)
    src.close()
    assert twin.calculate() == lnl and np.array_equal(twin.gradient(), g)
    twin.close()
