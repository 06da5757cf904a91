"""Time-tree chain: ratios / root height -> heights -> branch lengths -> lnL + gradients -> ratio, root-height
and clock-rate gradients, batched over samples.

CPU: the oracle's restatement (the reference's naive recursive forms) against the reference's own outputs -- the known answers of
tests/test_tree_likelihood.c and perturbed samples run through the unmodified reference (tests/golden/c1_time_tree.npz).
GPU: phb_tlk_gradient_batch_time (an adjoint sweep on the device) against both.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import GOLDEN, RTOL, grad_err, load_golden, rel_err


def _c1():
    pb, _ = load_golden("c1_jc69_fluA_tipstates")
    # the fixture carries the reference's closed-form JC69 matrices for ITS branch lengths (jc69.c:73-94); samples move the
    # branch lengths, so give the same model as an eigen system: Q = (J - 4 I) / 3 is diagonalised by the 4 x 4 Hadamard matrix
    # (entries +-1/2: exact in binary, so P(t) agrees with the closed form to an ulp), eigenvalues 0, -4/3, -4/3, -4/3
    vec = 0.5 * np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, 1, -1], [1, -1, -1, 1]], dtype=np.float64)
    pb.evec, pb.eval, pb.ivec = vec, np.array([0.0, -4.0 / 3.0, -4.0 / 3.0, -4.0 / 3.0]), vec.copy()
    pb.P_override = pb.dP_override = None
    z = dict(np.load(os.path.join(GOLDEN, "c1_time_tree.npz")))
    kat = json.load(open(os.path.join(GOLDEN, "c1_kat.json")))
    return pb, z, kat


def test_oracle_chain_reproduces_reference_known_answers():
    pb, z, kat = _c1()
    for jac in (False, True):
        out = O.time_evaluate(pb, z["tip_heights"], z["ratios"][0], z["rates"][:1], include_jacobian=jac)
        want = np.array((kat["ratio_jac_grad"] if jac else kat["ratio_grad"]) + [kat["root_height_jac_grad" if jac else "root_height_grad"]])
        assert abs(out["lnl"] - kat["logP"]) < 1e-8
        assert np.abs(out["grad_ratios"] - want).max() < 1e-8 and grad_err(out["grad_ratios"], want) < RTOL
        assert rel_err(out["grad_rates"][0], kat["rate_grad"]) < RTOL
        assert abs(out["lnl"] + out["log_jacobian"] - kat["logP_jacobian"]) < 1e-8
        np.testing.assert_allclose(out["bl"], pb.bl, rtol=1e-14, atol=0)


def test_oracle_chain_matches_reference_on_perturbed_samples():
    pb, z, _ = _c1()
    for b in range(z["ratios"].shape[0]):
        for jac, key in ((False, "ref_grad"), (True, "ref_grad_jacobian")):
            out = O.time_evaluate(pb, z["tip_heights"], z["ratios"][b], z["rates"][b:b + 1], include_jacobian=jac)
            assert rel_err(out["lnl"], z["ref_lnl"][b]) < RTOL
            assert rel_err(out["lnl"] + out["log_jacobian"], z["ref_lnl_jacobian"][b]) < RTOL
            got = np.concatenate([out["grad_ratios"], out["grad_rates"]])
            assert grad_err(got, z[key][b]) < RTOL


def _synthetic_time_problem(T=40, P=700, C=4, seed=90):
    from physher_b200 import models, synthetic as syn

    topo = syn.random_topology(T, seed)
    m = models.gtr([0.05, 0.3, 0.1, 0.15, 0.3, 0.1], [0.1, 0.2, 0.3, 0.4])
    rates, props = models.discrete_gamma(0.5, C)
    pb = O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=4,
                   tip_states=syn.random_patterns(T, P, 4, 0.2, seed + 1, unknown_frac=0.01),
                   weights=np.random.default_rng(seed + 2).integers(1, 4, P).astype(np.float64),
                   freqs=m.freqs, rates=rates, props=props, bl=np.zeros(2 * T - 1), evec=m.evec, eval=m.eval, ivec=m.ivec, unrooted=False)
    rng = np.random.default_rng(seed + 3)
    tip_heights = rng.uniform(0.0, 3.0, T)  # heterochronous tips
    return pb, tip_heights, rng


@pytest.mark.gpu
@pytest.mark.parametrize("jacobian", [False, True])
def test_gpu_batched_chain_c1(jacobian):
    import physher_b200 as phb

    pb, z, kat = _c1()
    pb.unrooted = False
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.set_time_tree(z["tip_heights"])
    n0 = tlk.launch_count()
    lnl, lj, gr, gc = tlk.gradient_batch_time(z["ratios"], z["rates"][:, None], include_jacobian=jacobian)
    assert tlk.launch_count() - n0 <= 6, "forward chain, matrices, one fused walk, finalize, backward chain (+ tip encoding)"
    key = "ref_grad_jacobian" if jacobian else "ref_grad"
    for b in range(z["ratios"].shape[0]):
        assert rel_err(lnl[b], z["ref_lnl"][b]) < RTOL
        assert rel_err(lnl[b] + lj[b], z["ref_lnl_jacobian"][b]) < RTOL
        assert grad_err(np.concatenate([gr[b], gc[b]]), z[key][b]) < RTOL
    want = np.array((kat["ratio_jac_grad"] if jacobian else kat["ratio_grad"]) + [kat["root_height_jac_grad" if jacobian else "root_height_grad"]])
    assert np.abs(gr[0] - want).max() < 1e-8 and abs(gc[0, 0] - kat["rate_grad"]) < 1e-8 * abs(kat["rate_grad"])
    # lnL only
    lnl2, lj2, none_r, none_c = tlk.gradient_batch_time(z["ratios"], z["rates"][:, None], want_gradient=False)
    assert none_r is None and np.allclose(lnl2, lnl, rtol=1e-13, atol=0) and np.array_equal(lj2, lj)
    tlk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("per_node_rates", [False, True])
@pytest.mark.parametrize("kernels", ["auto", "generic"])
def test_gpu_batched_chain_synthetic(per_node_rates, kernels):
    """GTR+G4, non-uniform pi, heterochronous tips, strict and per-node clock rates, both kernel families."""
    import physher_b200 as phb

    pb, tip_heights, rng = _synthetic_time_problem()
    T, N = pb.ntips, pb.nnodes
    B = 9
    ratios = rng.uniform(0.2, 0.9, size=(B, T - 1))
    ratios[:, -1] = tip_heights.max() + rng.uniform(0.5, 2.0, size=B)  # root height above every tip
    rates = rng.lognormal(np.log(0.02), 0.3, size=(B, N if per_node_rates else 1))
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_GENERIC if kernels == "generic" else phb.KERNELS_AUTO)
    tlk.set_time_tree(tip_heights)
    lnl, lj, gr, gc = tlk.gradient_batch_time(ratios, rates, include_jacobian=True)
    for b in range(B):
        want = O.time_evaluate(pb, tip_heights, ratios[b], rates[b], include_jacobian=True)
        assert rel_err(lnl[b], want["lnl"]) < RTOL and rel_err(lj[b], want["log_jacobian"]) < 1e-12
        assert grad_err(gr[b], want["grad_ratios"]) < RTOL
        assert grad_err(gc[b], want["grad_rates"]) < RTOL
    tlk.close()


@pytest.mark.gpu
def test_gpu_chain_rejects_negative_branch_lengths_and_missing_setup():
    import physher_b200 as phb

    pb, tip_heights, rng = _synthetic_time_problem(T=10, P=64, C=1)
    T = pb.ntips
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    ratios = rng.uniform(0.2, 0.9, size=(2, T - 1))
    ratios[:, -1] = tip_heights.max() + 1.0
    with pytest.raises(phb.PhysherB200Error, match="set_time_tree"):
        tlk.gradient_batch_time(ratios, np.full((2, 1), 0.01))
    tlk.set_time_tree(tip_heights)
    ratios[1, -1] = tip_heights.max() - 1.0  # root younger than a tip => negative length (treelikelihood.c:1659-1662 exits)
    with pytest.raises(phb.PhysherB200Error, match="negative branch length"):
        tlk.gradient_batch_time(ratios, np.full((2, 1), 0.01))
    tlk.close()
