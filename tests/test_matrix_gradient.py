"""Substitution-model parameter gradients: the node sweep of calculate_dlnl_dQ (treelikelihood.c:2337-2583)
with per-node matrices dP/d theta supplied by the caller (m->dPdp).

Golden: tests/golden/dlnl_dq_gtr_g4.npz -- the reference's own dPdp matrices for the five free GTR rate parameters and its own
calculate_dlnl_dQ values, with include_root_freqs false and true.  CPU: the oracle's restatement (from its own lower / upper
partials) against them.  GPU: phb_tlk_matrix_gradient against both, plus the tensor-core state counts against the oracle.
"""
import numpy as np
import pytest

import physher_b200 as phb
from oracle import oracle as O
from tests.util import RTOL, grad_err, load_golden, rel_err


def _case():
    pb, z = load_golden("dlnl_dq_gtr_g4")
    return pb, z


@pytest.mark.parametrize("irf,key", [(False, "ref_dlnl_dq"), (True, "ref_dlnl_dq_root_freqs")])
def test_oracle_reproduces_reference_dlnl_dq(irf, key):
    pb, z = _case()
    pb.include_root_freqs = irf
    got = O.matrix_gradient(pb, z["dPdp"])
    assert grad_err(got, z[key]) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("kernels", ["generic", "auto"])
@pytest.mark.parametrize("irf,key", [(False, "ref_dlnl_dq"), (True, "ref_dlnl_dq_root_freqs")])
def test_gpu_reproduces_reference_dlnl_dq(irf, key, kernels):
    import physher_b200 as phb
    from physher_b200.treelikelihood import OPT_INCLUDE_ROOT_FREQS

    pb, z = _case()
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_GENERIC if kernels == "generic" else phb.KERNELS_AUTO)
    tlk.set_option(OPT_INCLUDE_ROOT_FREQS, int(irf))
    got = tlk.matrix_gradient(z["dPdp"])
    assert grad_err(got, z[key]) < RTOL
    assert rel_err(tlk.calculate(), float(z["ref_lnl"])) < RTOL  # the sweep leaves lnL cached
    pb.include_root_freqs = irf
    if not irf:
        # rescaling on: the dlikelihood / likelihood form (:2464-2474) is the same derivative.  (With include_root_freqs and a
        # non-uniform pi the reference's per-node denominators are not the site likelihood -- SURVEY.md 0.4 iii -- so that
        # combination has no exact value to agree with.)
        tlk.use_rescaling(True)
        assert grad_err(tlk.matrix_gradient(z["dPdp"]), z[key]) < 1e-9
        pb.scale = True
    # the branch gradient still works afterwards
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    tlk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(9, 130, 20, 2), (7, 70, 61, 1), (10, 90, 5, 3)])
def test_gpu_matrix_gradient_other_state_counts(shape):
    """20 / 61 states: the forward phase once and one tensor-core gradient phase per matrix set (phbc_dmma_matrix_gradient; the 20-state
    request runs on the whole-tree walk); 5 states generic."""
    import physher_b200 as phb
    from tests.test_gpu_parity import _synthetic_problem

    T, P, S, C = shape
    pb = _synthetic_problem(T, P, S, C, seed=5000 + S, unknown=0.03)
    rng = np.random.default_rng(5001)
    M = rng.normal(size=(3, pb.nnodes, C, S, S))
    want = O.matrix_gradient(pb, M)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    got = tlk.matrix_gradient(M)
    assert grad_err(got, want) < RTOL
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    tlk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["rescaled", "ambiguous-tip-partials", "root-freqs-unrooted"])
@pytest.mark.parametrize("S,C", [(20, 2), (61, 1)])
def test_gpu_matrix_gradient_tensor_paths(S, C, variant):
    """the same request on the level-batched kernels the walk / message form decline to: under rescaling (exact gradient from the
    cumulative scaling factors), with tip partials that carry an ambiguity set, and with the root frequencies folded into the root's
    children on an unrooted tree (the root's right child is skipped, treelikelihood.c:2408)"""
    import physher_b200 as phb
    from physher_b200.treelikelihood import OPT_INCLUDE_ROOT_FREQS, OPT_UNROOTED
    from tests.test_gpu_parity import _synthetic_problem

    pb = _synthetic_problem(11, 150, S, C, seed=5100 + S, unknown=0.02)
    if variant == "ambiguous-tip-partials":
        pb.use_tip_states = False
        pb.tip_partials = np.eye(S)[np.minimum(pb.tip_states, S - 1)]
        pb.tip_partials[pb.tip_states >= S] = 1.0
        pb.tip_partials[2, 7, : S // 3] = 1.0
    if variant == "root-freqs-unrooted":
        pb.include_root_freqs = True
        pb.unrooted = True
    rng = np.random.default_rng(5101)
    M = rng.normal(size=(2, pb.nnodes, C, S, S))
    want = O.matrix_gradient(pb, M)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    if variant == "rescaled":
        tlk.use_rescaling(True)
    got = tlk.matrix_gradient(M)
    assert grad_err(got, want) < (1e-9 if variant == "rescaled" else RTOL)
    tlk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("S,C,scale", [(4, 4, False), (4, 4, True), (20, 2, False), (61, 1, False)], ids=["nuc4", "nuc4-scaled", "aa20", "codon61"])
def test_root_frequency_gradient(S, C, scale):
    """d lnL / d pi_i at fixed partials (root term of calculate_dlnl_dQ, treelikelihood.c:2371-2404) against the oracle's root partials"""
    from physher_b200 import models, synthetic as syn

    T, P = 8, 211
    topo = syn.random_topology(T, 5)
    m = models.gtr([0.05, 0.3, 0.1, 0.15, 0.3, 0.1], [0.1, 0.2, 0.3, 0.4]) if S == 4 else models.random_reversible(S, 9)
    rates, props = models.discrete_gamma(0.5, C) if C > 1 else (np.ones(1), np.ones(1))
    pb = O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=S,
                   tip_states=syn.random_patterns(T, P, S, 0.25, 6, unknown_frac=0.02), weights=np.random.default_rng(7).integers(1, 4, P).astype(np.float64),
                   freqs=m.freqs, rates=rates, props=props, bl=syn.random_branch_lengths(topo, 8), evec=m.evec, eval=m.eval, ivec=m.ivec)
    res = O.evaluate(pb, gradient=False, partials=True)
    R = np.einsum("c,cpi->pi", props, res["lower"][pb.root])
    want = (pb.weights[:, None] * R / (R @ pb.freqs)[:, None]).sum(0)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    if scale:
        tlk.set_option(phb.treelikelihood.OPT_SCALING_THRESHOLD_EXP, 2)
        tlk.use_rescaling(True)
    got = tlk.root_frequency_gradient()
    assert grad_err(got, want) < (1e-9 if scale else RTOL)
    # sum_i pi_i G_i = sum_k w_k exactly
    assert abs(got @ pb.freqs - pb.weights.sum()) < 1e-9 * pb.weights.sum()
    tlk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tips", ["states", "partials"])
@pytest.mark.parametrize("shape", [(37, 1500, 4), (21, 333, 1), (18, 257, 2), (26, 65, 8), (128, 900, 4)], ids=lambda s: "T%d-P%d-C%d" % s)
def test_gpu_fused_walk_transition_statistics(shape, tips):
    """4 states, unscaled: ONE launch of the fused walk accumulates the per-branch statistics G[n][c][i][j] every matrix set
    contracts with (no materialised upper partials); arbitrary matrices, unknown states / ambiguity sets at the tips, and the
    root term of the frequency gradient from the statistics' root entry.  Checked against the oracle's node sweep and against
    the node-at-a-time kernels."""
    import physher_b200 as phb
    from physher_b200 import synthetic as syn
    from tests.test_gpu_parity import _synthetic_problem

    T, P, C = shape
    topo = syn.caterpillar_topology(T) if T == 128 else None
    pb = _synthetic_problem(T, P, 4, C, seed=7000 + T + C, topo=topo, unknown=0.04)
    if tips == "partials":
        rng = np.random.default_rng(7100 + T)
        tp = np.zeros((T, P, 4))
        known = pb.tip_states < 4
        tp[known, pb.tip_states[known]] = 1.0
        tp[~known] = 1.0
        amb = rng.random((T, P)) < 0.05  # ambiguity sets (R, Y, ...): two or three states set
        tp[amb] = (rng.random((int(amb.sum()), 4)) < 0.5).astype(np.float64)
        tp[tp.sum(-1) == 0] = 1.0
        pb.tip_partials = tp
        pb.use_tip_states = False
    rng = np.random.default_rng(7200 + C)
    M = rng.normal(size=(6, pb.nnodes, C, 4, 4))
    want = O.matrix_gradient(pb, M)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    before = tlk.launch_count()
    got = tlk.matrix_gradient(M)
    assert tlk.launch_count() - before <= 7, "the fused walk serves the request in a handful of launches"
    assert grad_err(got, want) < RTOL
    res = O.evaluate(pb, gradient=False, partials=True)
    R = np.einsum("c,cpi->pi", pb.props, res["lower"][pb.root])
    want_root = (pb.weights[:, None] * R / (R @ pb.freqs)[:, None]).sum(0)
    before = tlk.launch_count()
    assert grad_err(tlk.root_frequency_gradient(), want_root) < RTOL
    assert tlk.launch_count() == before, "served from the statistics' root entry"
    assert rel_err(tlk.calculate(), res["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    # same request on the node-at-a-time kernels
    gen = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_GENERIC)
    assert grad_err(gen.matrix_gradient(M), want) < RTOL
    gen.close()
    # include_root_freqs (the reference's default request) on the fused path
    from physher_b200.treelikelihood import OPT_INCLUDE_ROOT_FREQS

    tlk.set_option(OPT_INCLUDE_ROOT_FREQS, 1)
    pb.include_root_freqs = True
    assert grad_err(tlk.matrix_gradient(M), O.matrix_gradient(pb, M)) < RTOL
    tlk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(37, 1500, 4), (21, 333, 1), (26, 65, 8), (128, 900, 4)], ids=lambda s: "T%d-P%d-C%d" % s)
def test_gpu_fused_walk_transition_statistics_rescaled(shape):
    """The same request with rescaling on and a threshold that makes nearly every node rescale (1e-2): each branch's statistics are
    normalised by the site likelihood in that branch's own scale (dlikelihood / likelihood, treelikelihood.c:2464-2474), still ONE fused
    launch; the values are the unscaled derivative.  Requests the walk cannot serve exactly (include_root_freqs, the reference-compatible
    per-category normalisation) go to the node-at-a-time kernels and must agree as well."""
    import physher_b200 as phb
    from physher_b200 import synthetic as syn
    from physher_b200.treelikelihood import OPT_COMPAT_SCALED_GRADIENT, OPT_SCALING_THRESHOLD_EXP
    from tests.test_gpu_parity import _synthetic_problem

    T, P, C = shape
    topo = syn.caterpillar_topology(T) if T == 128 else None
    pb = _synthetic_problem(T, P, 4, C, seed=7300 + T + C, topo=topo, unknown=0.04)
    rng = np.random.default_rng(7400 + C)
    M = rng.normal(size=(5, pb.nnodes, C, 4, 4))
    want = O.matrix_gradient(pb, M)
    res = O.evaluate(pb, gradient=True, partials=True)
    R = np.einsum("c,cpi->pi", pb.props, res["lower"][pb.root])
    want_root = (pb.weights[:, None] * R / (R @ pb.freqs)[:, None]).sum(0)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.set_option(OPT_SCALING_THRESHOLD_EXP, 2)
    tlk.use_rescaling(True)
    before = tlk.launch_count()
    got = tlk.matrix_gradient(M)
    assert tlk.launch_count() - before <= 7, "served by the fused walk"
    assert grad_err(got, want) < 1e-9
    before = tlk.launch_count()
    assert grad_err(tlk.root_frequency_gradient(), want_root) < 1e-9
    assert tlk.launch_count() == before
    assert rel_err(tlk.calculate(), res["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), res["grad"]) < 1e-9
    # node-at-a-time kernels on the same scaled request
    gen = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_GENERIC)
    gen.set_option(OPT_SCALING_THRESHOLD_EXP, 2)
    gen.use_rescaling(True)
    assert grad_err(gen.matrix_gradient(M), want) < 1e-9
    gen.close()
    # the reference-compatible normalisation is not something the walk's statistics can express: declined, served by the sweep
    tlk.set_option(OPT_COMPAT_SCALED_GRADIENT, 1)
    before = tlk.launch_count()
    assert grad_err(tlk.matrix_gradient(M), want) < 1e-9  # the sweep always uses dlikelihood / likelihood
    assert tlk.launch_count() - before > 7
    tlk.close()
