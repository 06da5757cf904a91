"""CUDA path vs the oracle and the reference's golden outputs, through the C ABI (-m gpu).

Tolerances (north_star): lnL and every branch gradient within 1e-10 relative in FP64; gradient
entries relative to max(|g|, |g|_inf * 1e-6).
"""
import numpy as np
import pytest

import physher_b200 as phb
from oracle import oracle as O
from physher_b200 import models, synthetic as syn
from physher_b200.treelikelihood import OPT_COMPAT_SCALED_GRADIENT, OPT_INCLUDE_ROOT_FREQS
from tests.util import RTOL, golden_names, grad_err, load_golden, rel_err

pytestmark = pytest.mark.gpu

KERNELS = [phb.KERNELS_GENERIC, phb.KERNELS_AUTO]
KIDS = ["generic", "auto"]


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
@pytest.mark.parametrize("name", golden_names())
def test_golden_lnl_and_gradient(name, kernels):
    pb, z = load_golden(name)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    lnl = tlk.calculate()
    assert rel_err(lnl, float(z["ref_lnl"])) < RTOL
    np.testing.assert_allclose(tlk.pattern_log_likelihoods(), z["ref_pattern_lnl"], rtol=1e-10, atol=0)
    want = O.evaluate(pb)
    g = tlk.gradient()
    assert grad_err(g, want["grad"]) < RTOL
    if "ref_grad_exact" in z:
        assert grad_err(g, z["ref_grad_exact"]) < RTOL
        # the reference's default request (include_root_freqs = true, SURVEY 0.4 iii)
        tlk.set_option(OPT_INCLUDE_ROOT_FREQS, 1)
        assert grad_err(tlk.gradient(), z["ref_grad_default"]) < RTOL
        tlk.set_option(OPT_INCLUDE_ROOT_FREQS, 0)
        tlk.gradient()
    cg = tlk.cat_branch_gradient()
    assert grad_err(cg.ravel(), np.where(np.arange(pb.nnodes)[:, None] == pb.root, 0, want["cat_grad"]).ravel()) < RTOL
    tlk.close()


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
def test_kat_jc69_fluA(kernels):
    """tests/test_tree_likelihood.c:28-40: lnL and clock-rate gradient of the reference's own known-answer test."""
    pb, z = load_golden("c1_jc69_fluA_tipstates")
    pb.include_root_freqs = True
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    lnl = tlk.calculate()
    assert abs(lnl - float(z["kat_lnl"])) < 1e-8
    assert rel_err(lnl, float(z["kat_lnl"])) < RTOL
    clock = float((tlk.gradient() * z["time_elapsed"]).sum())
    assert rel_err(clock, float(z["kat_clock_grad"])) < RTOL
    tlk.close()


def _jc69_eigen(pi=None):
    """Q = beta (1 pi^T - I), beta = 1 / (1 - sum pi^2): eigenvalues (0, -beta x 3), V = [1 | e_k - pi_k 1], V^-1 = [pi ; e_k - e_4] -- what
    integration/physher_glue.c hands over for JC69 and F81 (jc69.c:73-94, f81.c:45-110)."""
    pi = np.full(4, 0.25) if pi is None else np.asarray(pi, dtype=np.float64)
    beta = 1.0 / (1.0 - float((pi * pi).sum()))
    evec, ivec = np.zeros((4, 4)), np.zeros((4, 4))
    evec[:, 0] = 1.0
    ivec[0] = pi
    for k in range(1, 4):
        evec[:, k] = -pi[k - 1]
        evec[k - 1, k] += 1.0
        ivec[k, k - 1], ivec[k, 3] = 1.0, -1.0
    return evec, np.array([0.0, -beta, -beta, -beta]), ivec


def test_kat_jc69_fluA_through_the_fused_walk():
    """The reference's known answers (tests/test_tree_likelihood.c:28-40) with JC69 given as its closed-form EIGEN system, so that
    C1 runs on k_nuc4_walk (the fixture's explicit matrices can only reach the node-at-a-time kernels); the kernel family that ran
    is asserted, and the matrices the walk consumed are the reference's jc69_p_t / jc69_dp_dt values."""
    pb, z = load_golden("c1_jc69_fluA_tipstates")
    ref_P, ref_dP = pb.P_override, pb.dP_override
    pb.evec, pb.eval, pb.ivec = _jc69_eigen()
    pb.P_override = pb.dP_override = None
    pb.include_root_freqs = True
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_AUTO)
    lnl = tlk.calculate()
    assert tlk.last_kernels() == phb.treelikelihood.RAN_WALK
    assert abs(lnl - float(z["kat_lnl"])) < 1e-8 and rel_err(lnl, float(z["kat_lnl"])) < RTOL
    g = tlk.gradient()
    assert tlk.last_kernels() == phb.treelikelihood.RAN_WALK
    clock = float((g * z["time_elapsed"]).sum())
    assert rel_err(clock, float(z["kat_clock_grad"])) < RTOL
    Pm, dPm = tlk.get_matrices()
    keep = np.arange(pb.nnodes) != pb.root
    np.testing.assert_allclose(Pm[keep], ref_P[keep], rtol=0, atol=4e-16)
    np.testing.assert_allclose(dPm[keep], ref_dP[keep], rtol=0, atol=4e-15)
    np.testing.assert_allclose(tlk.pattern_log_likelihoods(), z["ref_pattern_lnl"], rtol=1e-10, atol=0)
    tlk.close()


def test_f81_closed_form_eigen_system():
    """F81 (modeltype JC69 in the reference, f81.c:33): the closed-form eigen system of the glue against f81_p_t's formula."""
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    evec, ev, ivec = _jc69_eigen(pi)
    beta = -ev[1]
    for t in (1e-4, 0.03, 0.7):
        Pt = np.abs(evec @ np.diag(np.exp(ev * t)) @ ivec)
        temp = np.exp(-t * beta)
        want = np.tile(pi * (1.0 - temp), (4, 1)) + temp * np.eye(4)
        np.testing.assert_allclose(Pt, want, rtol=0, atol=3e-16)  # both forms cancel at small t: 2e-12 relative on the 1e-5 entries


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
@pytest.mark.parametrize("name", [n for n in golden_names() if "deep_scaled" in n or n.startswith("synth_gtr_g4_tip")])
def test_rescaling(name, kernels):
    pb, z = load_golden(name)
    base = O.evaluate(pb)
    pb.scale = True
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    assert tlk.rescaling()
    lnl = tlk.calculate()
    assert rel_err(lnl, float(z["ref_lnl_scaled"])) < RTOL
    # exact form under rescaling reproduces the unscaled gradient
    g = tlk.gradient()
    assert grad_err(g, base["grad"]) < 1e-9
    assert grad_err(g, O.evaluate(pb)["grad"]) < RTOL
    # reference-compatible per-category normalisation (SURVEY 0.4 ii)
    tlk.set_option(OPT_COMPAT_SCALED_GRADIENT, 1)
    assert grad_err(tlk.gradient(), z["ref_grad_scaled_compat"]) < RTOL
    tlk.close()


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
def test_transition_matrices_and_partials(kernels):
    """A4 + K1-K4 + K8 entry by entry against the reference's own buffers.  phb_tlk_get_matrices hands back what the kernels of the
    LAST evaluation consumed: the node-at-a-time arrays (k_transition_matrices), or, after a fused evaluation, the walk-ordered
    set k_nuc4_matrices wrote (dP as the walk contracts it, Q P)."""
    pb, z = load_golden("tiny_gtr_g4")
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    tlk.gradient()
    assert tlk.last_kernels() == (phb.treelikelihood.RAN_GENERIC if kernels == phb.KERNELS_GENERIC else phb.treelikelihood.RAN_WALK)
    Pm, dPm = tlk.get_matrices()
    keep = np.arange(pb.nnodes) != pb.root
    np.testing.assert_allclose(Pm[keep], z["ref_matrices"][keep], rtol=0, atol=1e-13)
    np.testing.assert_allclose(dPm[keep], z["ref_dmatrices"][keep], rtol=0, atol=1e-12)
    # tlk->partials as the reference holds them (the fused walk keeps none: get_partials runs the node-at-a-time kernels)
    for n in range(pb.ntips, pb.nnodes):
        np.testing.assert_allclose(tlk.get_partials(n), z["ref_lower"][n], rtol=1e-10, atol=0)
    for n in range(pb.nnodes):
        if n != pb.root:
            np.testing.assert_allclose(tlk.get_partials(pb.nnodes + n), z["ref_upper"][n], rtol=1e-10, atol=1e-300)
    tlk.close()


@pytest.mark.parametrize("name", ["synth_lg_g4_tipstates", "synth_lg_g4_tippartials", "synth_gy94_tippartials", "synth_hky_g4_tipstates",
                                  "synth_gtr_g4_tippartials"])
def test_fast_path_matrix_images_match_reference(name):
    """k_dmma_pack's packed images (transposed tip images, zero-padded operand images, the pi-weighted transposed adjoint images of
    the message form) and k_nuc4_matrices' walk-ordered set, unpacked, against the reference's p_t / dp_dt entry by entry."""
    pb, z = load_golden(name)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_AUTO)
    tlk.gradient()
    S = pb.nstate
    assert tlk.last_kernels() == (phb.treelikelihood.RAN_WALK if S == 4 else phb.treelikelihood.RAN_TENSOR)
    Pm, dPm = tlk.get_matrices()
    keep = np.arange(pb.nnodes) != pb.root
    # 61 states: the exponentials come from the host's libm, so the images carry the reference's values exactly;
    # the adjoint images are divided by pi again on the way back (one rounding)
    np.testing.assert_allclose(Pm[keep], z["ref_matrices"][keep], rtol=0, atol=1e-13)
    np.testing.assert_allclose(dPm[keep], z["ref_dmatrices"][keep], rtol=1e-13, atol=1e-12)
    tlk.close()


def _synthetic_problem(T, P, S, C, seed, topo=None, unknown=0.01):
    topo = topo or syn.random_topology(T, seed)
    if S == 4:
        m = models.gtr([0.05, 0.3, 0.1, 0.15, 0.3, 0.1], [0.1, 0.2, 0.3, 0.4])
    else:
        m = models.random_reversible(S, seed + 5)
    rates, props = models.discrete_gamma(0.5, C)
    return O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=S,
                     tip_states=syn.random_patterns(T, P, S, 0.15, seed + 1, unknown_frac=unknown),
                     weights=np.random.default_rng(seed + 2).integers(1, 5, P).astype(np.float64),
                     freqs=m.freqs, rates=rates, props=props, bl=syn.random_branch_lengths(topo, seed + 3),
                     evec=m.evec, eval=m.eval, ivec=m.ivec)


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
@pytest.mark.parametrize("shape", [(50, 1000, 4, 4), (33, 517, 4, 1), (64, 300, 4, 2), (20, 130, 4, 3), (17, 1, 4, 4),
                                   (12, 333, 20, 4), (9, 200, 20, 1), (7, 150, 61, 1), (6, 90, 61, 2), (10, 100, 5, 2), (8, 64, 7, 1)])
def test_synthetic_against_oracle(shape, kernels):
    """Seeded synthetic inputs at sizes the oracle finishes in seconds; ragged pattern counts included."""
    T, P, S, C = shape
    pb = _synthetic_problem(T, P, S, C, seed=1000 + T + P)
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    tlk.close()


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
@pytest.mark.parametrize("topo_kind", ["caterpillar", "balanced"])
def test_extreme_topologies(topo_kind, kernels):
    T = 128
    topo = syn.caterpillar_topology(T) if topo_kind == "caterpillar" else syn.balanced_topology(T)
    pb = _synthetic_problem(T, 200, 4, 4, seed=77, topo=topo)
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    tlk.close()


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
def test_tip_partials_mode_matches_tip_states_for_known_states(kernels):
    pb = _synthetic_problem(15, 256, 4, 4, seed=5, unknown=0.0)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    a = tlk.calculate()
    ga = tlk.gradient()
    tlk.close()
    pb.use_tip_states = False
    pb.tip_partials = np.eye(4)[pb.tip_states]
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    assert rel_err(tlk.calculate(), a) < 1e-13
    assert grad_err(tlk.gradient(), ga) < 1e-12
    tlk.close()


def test_caching_semantics():
    """cached lnL if nothing is dirty (treelikelihood.c:1458); gradient cached until something changes (:323,:336)."""
    pb = _synthetic_problem(20, 128, 4, 4, seed=9)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    a = tlk.calculate()
    n0 = tlk.launch_count()
    assert tlk.calculate() == a and tlk.launch_count() == n0
    g = tlk.gradient()
    n1 = tlk.launch_count()
    assert np.array_equal(tlk.gradient(), g) and tlk.launch_count() == n1
    tlk.set_branch_length(3, pb.bl[3] * 1.5)
    b = tlk.calculate()
    assert b != a and tlk.launch_count() > n1
    pb.bl[3] *= 1.5
    assert rel_err(b, O.evaluate(pb, gradient=False)["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    tlk.close()


def test_store_restore_mcmc_reject():
    """Model.store / restore (treelikelihood.c:116-161): a rejected proposal returns to the stored lnL WITHOUT recomputation."""
    pb = _synthetic_problem(25, 300, 4, 4, seed=12)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    lnl0, g0 = tlk.calculate(), tlk.gradient()
    tlk.store()
    # proposal: new branch lengths, a new substitution model and a new site model
    m2 = models.gtr([0.2, 0.1, 0.3, 0.1, 0.2, 0.1], [0.3, 0.2, 0.2, 0.3])
    r2, p2 = models.discrete_gamma(1.7, 4)
    tlk.set_branch_lengths(pb.bl * 1.3)
    tlk.set_eigen(m2.evec, m2.eval, m2.ivec)
    tlk.set_frequencies(m2.freqs)
    tlk.set_site_model(r2, p2)
    lnl1 = tlk.calculate()
    assert lnl1 != lnl0
    n0 = tlk.launch_count()
    tlk.restore()
    assert tlk.calculate() == lnl0 and tlk.launch_count() == n0, "restore must not recompute"
    g = tlk.gradient()  # the gradient buffer belonged to the rejected state: recomputed from the restored inputs
    assert tlk.launch_count() > n0 and grad_err(g, g0) < 1e-13 and tlk.calculate() == lnl0
    # accept path: store after a move, then a second reject comes back to the moved state
    tlk.set_branch_length(5, pb.bl[5] * 2.0)
    lnl2 = tlk.calculate()
    tlk.store()
    tlk.set_branch_length(7, pb.bl[7] * 0.5)
    assert tlk.calculate() != lnl2
    tlk.restore()
    assert tlk.calculate() == lnl2
    pb.bl[5] *= 2.0
    assert rel_err(lnl2, O.evaluate(pb, gradient=False)["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    tlk.close()
    fresh = phb.SingleTreeLikelihood.from_problem(pb)
    with pytest.raises(phb.PhysherB200Error, match="without phb_tlk_store"):
        fresh.restore()
    fresh.close()


def test_unrooted_convention_and_root_entries():
    pb = _synthetic_problem(20, 128, 4, 4, seed=10)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    g = tlk.gradient()
    assert g[pb.root] == 0.0 and g[pb.right[pb.root]] == 0.0  # treelikelihood.c:3249-3255
    tlk.close()


def test_underflow_switches_rescaling_on():
    """-inf lnL => rescaling on and everything recomputed (treelikelihood.c:1496-1519)."""
    T = 600
    topo = syn.caterpillar_topology(T)
    pb = _synthetic_problem(T, 40, 4, 1, seed=3, topo=topo, unknown=0.0)
    pb.tip_states = syn.random_patterns(T, 40, 4, 0.75, 4)
    pb.bl = syn.random_branch_lengths(topo, 6, 0.8, 1.6)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    assert not tlk.rescaling()
    lnl = tlk.calculate()
    assert tlk.rescaling() and np.isfinite(lnl)
    pb.scale = True
    assert rel_err(lnl, O.evaluate(pb, gradient=False)["lnl"]) < RTOL
    tlk.close()


def test_negative_branch_length_is_rejected():
    pb = _synthetic_problem(8, 32, 4, 1, seed=2)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    bl = pb.bl.copy()
    bl[1] = -0.1
    with pytest.raises(phb.PhysherB200Error, match="branch length"):
        tlk.set_branch_lengths(bl)
    tlk.close()


def test_batched_branch_length_samples():
    """B samples sharing topology and patterns (BASELINE config 3 shape, scaled down)."""
    pb = _synthetic_problem(30, 500, 4, 4, seed=21)
    rng = np.random.default_rng(22)
    B = 5
    bls = pb.bl[None, :] * rng.lognormal(0, 0.1, size=(B, pb.nnodes))
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    lnl, grad = tlk.gradient_batch(bls)
    for b in range(B):
        pb.bl = bls[b]
        want = O.evaluate(pb)
        assert rel_err(lnl[b], want["lnl"]) < RTOL
        assert grad_err(grad[b], want["grad"]) < RTOL
    tlk.close()


@pytest.mark.parametrize("shape", [(12, 6400, 4, 7), (10, 100, 4, 600), (9, 70, 1, 500), (16, 1000, 2, 3)])
def test_batched_samples_share_one_launch(shape):
    """BASELINE config 3: many branch-length samples in ONE fused launch.  Work items are (sample, pattern tile) pairs in
    contiguous per-CTA ranges, so these shapes make single CTAs span 2, 3 and more samples; every sample is checked."""
    T, P, C, B = shape
    pb = _synthetic_problem(T, P, 4, C, seed=600 + P)
    rng = np.random.default_rng(601)
    bls = pb.bl[None, :] * rng.lognormal(0, 0.1, size=(B, pb.nnodes))
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    n0 = tlk.launch_count()
    lnl, grad = tlk.gradient_batch(bls)
    assert tlk.launch_count() - n0 <= 4, "the batch must not be a loop of per-sample launches"
    lnl2, grad2 = tlk.gradient_batch(bls)
    assert np.array_equal(lnl, lnl2) and np.array_equal(grad, grad2), "fixed-order reductions: bit-identical reruns"
    for b in range(B) if B <= 8 else list(range(0, B, 37)) + [B - 1]:
        pb.bl = bls[b]
        want = O.evaluate(pb)
        assert rel_err(lnl[b], want["lnl"]) < RTOL
        assert grad_err(grad[b], want["grad"]) < RTOL
    # lnL only
    lnl3, none = tlk.gradient_batch(bls, want_gradient=False)
    assert none is None and np.allclose(lnl3, lnl, rtol=1e-13, atol=0)
    tlk.close()


# ---------------------------------------------------------------------------------------------
# FP64 tensor-core path (20 / 61 states): KERNELS_AUTO routes there, KERNELS_GENERIC is the cross-check
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("shape", [(24, 1000, 20, 4), (24, 63, 20, 4), (31, 65, 20, 1), (16, 257, 61, 1), (11, 31, 61, 2), (40, 129, 61, 1)])
def test_tensor_core_path_against_oracle(shape):
    """Ragged pattern counts around the 32 / 64-pattern tiles, deeper trees, unknown states."""
    T, P, S, C = shape
    pb = _synthetic_problem(T, P, S, C, seed=4000 + T + P, unknown=0.03)
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_AUTO)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    np.testing.assert_allclose(tlk.pattern_log_likelihoods(), want["pattern_lnl"], rtol=1e-11, atol=0)
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    cg = tlk.cat_branch_gradient()
    assert grad_err(cg.ravel(), np.where(np.arange(pb.nnodes)[:, None] == pb.root, 0, want["cat_grad"]).ravel()) < RTOL
    # the reference's default request folds pi into the root's children (SURVEY 0.4 iii)
    pb.include_root_freqs = True
    tlk.set_option(OPT_INCLUDE_ROOT_FREQS, 1)
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    tlk.close()


@pytest.mark.parametrize("S", [60, 62, 63])
def test_tensor_core_path_other_codon_tables(S):
    """Genetic codes with 4, 2 or 1 stop codons (60 / 62 / 63 sense codons; the reference dispatches every state count >= 60 to its
    codon kernels, treelikelihood.c:1086-1090) run on the tensor cores like the universal code's 61."""
    pb = _synthetic_problem(11, 97, S, 2 if S == 62 else 1, seed=4050 + S, unknown=0.03)
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_AUTO)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    assert tlk.last_kernels() == phb.treelikelihood.RAN_TENSOR
    tlk.use_rescaling(True)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < 1e-9
    tlk.close()


@pytest.mark.parametrize("S", [20, 61])
def test_tensor_core_path_tip_partials(S):
    pb = _synthetic_problem(13, 200, S, 2 if S == 20 else 1, seed=4100 + S, unknown=0.0)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    a, ga = tlk.calculate(), tlk.gradient()
    tlk.close()
    pb.use_tip_states = False
    pb.tip_partials = np.eye(S)[pb.tip_states]
    pb.tip_partials[3, 5, :] = 1.0  # one fully ambiguous entry
    pb.tip_partials[7, 9, : S // 2] = 1.0  # and one ambiguity set
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    assert abs(want["lnl"] - a) > 0  # the ambiguity really changed the model
    tlk.close()


def test_codon_tip_partials_of_single_states_take_the_message_kernels():
    """0/1 tip partials (phycpp's default tip mode) are encoded to states for the message-form tensor-core kernels; an ambiguity
    SET keeps the evaluation on the kernels that read the partials"""
    pb = _synthetic_problem(13, 150, 61, 1, seed=4150, unknown=0.0)
    pb.use_tip_states = False
    pb.tip_partials = np.eye(61)[pb.tip_states]
    pb.tip_partials[3, 5, :] = 1.0  # a gap
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    assert tlk.last_kernels() == phb.treelikelihood.RAN_TENSOR
    pb.tip_partials[7, 9, :30] = 1.0  # the encoding is redone after a tip upload and now declines
    tlk.set_tip_partials(pb.tip_partials)
    want = O.evaluate(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    tlk.close()


@pytest.mark.parametrize("S,C", [(20, 4), (61, 1)])
def test_tensor_core_path_rescaling(S, C):
    """Deep caterpillar with long branches: partials underflow 1e-40, rescaling really triggers."""
    T = 90
    topo = syn.caterpillar_topology(T)
    pb = _synthetic_problem(T, 70, S, C, seed=4200 + S, topo=topo, unknown=0.0)
    pb.tip_states = syn.random_patterns(T, 70, S, 0.8, 4201)
    pb.bl = syn.random_branch_lengths(topo, 4202, 0.6, 1.4)
    base = O.evaluate(pb)
    pb.scale = True
    want = O.evaluate(pb, partials=True)
    assert (want["scaling"][pb.root] < 0).any()
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    g = tlk.gradient()
    assert grad_err(g, want["grad"]) < RTOL
    assert grad_err(g, base["grad"]) < 1e-9
    tlk.set_option(OPT_COMPAT_SCALED_GRADIENT, 1)
    pb.compat_scaled_gradient = True
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    tlk.close()


def test_tensor_core_path_matches_generic_partials():
    """Lower and (internal-node) upper partials of the tensor-core path against the node-at-a-time kernels."""
    pb = _synthetic_problem(10, 100, 20, 2, seed=4300)
    out = O.evaluate(pb, partials=True)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.gradient()
    for n in range(pb.ntips, pb.nnodes):
        np.testing.assert_allclose(tlk.get_partials(n), out["lower"][n], rtol=1e-11, atol=0)
        if n != pb.root:
            np.testing.assert_allclose(tlk.get_partials(pb.nnodes + n), out["upper"][n], rtol=1e-11, atol=1e-300)
    tlk.close()


@pytest.mark.parametrize("kernels", KERNELS, ids=KIDS)
@pytest.mark.parametrize("shape", [(2, 50, 4, 4), (3, 70, 4, 1), (2, 33, 20, 2), (3, 40, 61, 1), (2, 10, 5, 2), (4, 1, 4, 8)], ids=lambda s: "T%d-S%d-C%d" % (s[0], s[2], s[3]))
def test_smallest_trees(shape, kernels):
    """two and three taxa (the root's children are tips; one internal node besides the root at most), a single pattern, 8 categories"""
    T, P, S, C = shape
    pb = _synthetic_problem(T, P, S, C, seed=3000 + T * 7 + S, unknown=0.05)
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=kernels)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    M = np.random.default_rng(3100 + T).normal(size=(2, pb.nnodes, C, S, S))
    assert grad_err(tlk.matrix_gradient(M), O.matrix_gradient(pb, M)) < RTOL
    tlk.use_rescaling(True)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < 1e-9
    tlk.close()


def _gy94_problem(T=12, P=160, seed=8100):
    """GY94 (kappa 2.5, omega 0.3): codons two or three changes apart have transition probabilities of order t^2, t^3 -- the entries
    a one-ulp difference in an exponential moves by 1e-9 relative."""
    topo = syn.random_topology(T, seed)
    m = models.gy94(2.5, 0.3)
    bl = syn.random_branch_lengths(topo, seed + 3)
    # data evolved down the tree (bench.py's generator): columns a codon model can explain without piling changes on one short branch
    return O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=61,
                     tip_states=syn.simulate_patterns(topo, bl * 0.35, P, 61, seed + 1, unknown_frac=0.01),
                     weights=np.random.default_rng(seed + 2).integers(1, 4, P).astype(np.float64), freqs=m.freqs, rates=np.ones(1), props=np.ones(1),
                     bl=bl, evec=m.evec, eval=m.eval, ivec=m.ivec)


def test_codon_parity_holds_on_every_entry_point():
    """1e-10 on codon models needs the exponentials of the host's libm (PHB_OPT_HOST_EXPONENTIALS, default for >= 60 states): the
    batched, the time-tree and the single-branch entry points take them like the plain evaluation does."""
    pb = _gy94_problem()
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    want = O.evaluate(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL and grad_err(tlk.gradient(), want["grad"]) < RTOL
    assert tlk.last_kernels() == phb.treelikelihood.RAN_TENSOR
    # phb_tlk_gradient_batch
    rng = np.random.default_rng(8)
    bls = pb.bl[None, :] * rng.lognormal(0.0, 0.1, size=(3, pb.nnodes))
    lnl, grad = tlk.gradient_batch(bls)
    for b in range(3):
        q = _gy94_problem()
        q.bl = bls[b]
        w = O.evaluate(q)
        assert rel_err(lnl[b], w["lnl"]) < RTOL and grad_err(grad[b], w["grad"]) < RTOL
    # phb_tlk_calculate_branch
    from tests.test_branch import deriv_err

    tlk.set_branch_lengths(pb.bl)
    for n in (1, pb.ntips + 2):
        cands = pb.bl[n] * np.array([0.5, 1.0, 1.7])
        l0, d1, d2 = tlk.calculate_branch(n, cands)
        assert deriv_err(np.stack([l0, d1, d2], 1), O.branch_derivatives(pb, n, cands)) < RTOL
    # phb_tlk_gradient_batch_time (branch lengths are built on the device and fetched for the host's exp)
    T, N = pb.ntips, pb.nnodes
    tip_heights = rng.uniform(0.0, 0.5, T)
    ratios = rng.uniform(0.3, 0.9, size=(2, T - 1))
    ratios[:, -1] = tip_heights.max() + rng.uniform(0.5, 1.0, size=2)
    rates = np.full((2, 1), 0.05)
    pb.unrooted = False
    tlk.set_option(phb.treelikelihood.OPT_UNROOTED, 0)
    tlk.set_time_tree(tip_heights)
    lt, lj, gr, gc = tlk.gradient_batch_time(ratios, rates, include_jacobian=True)
    for b in range(2):
        w = O.time_evaluate(pb, tip_heights, ratios[b], rates[b], include_jacobian=True)
        assert rel_err(lt[b], w["lnl"]) < RTOL and grad_err(gr[b], w["grad_ratios"]) < RTOL and grad_err(gc[b], w["grad_rates"]) < RTOL
    tlk.close()


def test_evaluations_invalidate_what_matrix_gradient_left_behind():
    """matrix_gradient -> gradient -> root_frequency_gradient must not reuse statistics / root partials that the
    gradient call overwrote or outdated (4 states: fused walk statistics; the walk never writes d_lower)."""
    pb = _synthetic_problem(14, 300, 4, 4, seed=8200)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    M = np.random.default_rng(3).normal(size=(2, pb.nnodes, pb.ncat, 4, 4))
    tlk.matrix_gradient(M)
    first = tlk.root_frequency_gradient()
    tlk.gradient()  # same inputs: a fused walk WITHOUT statistics takes the place of the one that left them
    assert grad_err(tlk.root_frequency_gradient(), first) < 1e-12
    pb.bl = pb.bl * 1.3
    tlk.set_branch_lengths(pb.bl)
    tlk.gradient()  # a fused walk without statistics
    got = tlk.root_frequency_gradient()
    gen = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_GENERIC)
    want = gen.root_frequency_gradient()
    gen.close()
    assert grad_err(got, want) < RTOL and grad_err(got, first) > 1e-6
    tlk.close()


def test_incremental_reads_see_the_stale_ancestors_recomputed():
    """incremental mode, set_branch_length + gradient (fused walk: the resident buffers are not touched), then
    get_partials / root_frequency_gradient must recompute the dirty ancestors first."""
    pb = _synthetic_problem(12, 200, 4, 2, seed=8300)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.set_option(phb.OPT_INCREMENTAL, 1)
    tlk.update_uppers()
    node = 3
    pb.bl[node] *= 2.0
    tlk.set_branch_length(node, pb.bl[node])
    want = O.evaluate(pb, partials=True)
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    gen = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_GENERIC)
    anc = int(pb.parent[node])
    np.testing.assert_allclose(tlk.get_partials(anc), gen.get_partials(anc), rtol=1e-12, atol=0)
    np.testing.assert_allclose(tlk.get_partials(pb.root), gen.get_partials(pb.root), rtol=1e-12, atol=0)
    assert grad_err(tlk.root_frequency_gradient(), gen.root_frequency_gradient()) < RTOL
    gen.close()
    tlk.close()


def _eval_tuned(pb, tune):
    from physher_b200.treelikelihood import OPT_TUNE
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.set_option(OPT_TUNE, tune)
    lnl, plnl, g, launches = tlk.calculate(), tlk.pattern_log_likelihoods().copy(), tlk.gradient().copy(), tlk.launch_count()
    assert tlk.last_kernels() == phb.treelikelihood.RAN_TENSOR
    tlk.close()
    return lnl, plnl, g, launches


@pytest.mark.parametrize("shape", [(61, 9, 300, 2), (61, 2, 50, 1), (60, 6, 33, 1), (63, 12, 129, 1), (61, 7, 77, 2)], ids=lambda s: "S%d-T%d-P%d-C%d" % s)
def test_cherry_tables(shape):
    """cherries evaluated once per PAIR of tip states (forced on by PHB_OPT_TUNE 20, off by 21).  Post-order pass: the pair's message
    copied per pattern (k_dmma_cherry_gather) -- the same arithmetic in the same order, so not one bit of any pattern likelihood
    differs.  Pre-order pass: the cherry's three branch terms as dot products of U_n with the pair's coefficient rows
    (k_dmma_cherry_upper) -- another summation order, same gradient to rounding.  Both against the oracle; unknown states included."""
    S, T, P, C = shape
    pb = _synthetic_problem(T, P, S, C, seed=4200 + S + T, unknown=0.05)
    want = O.evaluate(pb)
    a, b = _eval_tuned(pb, 20), _eval_tuned(pb, 21)
    assert rel_err(a[0], want["lnl"]) < RTOL and grad_err(a[2], want["grad"]) < RTOL
    assert rel_err(b[0], want["lnl"]) < RTOL and grad_err(b[2], want["grad"]) < RTOL
    assert a[0] == b[0]
    np.testing.assert_array_equal(a[1], b[1])
    assert grad_err(a[2], b[2]) < 1e-11
    assert a[3] != b[3]  # the table path really ran


def test_cherry_tables_include_root_freqs_and_matrix_gradient():
    """the frequency weights move from the gradient sums into the upper partials (tlk->include_root_freqs); the substitution-model
    sweep runs the pre-order pass once per matrix set on swapped derivative matrices, tables included"""
    from physher_b200.treelikelihood import OPT_TUNE
    pb = _synthetic_problem(10, 120, 61, 1, seed=4290, unknown=0.04)
    pb.include_root_freqs = True
    want = O.evaluate(pb)
    res = {}
    for tune in (20, 21):
        tlk = phb.SingleTreeLikelihood.from_problem(pb)
        tlk.set_option(OPT_TUNE, tune)
        tlk.set_option(OPT_INCLUDE_ROOT_FREQS, 1)
        assert grad_err(tlk.gradient(), want["grad"]) < RTOL
        rng = np.random.default_rng(5)
        M = rng.standard_normal((2, pb.nnodes, len(pb.rates), 61, 61)) * 1e-2
        res[tune] = tlk.matrix_gradient(M)
        tlk.close()
    np.testing.assert_allclose(res[20], res[21], rtol=1e-9, atol=1e-9 * np.abs(res[21]).max())


def test_cherry_tables_switch_on_by_pattern_count_and_take_encoded_tip_partials():
    """20 states on the level-batched kernels (PHB_OPT_TUNE 9): 441 pairs, tables from 1,764 patterns on; 0/1 tip partials"""
    pb = _synthetic_problem(14, 2000, 20, 2, seed=4260, unknown=0.03)
    want = O.evaluate(pb)
    lnl, plnl, g, launches = _eval_tuned(pb, 9)
    assert rel_err(lnl, want["lnl"]) < RTOL and grad_err(g, want["grad"]) < RTOL
    small = _synthetic_problem(14, 1000, 20, 2, seed=4260, unknown=0.03)
    assert _eval_tuned(small, 9)[3] < launches
    pb.use_tip_states = False
    pb.tip_partials = np.eye(21)[np.minimum(pb.tip_states, 20)][:, :, :20]
    pb.tip_partials[pb.tip_states >= 20] = 1.0
    lnl2, plnl2, g2, _ = _eval_tuned(pb, 9)
    assert lnl2 == lnl
    np.testing.assert_array_equal(g2, g)


@pytest.mark.parametrize("tune", [0, 20, 22], ids=["message-form", "message-form+cherry-tables", "node-at-a-time"])
@pytest.mark.parametrize("S,T,P,C", [(61, 14, 130, 1), (61, 9, 75, 2), (20, 17, 300, 4), (63, 8, 40, 1)])
def test_rescaled_message_form(S, T, P, C, tune):
    """Rescaling on the message form of the tensor-core level kernels (k_dmma_lower_msg leaves the row maxima of L_n,
    k_dmma_scale_from_max divides the message; the pre-order ops carry exp(-(sf[N+n] + sf[a] + sf[b]))) with a threshold (1e-2) that
    makes nearly every node rescale, with and without the cherry tables, against the oracle and the node-at-a-time form
    (PHB_OPT_TUNE 22).  SingleTreeLikelihood_scalePartials, treelikelihood.c:1790-1836; gradient_cat_branch_lengths :2715-2789."""
    from physher_b200.treelikelihood import OPT_SCALING_THRESHOLD_EXP, OPT_TUNE
    pb = _synthetic_problem(T, P, S, C, seed=4400 + S + T, unknown=0.04)
    base = O.evaluate(pb)
    pb.scale, pb.scaling_threshold = True, 1e-2
    want = O.evaluate(pb, partials=True)
    assert (want["scaling"][pb.root] < 0).any()
    tlk = phb.SingleTreeLikelihood.from_problem(pb, kernels=phb.KERNELS_FUSED if S != 20 else phb.KERNELS_AUTO)
    tlk.set_option(OPT_SCALING_THRESHOLD_EXP, 2)
    tlk.set_option(OPT_TUNE, 9 if (S == 20 and tune == 0) else tune)  # 20 states: 9 = the level kernels (the walk does not rescale anyway)
    assert tlk.rescaling()
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    np.testing.assert_allclose(tlk.pattern_log_likelihoods(), want["pattern_lnl"], rtol=1e-11, atol=0)
    g = tlk.gradient()
    assert tlk.last_kernels() == phb.treelikelihood.RAN_TENSOR
    assert grad_err(g, want["grad"]) < RTOL
    assert grad_err(g, base["grad"]) < 1e-9
    tlk.use_rescaling(False)
    assert grad_err(tlk.gradient(), base["grad"]) < RTOL  # and back
    tlk.close()
