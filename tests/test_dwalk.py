"""Whole-tree walk on the FP64 tensor cores (physher_b200/csrc/phb_dwalk.cu, 20 states) against the CPU oracle and against the
level-batched tensor-core kernels it replaces (PHB_OPT_TUNE 9).  What the reference computes here: update_partials_20_SSE
(treelikelihood20.c:114-647), update_upper_partials (treelikelihood.c:2129-2162), calculate_branch_partials_20_SSE
(treelikelihood20.c:834-1025), gradient_cat_branch_lengths (treelikelihood.c:2793-2941)."""
import numpy as np
import pytest

import physher_b200 as phb
from oracle import oracle as O
from physher_b200 import models, synthetic as syn
from physher_b200.treelikelihood import OPT_INCLUDE_ROOT_FREQS, OPT_TUNE, RAN_TENSOR

pytestmark = pytest.mark.gpu
RTOL = 1e-10  # north_star: lnL and every branch gradient within 1e-10 relative

TUNE_LEVELS, TUNE_ONE_SLOT, TUNE_8_WARPS, TUNE_ONE_SLOT_8_WARPS, TUNE_4_WARPS, TUNE_TURNS, TUNE_12_NARROW_WARPS = 9, 11, 12, 13, 14, 15, 16


def rel_err(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def grad_err(g, want):
    scale = np.maximum(np.abs(want), 1e-6 * np.abs(want).max())
    return float(np.max(np.abs(g - want) / scale))


def problem(T, P, C, seed, topo=None, unknown=0.02):
    topo = topo or syn.random_topology(T, seed)
    m = models.random_reversible(20, seed + 5)
    rates, props = models.discrete_gamma(0.5, C)
    return O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=20,
                     tip_states=syn.random_patterns(T, P, 20, 0.15, seed + 1, unknown_frac=unknown),
                     weights=np.random.default_rng(seed + 2).integers(1, 5, P).astype(np.float64),
                     freqs=m.freqs, rates=rates, props=props, bl=syn.random_branch_lengths(topo, seed + 3),
                     evec=m.evec, eval=m.eval, ivec=m.ivec)


def check(pb, tune, want=None):
    want = want or O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.set_option(OPT_TUNE, tune)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    np.testing.assert_allclose(tlk.pattern_log_likelihoods(), want["pattern_lnl"], rtol=1e-11, atol=0)
    g = tlk.gradient()
    assert tlk.last_kernels() == RAN_TENSOR
    assert grad_err(g, want["grad"]) < RTOL
    cg = tlk.cat_branch_gradient()
    assert grad_err(cg.ravel(), np.where(np.arange(pb.nnodes)[:, None] == pb.root, 0, want["cat_grad"]).ravel()) < RTOL
    launches = tlk.launch_count()
    tlk.close()
    return g, launches


@pytest.mark.parametrize("tune", [0, TUNE_8_WARPS, TUNE_ONE_SLOT, TUNE_ONE_SLOT_8_WARPS, TUNE_TURNS, TUNE_12_NARROW_WARPS],
                         ids=["auto", "8warps", "spill", "spill8", "turns", "12x8"])
@pytest.mark.parametrize("shape", [(24, 1000, 4), (24, 63, 4), (31, 65, 1), (57, 129, 2), (2, 40, 2), (3, 17, 4), (9, 1, 4), (10, 5, 8)],
                         ids=lambda s: "T%d-P%d-C%d" % s)
def test_walk_against_oracle(shape, tune):
    """ragged pattern counts around the 64 / 128-pattern tiles, unknown states, two and three taxa; every launch geometry"""
    T, P, C = shape
    check(problem(T, P, C, seed=9000 + T + P), tune)


@pytest.mark.parametrize("tune", [0, TUNE_8_WARPS, TUNE_ONE_SLOT_8_WARPS], ids=["auto", "8warps", "spill8"])
@pytest.mark.parametrize("kind", ["caterpillar", "balanced"])
def test_walk_extreme_topologies(kind, tune):
    """a chain (every op hands over to the next) and a perfect tree (the most parked values: at 64 taxa and 8 warps they fit only beside a
    TWO-stage image ring, the geometry no other test reaches)"""
    T = 64
    topo = syn.caterpillar_topology(T) if kind == "caterpillar" else syn.balanced_topology(T)
    check(problem(T, 300, 2, seed=9100, topo=topo), tune)


def test_walk_equals_level_kernels_and_uses_fewer_launches():
    pb = problem(40, 2000, 4, seed=9200)
    want = O.evaluate(pb)
    g_walk, l_walk = check(pb, 0, want)
    g_lvl, l_lvl = check(pb, TUNE_LEVELS, want)
    assert grad_err(g_walk, g_lvl) < 1e-12
    assert l_walk < l_lvl


def test_walk_include_root_freqs_and_repeat():
    pb = problem(19, 500, 4, seed=9300)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    g1 = tlk.gradient().copy()
    bl2 = pb.bl * 1.1
    tlk.set_branch_lengths(bl2)
    tlk.gradient()
    tlk.set_branch_lengths(pb.bl)
    np.testing.assert_array_equal(tlk.gradient(), g1)  # warp-private rows, fixed-order sums: run-to-run identical
    pb.include_root_freqs = True
    tlk.set_option(OPT_INCLUDE_ROOT_FREQS, 1)
    assert grad_err(tlk.gradient(), O.evaluate(pb)["grad"]) < RTOL
    tlk.close()


def test_walk_takes_tip_partials_of_single_states():
    """0/1 tip partials (phycpp's default, examples/fluA/*.json) encode to state codes; an ambiguity SET declines to the level kernels"""
    pb = problem(13, 200, 2, seed=9400, unknown=0.0)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    a = tlk.calculate()
    tlk.close()
    pb.use_tip_states = False
    pb.tip_partials = np.eye(20)[pb.tip_states]
    pb.tip_partials[3, 5, :] = 1.0  # fully ambiguous: still a code
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    few = tlk.launch_count()
    tlk.close()
    pb.tip_partials[7, 9, :10] = 1.0  # an ambiguity set is a sum of columns, not a gather
    want = O.evaluate(pb)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    assert tlk.launch_count() > few
    assert abs(want["lnl"] - a) > 0
    tlk.close()


def test_walk_full_width_tiles_and_partial_reads_after_it():
    """enough patterns for the 8-warp geometry on its own (items >= 6 x SMs), then get_partials / a changed topology / rescaling"""
    pb = problem(12, 30011, 4, seed=9500)
    want = O.evaluate(pb, partials=True)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    n = pb.ntips + 3
    np.testing.assert_allclose(tlk.get_partials(n), want["lower"][n], rtol=1e-11, atol=0)
    if n != pb.root:
        np.testing.assert_allclose(tlk.get_partials(pb.nnodes + n), want["upper"][n], rtol=1e-11, atol=1e-300)
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    tlk.use_rescaling(True)
    assert rel_err(tlk.calculate(), want["lnl"]) < RTOL
    assert grad_err(tlk.gradient(), want["grad"]) < 1e-9
    tlk.use_rescaling(False)
    assert grad_err(tlk.gradient(), want["grad"]) < RTOL
    tlk.close()


def test_walk_serves_batched_branch_length_samples():
    """phb_tlk_gradient_batch at 20 states: every sample of the batch is one pair of walk launches into its own result slot"""
    pb = problem(14, 333, 2, seed=9600)
    rng = np.random.default_rng(9601)
    bls = pb.bl[None, :] * rng.uniform(0.5, 1.5, size=(3, pb.nnodes))
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    lnl, grad = tlk.gradient_batch(bls)
    assert tlk.last_kernels() == RAN_TENSOR
    for b in range(3):
        pb.bl = bls[b]
        want = O.evaluate(pb)
        assert rel_err(lnl[b], want["lnl"]) < RTOL
        assert grad_err(grad[b], want["grad"]) < RTOL
    tlk.close()


@pytest.mark.parametrize("tune", [0, TUNE_ONE_SLOT_8_WARPS], ids=["auto", "spill8"])
def test_walk_is_repeatable_under_load(tune):
    """150 back-to-back evaluations that alternate between two sets of branch lengths: every result must be bit-identical to the first
    one of its set.  The walk's warps synchronise only through mbarriers (image ring, landing tiles); a lost or doubled arrival shows
    up here as a wrong number or as a hang long before it shows up in a benchmark."""
    pb = problem(40, 5000, 4, seed=9700)
    tlk = phb.SingleTreeLikelihood.from_problem(pb)
    tlk.set_option(OPT_TUNE, tune)
    bls = [pb.bl, pb.bl * 1.07]
    first = []
    for b in bls:
        tlk.set_branch_lengths(b)
        first.append((tlk.calculate(), tlk.gradient().copy()))
    pb2 = problem(40, 5000, 4, seed=9700)
    pb2.bl = bls[1]
    want = O.evaluate(pb2)
    assert rel_err(first[1][0], want["lnl"]) < RTOL and grad_err(first[1][1], want["grad"]) < RTOL
    for i in range(150):
        k = i & 1
        tlk.set_branch_lengths(bls[k])
        g = tlk.gradient()
        assert tlk.calculate() == first[k][0]
        np.testing.assert_array_equal(g, first[k][1])
    tlk.close()
