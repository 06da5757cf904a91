"""BASELINE.json's configurations at FULL size on the GPU, checked through properties that do not need the oracle to run the whole
workload (it would take hours): additivity over pattern shards with one shard pinned on the oracle, exact weight linearity, the
per-pattern checksum, invariance under a pattern permutation, central finite differences of lnL against the analytic gradient, and
agreement between kernel families.  Inputs are bench.py's (same seeds, same generator)."""
import numpy as np
import pytest

import bench
import physher_b200 as phb
from oracle import oracle as O
from tests.util import FLOOR_LARGE, RTOL, grad_err, rel_err

pytestmark = pytest.mark.gpu


def _make(cfg, patterns, weights, kernels=phb.KERNELS_AUTO, bl=None, inputs=None):
    topo, bl0, m, rates, props = inputs
    tlk = phb.SingleTreeLikelihood(topo.left, topo.right, topo.root, cfg["states"], cfg["cats"], patterns.shape[1], use_tip_states=True, device=0)
    tlk.set_option(phb.treelikelihood.OPT_KERNELS, kernels)
    tlk.set_tip_states(patterns)
    tlk.set_pattern_weights(weights)
    tlk.set_eigen(m.evec, m.eval, m.ivec)
    tlk.set_frequencies(m.freqs)
    tlk.set_site_model(rates, props)
    tlk.set_branch_lengths(bl0 if bl is None else bl)
    return tlk


def _fd_check(tlk, bl, grad, topo, rng, nbranches=3, h=1e-6):
    cands = [n for n in range(bl.size) if n != topo.root and n != topo.right[topo.root]]
    for n in rng.choice(cands, nbranches, replace=False):
        up, dn = bl.copy(), bl.copy()
        up[n] += h
        dn[n] -= h
        tlk.set_branch_lengths(up)
        lu = tlk.calculate()
        tlk.set_branch_lengths(dn)
        ld = tlk.calculate()
        fd = (lu - ld) / (2 * h)
        # truncation ~ h^2, rounding of the two lnL values ~ a few ulp of |lnL| divided by h
        assert abs(fd - grad[n]) <= 2e-5 * abs(grad[n]) + 8e-15 * abs(lu) / h, (int(n), fd, grad[n])
    tlk.set_branch_lengths(bl)


def test_c2_full_size_properties():
    """GTR+G4, 1000 taxa x 100,000 patterns (BASELINE config 2)"""
    cfg = dict(bench.CONFIGS["c2"])
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    inputs = (topo, bl, m, rates, props)
    P, N = cfg["patterns"], 2 * cfg["taxa"] - 1
    rng = np.random.default_rng(11)
    weights = rng.integers(1, 4, P).astype(np.float64)  # non-trivial multiplicities
    tlk = _make(cfg, patterns, weights, inputs=inputs)
    lnl, g = tlk.calculate(), tlk.gradient().copy()
    assert np.isfinite(lnl) and np.isfinite(g).all() and not tlk.rescaling()
    # checksum of checksums: the weighted sum of the per-pattern log likelihoods is lnL
    plk = tlk.pattern_log_likelihoods()
    assert rel_err(float(np.dot(weights, plk)), lnl) < 1e-12
    # additivity over pattern shards; the first shard (400 patterns) is pinned on the oracle
    edges = [0, 400, 33_333, 70_001, P]
    acc_l, acc_g = 0.0, np.zeros(N)
    for k, (b, e) in enumerate(zip(edges, edges[1:])):
        sub = _make(cfg, np.ascontiguousarray(patterns[:, b:e]), weights[b:e], inputs=inputs)
        sl, sg = sub.calculate(), sub.gradient().copy()
        if k == 0:
            pb = O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=4, tip_states=np.ascontiguousarray(patterns[:, b:e]),
                           weights=weights[b:e], freqs=m.freqs, rates=rates, props=props, bl=bl, evec=m.evec, eval=m.eval, ivec=m.ivec)
            want = O.evaluate(pb)
            assert rel_err(sl, want["lnl"]) < RTOL and grad_err(sg, want["grad"]) < RTOL
            assert np.max(np.abs(plk[b:e] - want["pattern_lnl"]) / np.abs(want["pattern_lnl"])) < RTOL
        acc_l += sl
        acc_g += sg
        sub.close()
    assert rel_err(acc_l, lnl) < 1e-12 and grad_err(acc_g, g) < 1e-11
    # exact linearity in the weights (a factor 2 is exact in binary floating point)
    tlk.set_pattern_weights(2.0 * weights)
    assert tlk.calculate() == 2.0 * lnl and np.array_equal(tlk.gradient(), 2.0 * g)
    tlk.set_pattern_weights(weights)
    # central finite differences on three branches
    _fd_check(tlk, bl, g, topo, rng)
    tlk.close()
    # a permutation of the patterns only reorders the sums
    perm = rng.permutation(P)
    shuf = _make(cfg, np.ascontiguousarray(patterns[:, perm]), weights[perm], inputs=inputs)
    assert rel_err(shuf.calculate(), lnl) < 1e-12 and grad_err(shuf.gradient(), g) < 1e-11
    shuf.close()
    # the node-at-a-time kernels agree at full size (38 GB of partials)
    gen = _make(cfg, patterns, weights, kernels=phb.KERNELS_GENERIC, inputs=inputs)
    assert rel_err(gen.calculate(), lnl) < 1e-12 and grad_err(gen.gradient(), g) < RTOL
    gen.close()


@pytest.mark.parametrize("name", ["c4", "c5"])
def test_tensor_core_full_size_properties(name):
    """LG+G4 200 taxa x 200,000 patterns and GY94 100 taxa x 1,000,000 patterns (BASELINE configs 4 and 5) on one GPU"""
    cfg = dict(bench.CONFIGS[name])
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    inputs = (topo, bl, m, rates, props)
    P, N, S = cfg["patterns"], 2 * cfg["taxa"] - 1, cfg["states"]
    rng = np.random.default_rng(12)
    weights = rng.integers(1, 4, P).astype(np.float64)
    tlk = _make(cfg, patterns, weights, inputs=inputs)
    lnl, g = tlk.calculate(), tlk.gradient().copy()
    assert np.isfinite(lnl) and np.isfinite(g).all()
    plk = tlk.pattern_log_likelihoods()
    assert rel_err(float(np.dot(weights, plk)), lnl) < 1e-12
    tlk.set_pattern_weights(2.0 * weights)
    assert tlk.calculate() == 2.0 * lnl and np.array_equal(tlk.gradient(), 2.0 * g)
    tlk.set_pattern_weights(weights)
    _fd_check(tlk, bl, g, topo, rng, nbranches=2)
    tlk.close()
    # a 300-pattern shard against the oracle, and its share of the full evaluation
    b, e = P // 2, P // 2 + 300
    pb = O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=S, tip_states=np.ascontiguousarray(patterns[:, b:e]),
                   weights=weights[b:e], freqs=m.freqs, rates=rates, props=props, bl=bl, evec=m.evec, eval=m.eval, ivec=m.ivec)
    want = O.evaluate(pb)
    assert np.max(np.abs(plk[b:e] - want["pattern_lnl"]) / np.abs(want["pattern_lnl"])) < RTOL
    sub = _make(cfg, np.ascontiguousarray(patterns[:, b:e]), weights[b:e], inputs=inputs)
    assert rel_err(sub.calculate(), want["lnl"]) < RTOL and grad_err(sub.gradient(), want["grad"]) < RTOL
    if S == 61:
        # Codons two or three changes apart have transition probabilities ~ t^2, t^3 that come out of the eigen sum by cancellation:
        # one ulp in an exponential moves them by 1e-9 relative.  The 1e-10 above holds because the exponentials come from the host's
        # libm, the exp the reference (and the oracle) calls -- PHB_OPT_HOST_EXPONENTIALS, on by default for >= 60 states.  With the
        # device's own (equally accurate) exp the same evaluation sits at the conditioning of the inputs, not at 1e-10:
        sub.set_option(phb.treelikelihood.OPT_HOST_EXPONENTIALS, 0)
        off = grad_err(sub.gradient(), want["grad"])
        assert rel_err(sub.calculate(), want["lnl"]) < 1e-11 and off < 1e-6
        print(f"codon gradient vs oracle with the device's exp: {off:.2e}")
    sub.close()


def test_c2_one_million_patterns_properties():
    """GTR+G4, 1000 taxa x 1,000,000 patterns: the configuration the north-star's >= 50 % roofline bar is quoted on.  The fused walk
    keeps no partials, so the whole job needs ~3.5 GB of device memory (the reference would allocate 512 GB)."""
    cfg = dict(bench.CONFIGS["c2_1m"])
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    inputs = (topo, bl, m, rates, props)
    P, N = cfg["patterns"], 2 * cfg["taxa"] - 1
    rng = np.random.default_rng(21)
    weights = rng.integers(1, 4, P).astype(np.float64)
    tlk = _make(cfg, patterns, weights, inputs=inputs)
    lnl, g = tlk.calculate(), tlk.gradient().copy()
    assert tlk.last_kernels() == phb.treelikelihood.RAN_WALK
    assert np.isfinite(lnl) and np.isfinite(g).all() and not tlk.rescaling()
    plk = tlk.pattern_log_likelihoods()
    assert rel_err(float(np.dot(weights, plk)), lnl) < 1e-12
    # additivity over ragged pattern shards; the first (400 patterns) and a middle one (333 patterns) are pinned on the oracle
    edges = [0, 400, 250_001, 250_334, 777_777, P]
    acc_l, acc_g = 0.0, np.zeros(N)
    for k, (b, e) in enumerate(zip(edges, edges[1:])):
        sub = _make(cfg, np.ascontiguousarray(patterns[:, b:e]), weights[b:e], inputs=inputs)
        sl, sg = sub.calculate(), sub.gradient().copy()
        if e - b < 1000:
            pb = O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=4, tip_states=np.ascontiguousarray(patterns[:, b:e]),
                           weights=weights[b:e], freqs=m.freqs, rates=rates, props=props, bl=bl, evec=m.evec, eval=m.eval, ivec=m.ivec)
            want = O.evaluate(pb)
            assert rel_err(sl, want["lnl"]) < RTOL and grad_err(sg, want["grad"]) < RTOL
            assert np.max(np.abs(plk[b:e] - want["pattern_lnl"]) / np.abs(want["pattern_lnl"])) < RTOL
        acc_l += sl
        acc_g += sg
        sub.close()
    assert rel_err(acc_l, lnl) < 1e-12 and grad_err(acc_g, g) < 1e-11
    tlk.set_pattern_weights(2.0 * weights)
    assert tlk.calculate() == 2.0 * lnl and np.array_equal(tlk.gradient(), 2.0 * g)
    tlk.set_pattern_weights(weights)
    _fd_check(tlk, bl, g, topo, rng, nbranches=2)
    tlk.close()


def test_c3_full_size_batch():
    """HKY+G4, 500 taxa x 50,000 patterns x 128 branch-length samples (BASELINE config 3) through phb_tlk_gradient_batch: ONE fused
    launch for the whole batch; three samples pinned on oracle shards, additivity over pattern shards per sample, and agreement
    with the single-sample entry point."""
    cfg = dict(bench.CONFIGS["c3"])
    topo, bl, m, rates, props, patterns, weights = bench.make_inputs(cfg, 0)
    inputs = (topo, bl, m, rates, props)
    P, N, B = cfg["patterns"], 2 * cfg["taxa"] - 1, cfg["batch"]
    rng = np.random.default_rng(31)
    weights = rng.integers(1, 4, P).astype(np.float64)
    bls = bl[None, :] * rng.lognormal(0.0, 0.1, size=(B, N))
    bls[:, topo.root] = 0.0
    bls[:, topo.right[topo.root]] = 0.0
    tlk = _make(cfg, patterns, weights, inputs=inputs)
    tlk.gradient_batch(bls[:2])  # tip encoding and scratch allocation happen once
    n0 = tlk.launch_count()
    lnls, grads = tlk.gradient_batch(bls)
    assert tlk.launch_count() - n0 == 3, "matrices of all samples, ONE walk over (sample, pattern tile) items, fixed-order finalize"
    assert tlk.last_kernels() == phb.treelikelihood.RAN_WALK
    assert np.isfinite(lnls).all() and np.isfinite(grads).all() and not tlk.rescaling()
    # the single-sample entry point sees the same numbers
    for k in (5, 127):
        tlk.set_branch_lengths(bls[k])
        assert rel_err(tlk.calculate(), lnls[k]) < 1e-12
        assert grad_err(tlk.gradient(), grads[k]) < RTOL  # another CTA -> tile assignment: another summation order
    tlk.close()
    # pattern shards, three samples at once; the 1500-pattern shards are pinned on the oracle
    pick = [0, 63, 127]
    edges = [0, 1500, 20_000, 21_500, P]
    acc_l, acc_g = np.zeros(3), np.zeros((3, N))
    for b, e in zip(edges, edges[1:]):
        sub = _make(cfg, np.ascontiguousarray(patterns[:, b:e]), weights[b:e], inputs=inputs)
        sl, sg = sub.gradient_batch(bls[pick])
        if e - b == 1500:  # (small shards leave branch gradients that are near-cancelling sums: 1e-10 of THEIR value is below the summation noise)
            for i, k in enumerate(pick):
                pb = O.Problem(left=topo.left, right=topo.right, parent=topo.parent, root=topo.root, nstate=4, tip_states=np.ascontiguousarray(patterns[:, b:e]),
                               weights=weights[b:e], freqs=m.freqs, rates=rates, props=props, bl=bls[k], evec=m.evec, eval=m.eval, ivec=m.ivec)
                want = O.evaluate(pb)
                assert rel_err(sl[i], want["lnl"]) < RTOL
                assert grad_err(sg[i], want["grad"], floor=FLOOR_LARGE) < RTOL
                assert grad_err(sg[i], want["grad"]) < 1e-8  # near-cancelling entries: within the noise of the reference's own exp(log L_k)
        acc_l += sl
        acc_g += sg
        sub.close()
    for i, k in enumerate(pick):
        assert rel_err(acc_l[i], lnls[k]) < 1e-12
        assert grad_err(acc_g[i], grads[k]) < RTOL
