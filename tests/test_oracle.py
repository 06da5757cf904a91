"""The CPU oracle (oracle/phb_oracle.c) against the reference's known answers and golden outputs.

Pins the oracle before it is trusted as the checker of the CUDA path (task ③):
 * tests/test_tree_likelihood.c:28-40 known answers (C1, JC69 strict clock on fluA);
 * outputs of the unmodified reference compiled by oracle/Makefile (tests/golden/*.npz).
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import RTOL, golden_names, grad_err, load_golden, rel_err


def test_kat_jc69_fluA_lnl_and_clock_gradient():
    pb, z = load_golden("c1_jc69_fluA_tipstates")
    pb.include_root_freqs = True  # what TreeLikelihood_initialize_gradient sets (treelikelihood.c:241)
    out = O.evaluate(pb)
    assert abs(out["lnl"] - float(z["kat_lnl"])) < 1e-8  # tolerance of the reference's own test
    assert rel_err(out["lnl"], float(z["kat_lnl"])) < RTOL
    clock = float((out["grad"] * z["time_elapsed"]).sum())  # gradient_clock, treelikelihood.c:3054-3063
    assert abs(clock - float(z["kat_clock_grad"])) < 1e-8 * 10
    assert rel_err(clock, float(z["kat_clock_grad"])) < RTOL


def test_kat_jc69_fluA_tip_partials_is_a_different_model():
    pb, z = load_golden("c1_jc69_fluA_tippartials")
    out = O.evaluate(pb, gradient=False)
    assert rel_err(out["lnl"], -4777.616437105040) < 1e-12
    assert rel_err(out["lnl"], float(z["ref_lnl"])) < RTOL


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    pb, z = load_golden(name)
    out = O.evaluate(pb, matrices=True)
    assert rel_err(out["lnl"], float(z["ref_lnl"])) < RTOL
    np.testing.assert_allclose(out["pattern_lnl"], z["ref_pattern_lnl"], rtol=1e-11, atol=0)
    keep = np.arange(pb.nnodes) != pb.root
    np.testing.assert_allclose(out["matrices"][keep], z["ref_matrices"][keep], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["dmatrices"][keep], z["ref_dmatrices"][keep], rtol=0, atol=1e-12)
    if "ref_grad_exact" in z:
        assert grad_err(out["grad"], z["ref_grad_exact"]) < RTOL
        pb.include_root_freqs = True
        assert grad_err(O.evaluate(pb)["grad"], z["ref_grad_default"]) < RTOL
        pb.include_root_freqs = False
    if "ref_lnl_scaled" in z:
        pb.scale = True
        pb.compat_scaled_gradient = True
        outs = O.evaluate(pb, partials=True)
        assert rel_err(outs["lnl"], float(z["ref_lnl_scaled"])) < RTOL
        np.testing.assert_allclose(outs["scaling"][pb.root], z["ref_root_scaling"], rtol=1e-12, atol=1e-12)
        if "ref_grad_scaled_compat" in z:
            assert grad_err(outs["grad"], z["ref_grad_scaled_compat"]) < RTOL


def test_rescaling_really_triggers_in_deep_fixture():
    _, z = load_golden("synth_gtr_g4_deep_scaled")
    assert (z["ref_root_scaling"] < 0).any()


def test_oracle_partials_match_reference():
    pb, z = load_golden("tiny_gtr_g4")
    out = O.evaluate(pb, partials=True)
    internal = pb.left >= 0
    np.testing.assert_allclose(out["lower"][internal], z["ref_lower"][internal], rtol=1e-12, atol=0)
    keep = np.arange(pb.nnodes) != pb.root
    np.testing.assert_allclose(out["upper"][keep], z["ref_upper"][keep], rtol=1e-12, atol=1e-300)


def test_exact_scaled_gradient_equals_unscaled_gradient():
    """With rescaling on, the exact form must reproduce the unscaled gradient (the reference's
    per-category normalisation does not when C > 1: SURVEY.md §0.4 ii)."""
    pb, z = load_golden("synth_gtr_g4_deep_scaled")
    base = O.evaluate(pb)["grad"]
    pb.scale = True
    exact = O.evaluate(pb)["grad"]
    assert grad_err(exact, base) < 1e-9
    pb.compat_scaled_gradient = True
    compat = O.evaluate(pb)["grad"]
    assert grad_err(compat, base) > 1e-3  # the quirk is real


def test_gradient_matches_finite_differences():
    pb, _ = load_golden("synth_gtr_g4_tipstates")
    g = O.evaluate(pb)["grad"]
    for n in (0, 3, 13, 20):
        if n == pb.root or n == pb.right[pb.root]:
            continue
        h = 1e-6
        bl = pb.bl.copy()
        pb.bl = bl.copy(); pb.bl[n] += h
        up = O.evaluate(pb, gradient=False)["lnl"]
        pb.bl = bl.copy(); pb.bl[n] -= h
        dn = O.evaluate(pb, gradient=False)["lnl"]
        pb.bl = bl
        assert abs((up - dn) / (2 * h) - g[n]) < 1e-4 * max(1.0, abs(g[n]))
