"""The algebraic identities the device kernels are built on, checked on the CPU against the oracle's node-at-a-time values (which are
pinned on the unmodified reference).  They document WHY the fused kernels may compute what they compute:

* message form: M_n = P_n L_n travels between nodes, L_n = M_a o M_b (csrc/phb_nuc4.cu, csrc/phb_dmma.cu);
* dP/dt L = Q (P L) for any rate matrix (4-state walk);
* adjoint form of the branch gradient: sum_i f_i U_n[i] (dP_n L_n)[i] = sum_j L_n[j] (U_n (f o dP_n))[j] (tensor-core kernels);
* transition statistics: G_n[c] = sum_k w_k / L_k (f o U_n)^T L_n, and sum_ij G_n[c][i][j] M[i][j] is the node sweep of calculate_dlnl_dQ
  (treelikelihood.c:2337-2583) for ANY matrix set M, the root entry being the root term of the frequency gradient (:2371-2404).
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_parity import _synthetic_problem
from tests.util import grad_err


def _case(S, C, seed):
    pb = _synthetic_problem(11, 90, S, C, seed=seed, unknown=0.04)
    res = O.evaluate(pb, partials=True, matrices=True)
    site = np.exp(res["pattern_lnl"])
    # the oracle keeps state tips as states; give them their partial vectors (one-hot, all ones for an unknown state)
    L = res["lower"].copy()
    for t in range(pb.ntips):
        s_ = pb.tip_states[t]
        x = np.ones((pb.npatterns, S))
        known = s_ < S
        x[known] = 0.0
        x[known, s_[known]] = 1.0
        L[t] = x[None]
    res["lower"] = L
    return pb, res, site


@pytest.mark.parametrize("S,C", [(4, 4), (20, 2), (61, 1)])
def test_message_form_and_q_times_p(S, C):
    pb, res, _ = _case(S, C, 6100 + S)
    L, Pm, dPm = res["lower"], res["matrices"], res["dmatrices"]
    Q = (pb.evec * pb.eval) @ pb.ivec
    for n in range(pb.ntips, pb.nnodes):
        a, b = int(pb.left[n]), int(pb.right[n])
        Ma = np.einsum("cij,ckj->cki", Pm[a], L[a])
        Mb = np.einsum("cij,ckj->cki", Pm[b], L[b])
        np.testing.assert_allclose(L[n], Ma * Mb, rtol=1e-12, atol=0)  # L_n = M_a o M_b
        for x, M in ((a, Ma), (b, Mb)):
            want = np.einsum("cij,ckj->cki", dPm[x], L[x])
            got = np.einsum("ij,ckj->cki", Q, M)  # dP/dt L = Q (P L), up to the fabs() the reference puts on P only
            scale = np.abs(want).max()
            assert np.abs(got - want).max() <= 1e-12 * scale


@pytest.mark.parametrize("S,C", [(4, 4), (20, 2), (61, 1)])
def test_adjoint_form_of_the_branch_gradient(S, C):
    pb, res, site = _case(S, C, 6200 + S)
    L, U, dPm = res["lower"], res["upper"], res["dmatrices"]
    coef = pb.weights / site
    for n in range(pb.nnodes):
        if n == pb.root or (pb.unrooted and n == pb.right[pb.root]):  # the unrooted convention zeroes that entry (treelikelihood.c:3249-3255)
            continue
        Z = np.einsum("cki,i,cij->ckj", U[n], pb.freqs, dPm[n])  # U_n (f o dP_n)
        g_adj = np.einsum("k,ckj,ckj->c", coef, L[n], Z)
        np.testing.assert_allclose(g_adj, res["cat_grad"][n], rtol=1e-10, atol=1e-10 * np.abs(res["cat_grad"]).max())


@pytest.mark.parametrize("S,C", [(4, 4), (4, 1), (20, 2)])
def test_transition_statistics_contract_to_the_matrix_gradient(S, C):
    pb, res, site = _case(S, C, 6300 + S + C)
    L, U = res["lower"], res["upper"]
    coef = pb.weights / site
    G = np.einsum("k,ncki,i,nckj->ncij", coef, U, pb.freqs, L)  # [N][C][S][S]
    M = np.random.default_rng(6301).normal(size=(4, pb.nnodes, C, S, S))
    keep = np.ones(pb.nnodes, bool)
    keep[pb.root] = False
    if pb.unrooted:
        keep[pb.right[pb.root]] = False  # treelikelihood.c:2408
    got = np.einsum("n,c,ncij,sncij->s", keep.astype(float), pb.props, G, M)
    assert grad_err(got, O.matrix_gradient(pb, M)) < 1e-10
    # the branch gradient is the same contraction with dP/dt (times the category rate): what cat_grad holds
    cat = np.einsum("ncij,ncij->nc", G, res["dmatrices"])
    np.testing.assert_allclose(cat[keep], res["cat_grad"][keep], rtol=1e-10, atol=1e-10 * np.abs(res["cat_grad"]).max())
    # root entry of the statistics = root term of the frequency gradient
    R = np.einsum("c,cki->ki", pb.props, L[pb.root])
    want_root = (coef[:, None] * R).sum(0)
    got_root = np.einsum("c,k,cki->i", pb.props, coef, L[pb.root])
    np.testing.assert_allclose(got_root, want_root, rtol=1e-13)
    assert abs(got_root @ pb.freqs - pb.weights.sum()) < 1e-9 * pb.weights.sum()
