"""Drop-in boundary: the UNMODIFIED reference runs its own JSON fixture with the tree likelihood
re-pointed at libphysher_b200.so through integration/physher_glue.c, and must reproduce the known answers of its own
test (tests/test_tree_likelihood.c:28-116: lnL, clock-rate gradient, 67 ratio gradients + root height, with and
without the Jacobian) to the test's own tolerance (1e-8 absolute) and the north-star 1e-10 relative.

Needs oracle/_ref (built where /root/reference exists; travels to the GPU box) and a GPU.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import GOLDEN, RTOL, grad_err, rel_err

pytestmark = pytest.mark.gpu

GLUE_SO = os.path.join(os.path.dirname(O.REF_SO), "libphysher_glue.so")
_probe = []


def phb_launches():
    """kernel launches issued by libphysher_b200 so far in this process, through a throw-away object's counter is not possible -- the
    wrapped phycpp owns its objects -- so the library-wide counter of device evaluations kept by the glue is used"""
    G = C.CDLL(GLUE_SO)
    G.phb_physher_total_evaluations.restype = C.c_longlong
    return int(G.phb_physher_total_evaluations())


@pytest.fixture(scope="module")
def libs():
    if not (O.reference_available() and os.path.exists(GLUE_SO)):
        pytest.skip("oracle/_ref (compiled reference + glue) not present")
    L = O._reflib()
    L.refh_model_handle.argtypes = [C.c_void_p]
    L.refh_model_handle.restype = C.c_void_p
    L.refh_initialize_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.refh_initialize_gradient.restype = C.c_size_t
    L.refh_mark_dirty.argtypes = [C.c_void_p]
    L.refh_kat_dlogP.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    L.refh_kat_dlogP.restype = C.c_int
    G = C.CDLL(GLUE_SO)
    G.phb_physher_attach.argtypes = [C.c_void_p, C.c_int]
    G.phb_physher_attach.restype = C.c_int
    G.phb_physher_detach.argtypes = [C.c_void_p]
    G.phb_physher_gradient.argtypes = [C.c_void_p]
    G.phb_physher_gradient.restype = C.POINTER(C.c_double)
    G.phb_physher_evaluations.argtypes = [C.c_void_p]
    G.phb_physher_evaluations.restype = C.c_longlong
    G.phb_physher_update_uppers.argtypes = [C.c_void_p]
    L.refh_upper_walk.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    L.refh_upper_walk.restype = C.c_int
    L.refh_d2logP_branch.argtypes = [C.c_void_p, C.c_int]
    L.refh_d2logP_branch.restype = C.c_double
    return L, G


def _spec(name="c1_jc69_time.json"):
    return json.load(open(os.path.join(GOLDEN, name)))["model"]


def test_reference_kat_through_the_glue(libs):
    L, G = libs
    kat = json.load(open(os.path.join(GOLDEN, "c1_kat.json")))
    ref = O.Reference(_spec())
    model = L.refh_model_handle(ref.h)
    assert G.phb_physher_attach(model, 0) == 0
    ref.set_include_jacobian(False)
    # logP through the Model vtable -> tlk->calculate -> phb_tlk_calculate
    lnl = ref.logP()
    assert G.phb_physher_evaluations(model) == 1, "the likelihood did not run through libphysher_b200"
    assert abs(lnl - kat["logP"]) < 1e-8 and rel_err(lnl, kat["logP"]) < RTOL
    # the request sequence of the reference's own test through model->prepare_gradient / model->dlogP
    out = np.zeros(80)
    n = L.refh_kat_dlogP(ref.h, out.ctypes.data_as(C.POINTER(C.c_double)), 80)
    assert n == 1 + 68
    want = np.array([kat["rate_grad"]] + kat["ratio_grad"] + [kat["root_height_grad"]])
    assert np.abs(out[:n] - want).max() < 1e-8 or grad_err(out[:n], want) < RTOL
    assert grad_err(out[:n], want) < RTOL
    assert G.phb_physher_evaluations(model) == 2, "one fused device evaluation per gradient request"
    # with the Jacobian of the ratio transform (host chain, the reference's own code on top of device gradients)
    ref.set_include_jacobian(True)
    lnl = ref.logP()
    assert abs(lnl - kat["logP_jacobian"]) < 1e-8
    n = L.refh_kat_dlogP(ref.h, out.ctypes.data_as(C.POINTER(C.c_double)), 80)
    want = np.array([kat["rate_grad"]] + kat["ratio_jac_grad"] + [kat["root_height_jac_grad"]])
    assert grad_err(out[:n], want) < RTOL
    G.phb_physher_detach(model)
    ref.close()


def test_gradient_vector_contract_and_caching(libs):
    """TreeLikelihood_gradient's contract (treelikelihood.c:320-340): tlk-owned buffer, order (ratios | clock), cached."""
    L, G = libs
    kat = json.load(open(os.path.join(GOLDEN, "c1_kat.json")))
    ref = O.Reference(_spec())
    model = L.refh_model_handle(ref.h)
    ref.set_include_jacobian(False)
    G.phb_physher_attach(model, 0)
    n = L.refh_initialize_gradient(ref.h, O.FLAG_TREE_MODEL | O.FLAG_BRANCH_MODEL, -1)
    assert n == 69
    L.refh_mark_dirty(ref.h)
    p1 = G.phb_physher_gradient(model)
    g = np.ctypeslib.as_array(p1, shape=(n,)).copy()
    assert grad_err(g, np.array(kat["ref_gradient_tree_branch"])) < RTOL
    evals = G.phb_physher_evaluations(model)
    p2 = G.phb_physher_gradient(model)  # nothing changed: cached, same buffer, no device work
    assert C.addressof(p1.contents) == C.addressof(p2.contents) and G.phb_physher_evaluations(model) == evals
    assert abs(ref.L.refh_logP(ref.h) - kat["logP"]) < 1e-8  # marks dirty -> recomputed
    assert G.phb_physher_evaluations(model) == evals + 1
    # detach: the reference's own CPU path is back and agrees
    G.phb_physher_detach(model)
    assert abs(ref.logP() - kat["logP"]) < 1e-8
    ref.close()


@pytest.mark.parametrize("tipstates", [True, False])
def test_unrooted_gtr_gamma_through_the_glue(libs, tipstates):
    """Branch-length tree, GTR+G4 with non-uniform pi: eigen system path, unrooted convention, both gradient variants."""
    from physher_b200 import synthetic as syn

    L, G = libs
    T, sites = 14, 500
    topo = syn.random_topology(T, 41)
    bl = syn.random_branch_lengths(topo, 42)
    pat = syn.random_patterns(T, sites, 4, 0.3, 43, unknown_frac=0.02)
    names = [f"t{i}" for i in range(T)]
    seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)))
    gtr = O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1])
    spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, gtr, categories=4, alpha=0.5, tipstates=tipstates)
    ref = O.Reference(spec)
    lnl_cpu = ref.logP()
    g_cpu_exact = ref.gradient(O.FLAG_TREE_MODEL, include_root_freqs=0)
    g_cpu_default = ref.gradient(O.FLAG_TREE_MODEL, include_root_freqs=-1)
    g_cpu_site = ref.gradient(O.FLAG_TREE_MODEL | O.FLAG_SITE_MODEL, include_root_freqs=0)  # [branches | d lnL / d gamma shape]
    assert g_cpu_site.shape[0] == ref.N + 1
    model = L.refh_model_handle(ref.h)
    assert G.phb_physher_attach(model, 0) == 0
    n = L.refh_initialize_gradient(ref.h, O.FLAG_TREE_MODEL | O.FLAG_SITE_MODEL, 0)
    L.refh_mark_dirty(ref.h)
    g = np.ctypeslib.as_array(G.phb_physher_gradient(model), shape=(n,)).copy()
    assert grad_err(g[:-1], g_cpu_site[:-1]) < RTOL
    assert rel_err(g[-1], g_cpu_site[-1]) < 1e-9, "site-model (shape) gradient: device cat_branch_gradient + the reference's own chain"
    G.phb_physher_detach(model)
    assert G.phb_physher_attach(model, 0) == 0
    assert rel_err(ref.logP(), lnl_cpu) < RTOL
    for irf, want in ((0, g_cpu_exact), (-1, g_cpu_default)):
        n = L.refh_initialize_gradient(ref.h, O.FLAG_TREE_MODEL, irf)
        L.refh_mark_dirty(ref.h)
        g = np.ctypeslib.as_array(G.phb_physher_gradient(model), shape=(n,)).copy()
        assert grad_err(g, want) < RTOL
        assert g[ref.root] == 0.0 and g[ref.root_right] == 0.0
    G.phb_physher_detach(model)
    ref.close()


def test_single_branch_fast_path_through_the_glue(libs):
    """tlk->use_upper: the access pattern of serial_brent_optimize_tree (optimizer.c:111-152) driven through tlk->calculate, once on
    the reference's CPU path and once with the likelihood on the device; model->d2logP of a branch length the same way."""
    from physher_b200 import synthetic as syn

    L, G = libs
    T, sites = 11, 400
    topo = syn.random_topology(T, 61)
    bl = syn.random_branch_lengths(topo, 62)
    pat = syn.random_patterns(T, sites, 4, 0.3, 63, unknown_frac=0.02)
    names = [f"t{i}" for i in range(T)]
    seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)))
    gtr = O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1])
    spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, gtr, categories=4, alpha=0.5, tipstates=True)
    factors = np.array([0.5, 2.0, 1.3])
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

    def walk(ref, update_uppers):
        out, ids = np.zeros(3 * 2 * T), np.zeros(3 * 2 * T, dtype=np.int32)
        n = L.refh_upper_walk(ref.h, update_uppers, factors.ctypes.data_as(dp), 3, out.ctypes.data_as(dp), ids.ctypes.data_as(ip), out.size)
        return out[:n], ids[:n]

    cpu = O.Reference(spec)
    want, want_ids = walk(cpu, None)
    d2_cpu = [L.refh_d2logP_branch(cpu.h, n) for n in (0, 5, T + 2)]
    lnl_cpu_end = cpu.logP()
    cpu.close()

    dev = O.Reference(spec)
    model = L.refh_model_handle(dev.h)
    assert G.phb_physher_attach(model, 0) == 0
    got, got_ids = walk(dev, C.cast(G.phb_physher_update_uppers, C.c_void_p))
    assert got.size == want.size == 3 * (2 * T - 3) and (got_ids == want_ids).all()
    assert np.max(np.abs(got - want) / np.abs(want)) < RTOL
    assert G.phb_physher_evaluations(model) >= got.size, "the single-branch evaluations did not run through libphysher_b200"
    d2_dev = [L.refh_d2logP_branch(dev.h, n) for n in (0, 5, T + 2)]
    assert grad_err(np.array(d2_dev), np.array(d2_cpu)) < RTOL
    assert rel_err(dev.logP(), lnl_cpu_end) < RTOL  # all kept lengths reached the device object
    G.phb_physher_detach(model)
    dev.close()


@pytest.mark.parametrize("name,kw", [("gtr", dict(freqs=[0.1, 0.2, 0.3, 0.4], rates=[0.05, 0.3, 0.1, 0.15, 0.3, 0.1])),
                                     ("hky", dict(freqs=[0.3, 0.2, 0.2, 0.3], kappa=3.0))])
def test_substitution_model_gradient_through_the_glue(libs, name, kw):
    """TREELIKELIHOOD_FLAG_SUBSTITUTION_MODEL[_RATES|_FREQUENCIES] (what torchtree-physher requests, physher.hpp:27-34): the reference's
    calculate_dlnl_dQ values (rates, then frequencies with their root term) from the device sweep over the reference's own dPdp matrices."""
    from physher_b200 import synthetic as syn

    L, G = libs
    T, sites = 10, 300
    topo = syn.random_topology(T, 71)
    bl = syn.random_branch_lengths(topo, 72)
    pat = syn.random_patterns(T, sites, 4, 0.3, 73, unknown_frac=0.02)
    names = [f"t{i}" for i in range(T)]
    seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)))
    spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, O.nucleotide_model_spec(name, **kw), categories=4, alpha=0.5, tipstates=True)
    SUBST, RATES, FREQS = 1 << 2, 1 << 3, 1 << 4
    cpu = O.Reference(spec)
    want = {f: cpu.gradient(O.FLAG_TREE_MODEL | f, include_root_freqs=-1) for f in (SUBST, RATES, FREQS)}
    cpu.close()
    dev = O.Reference(spec)
    model = L.refh_model_handle(dev.h)
    assert G.phb_physher_attach(model, 0) == 0
    for f in (SUBST, RATES, FREQS):
        n = L.refh_initialize_gradient(dev.h, O.FLAG_TREE_MODEL | f, -1)
        assert n == want[f].size
        L.refh_mark_dirty(dev.h)
        g = np.ctypeslib.as_array(G.phb_physher_gradient(model), shape=(n,)).copy()
        assert grad_err(g[:dev.N], want[f][:dev.N]) < RTOL
        assert grad_err(g[dev.N:], want[f][dev.N:]) < RTOL, (f, g[dev.N:], want[f][dev.N:])
    G.phb_physher_detach(model)
    dev.close()


def test_model_clone_gets_its_own_device_object(libs):
    """Model.clone of an attached tree likelihood (_treeLikelihood_model_clone, treelikelihood.c:715-790; what gradascent.c:166-170 does
    per worker): the reference copies the function pointers of the source object, so the glue's clone wrapper must give the
    clone a device object of its own (phb_tlk_clone).  Both reproduce the known answer, move independently, and free cleanly."""
    L, G = libs
    L.refh_clone.argtypes = [C.c_void_p]
    L.refh_clone.restype = C.c_void_p
    L.refh_free.argtypes = [C.c_void_p]
    L.refh_set_clock_rate.argtypes = [C.c_void_p, C.c_double]
    L.refh_logP.argtypes = [C.c_void_p]
    L.refh_logP.restype = C.c_double
    kat = json.load(open(os.path.join(GOLDEN, "c1_kat.json")))
    ref = O.Reference(_spec())
    ref.set_include_jacobian(False)
    model = L.refh_model_handle(ref.h)
    assert G.phb_physher_attach(model, 0) == 0
    assert rel_err(ref.logP(), kat["logP"]) < RTOL
    twin = L.refh_clone(ref.h)
    twin_model = L.refh_model_handle(twin)
    assert G.phb_physher_evaluations(twin_model) == 0, "the clone has its own backend record"
    # (the reference's clone of this time tree does not evaluate to the source's value -- its cloned tree carries different
    # heights -- so the clone is checked against the reference's own CPU path ON THE CLONE, below)
    base = L.refh_logP(twin)
    assert np.isfinite(base) and G.phb_physher_evaluations(twin_model) == 1
    # the clone moves (a different clock rate); the source still gives the known answer
    L.refh_set_clock_rate(twin, 0.002)
    moved = L.refh_logP(twin)
    assert np.isfinite(moved) and moved != base
    evals = G.phb_physher_evaluations(model)
    assert rel_err(ref.logP(), kat["logP"]) < RTOL and G.phb_physher_evaluations(model) == evals + 1
    # same values as the reference's own CPU path on the clone
    G.phb_physher_detach(twin_model)
    assert rel_err(L.refh_logP(twin), moved) < RTOL
    L.refh_free(twin)
    G.phb_physher_detach(model)
    ref.close()
    ref2 = O.Reference(_spec())  # the clone's starting value, from an untouched CPU clone
    ref2.set_include_jacobian(False)
    twin2 = L.refh_clone(ref2.h)
    assert rel_err(base, L.refh_logP(twin2)) < RTOL
    L.refh_free(twin2)
    ref2.close()


# ---------------------------------------------------------------------------------------------------------------------------------
# round 2: the rest of SURVEY.md 8b -- JSON factory with "backend", C1 on the fused walk, the struct slots callers drive themselves,
# Model.store / restore, invariant-site / mu gradients, finite-difference substitution routes, use_upper on a time tree
# ---------------------------------------------------------------------------------------------------------------------------------

def _bind_round2(L, G):
    import physher_b200 as phb

    lib = phb.load_library()
    L.refh_create_with.argtypes = [C.c_char_p, C.c_void_p]
    L.refh_create_with.restype = C.c_void_p
    L.refh_update_uppers.argtypes = [C.c_void_p]
    L.refh_update_uppers.restype = C.c_double
    L.refh_pinned_state_pattern_lnl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.refh_slot_update_upper.argtypes = [C.c_void_p, C.c_int]
    L.refh_set_distance.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.refh_store.argtypes = [C.c_void_p]
    L.refh_restore.argtypes = [C.c_void_p]
    L.refh_plain_logP.argtypes = [C.c_void_p]
    L.refh_plain_logP.restype = C.c_double
    G.phb_physher_handle.argtypes = [C.c_void_p]
    G.phb_physher_handle.restype = C.c_void_p
    G.phb_physher_sync_partials_to_host.argtypes = [C.c_void_p]
    return lib


def _gtr_spec(T=11, sites=400, seed=161, tipstates=True, categories=4, model=None, sitemodel_extra=None):
    from physher_b200 import synthetic as syn

    topo = syn.random_topology(T, seed)
    bl = syn.random_branch_lengths(topo, seed + 1)
    pat = syn.random_patterns(T, sites, 4, 0.3, seed + 2, unknown_frac=0.02)
    names = [f"t{i}" for i in range(T)]
    seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)))
    model = model or O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1])
    spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, model, categories=categories, alpha=0.5, tipstates=tipstates)
    if sitemodel_extra:
        spec["sitemodel"].update(sitemodel_extra)
    return spec, topo


def test_json_backend_key_and_c1_on_the_fused_walk(libs):
    """`"backend": "b200"` on the reference's own "treelikelihood" JSON object (treelikelihood.c:819-943): the glue's factory strips
    the key, runs the reference's constructor and attaches.  C1 (JC69, closed-form p_t in the reference) reaches the fused 4-state
    walk through the closed-form eigen system, and reproduces the known answers of tests/test_tree_likelihood.c:28-84."""
    L, G = libs
    lib = _bind_round2(L, G)
    kat = json.load(open(os.path.join(GOLDEN, "c1_kat.json")))
    spec = dict(_spec(), backend="b200", device=0)
    factory = C.cast(G.phb_physher_new_TreeLikelihoodModel_from_json, C.c_void_p)
    h = L.refh_create_with(json.dumps({"model": spec}).encode(), factory)
    assert h
    model = L.refh_model_handle(h)
    assert G.phb_physher_evaluations(model) == 0, "the factory attached the device backend"
    L.refh_set_include_jacobian(h, 0)
    lnl = L.refh_logP(h)
    assert G.phb_physher_evaluations(model) == 1
    assert abs(lnl - kat["logP"]) < 1e-8 and rel_err(lnl, kat["logP"]) < RTOL
    assert lib.phb_tlk_last_kernels(G.phb_physher_handle(model)) == 2, "C1 must run on k_nuc4_walk, not on the node-at-a-time kernels"
    out = np.zeros(80)
    n = L.refh_kat_dlogP(h, out.ctypes.data_as(C.POINTER(C.c_double)), 80)
    want = np.array([kat["rate_grad"]] + kat["ratio_grad"] + [kat["root_height_grad"]])
    assert n == 69 and grad_err(out[:n], want) < RTOL
    assert lib.phb_tlk_last_kernels(G.phb_physher_handle(model)) == 2
    # without the key the same factory leaves the reference's CPU path in place
    h2 = L.refh_create_with(json.dumps({"model": _spec()}).encode(), factory)
    assert G.phb_physher_evaluations(L.refh_model_handle(h2)) == -1
    L.refh_set_include_jacobian(h2, 0)
    assert abs(L.refh_logP(h2) - kat["logP"]) < 1e-8
    G.phb_physher_detach(model)


@pytest.mark.parametrize("tipstates", [True, False])
def test_direct_struct_slots_run_on_the_device(libs, tipstates):
    """tlk->update_partials / integrate_partials / node_log_likelihoods / calculate_per_cat_partials (treelikelihood.h:90-94,111) as
    the reference's own non-virtual code drives them: SingleTreeLikelihood_update_uppers (_calculate_simple + update_upper_partials),
    the masked-state evaluation of asr_marginal (asr.c:60-69) and one tripod-style upper update of SPR (spropt.c:1578-1608).
    Same calls on the CPU reference and on an attached model; tlk->partials on the host must mirror what the device computed."""
    L, G = libs
    _bind_round2(L, G)
    spec, topo = _gtr_spec(tipstates=tipstates)
    dp = C.POINTER(C.c_double)

    def run(ref):
        out = {"lnl": L.refh_update_uppers(ref.h)}
        N, T = ref.N, ref.T
        out["lower"] = {n: ref.partials(n) for n in range(T, N)}
        out["upper"] = {n: ref.partials(N + n) for n in range(N) if n != ref.root}
        pinned = np.zeros((2, 4, ref.P))
        for i, node in enumerate((T + 1, ref.root)):
            for s in range(4):
                L.refh_pinned_state_pattern_lnl(ref.h, node, s, pinned[i, s].ctypes.data_as(dp))
        out["pinned"] = pinned
        # SPR tripod step: the sibling's length changes, the caller refreshes one upper partial through the slot
        node = 2
        sib = int(topo.right[topo.parent[node]]) if int(topo.left[topo.parent[node]]) == node else int(topo.left[topo.parent[node]])
        L.refh_set_distance(ref.h, sib, 0.123)
        L.refh_slot_update_upper(ref.h, node)
        out["tripod_upper"] = ref.partials(N + node)
        return out

    cpu = O.Reference(spec)
    want = run(cpu)
    cpu.close()
    dev = O.Reference(spec)
    model = L.refh_model_handle(dev.h)
    assert G.phb_physher_attach(model, 0) == 0
    got = run(dev)
    assert G.phb_physher_evaluations(model) > 0
    assert rel_err(got["lnl"], want["lnl"]) < RTOL
    for key in ("lower", "upper"):
        for n in want[key]:
            np.testing.assert_allclose(got[key][n], want[key][n], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(got["pinned"], want["pinned"], rtol=1e-10, atol=0)
    np.testing.assert_allclose(got["tripod_upper"], want["tripod_upper"], rtol=1e-10, atol=1e-300)
    G.phb_physher_detach(model)
    dev.close()


def test_model_store_restore_through_the_glue(libs):
    """Model.store / Model.restore of an attached tree likelihood (_singleTreeLikelihood_store / _restore, treelikelihood.c:126-161):
    an MCMC reject returns to the stored lnL; the device object is not asked to recompute it.  On the JC69 time tree: the reference's
    own tree-model store dereferences the tree transform (tree.c:1017-1027), so only reparameterised trees can be stored at all."""
    L, G = libs
    _bind_round2(L, G)
    kat = json.load(open(os.path.join(GOLDEN, "c1_kat.json")))
    dev = O.Reference(_spec())
    dev.set_include_jacobian(False)
    model = L.refh_model_handle(dev.h)
    assert G.phb_physher_attach(model, 0) == 0
    base = L.refh_plain_logP(dev.h)
    assert rel_err(base, kat["logP"]) < RTOL
    L.refh_store(dev.h)
    dev.set_clock_rate(0.003)  # a proposal
    moved = L.refh_plain_logP(dev.h)
    assert abs(moved - base) > 1e-3
    evals = G.phb_physher_evaluations(model)
    L.refh_restore(dev.h)  # reject
    assert L.refh_plain_logP(dev.h) == pytest.approx(base, rel=1e-13)
    assert G.phb_physher_evaluations(model) == evals, "a rejected proposal must not cost a device evaluation"
    # accept: store after the move; the moved value is what a later reject returns to
    dev.set_clock_rate(0.003)
    assert rel_err(L.refh_plain_logP(dev.h), moved) < 1e-12
    L.refh_store(dev.h)
    dev.set_clock_rate(0.0007)
    assert abs(L.refh_plain_logP(dev.h) - moved) > 1e-3
    L.refh_restore(dev.h)
    assert rel_err(L.refh_plain_logP(dev.h), moved) < 1e-12
    # the reference's own CPU path agrees on the value the chain returned to
    G.phb_physher_detach(model)
    dev.set_clock_rate(0.003)
    assert rel_err(dev.logP(), moved) < RTOL
    dev.close()


@pytest.mark.parametrize("case", ["pinv", "pinv_gamma", "mu"])
def test_site_model_gradients_with_invariant_sites_and_mu(libs, case):
    """TREELIKELIHOOD_FLAG_SITE_MODEL with an invariant-site proportion (gradient_pinv_sitemodel / gradient_pinv_W_sitemodel,
    treelikelihood.c:2943-3001: they read the root partials) and with a mutation-rate multiplier mu (:3245-3249, :3283-3296)."""
    L, G = libs
    _bind_round2(L, G)
    simplex = {"id": "props", "type": "Simplex", "values": [0.2, 0.8]}
    if case == "pinv":  # +I alone: an invariant category and one variable category (sitemodel.c:1182-1197)
        spec, topo = _gtr_spec(seed=361, categories=1)
        spec["sitemodel"]["distribution"] = {"distribution": "discrete", "categories": 1, "proportions": simplex}
    elif case == "pinv_gamma":  # Weibull + I: five categories, the shape and the invariant proportion both carry a gradient
        spec, topo = _gtr_spec(seed=361, categories=4)
        spec["sitemodel"]["distribution"] = {"distribution": "weibull", "categories": 4, "proportions": simplex,
                                             "parameters": {"shape": {"id": "alpha", "type": "parameter", "value": 0.5, "lower": 0}}}
    else:
        spec, topo = _gtr_spec(seed=361, categories=4)
        spec["sitemodel"]["mu"] = {"id": "mu", "type": "parameter", "value": 1.7, "lower": 0}
    cpu = O.Reference(spec)
    flags = O.FLAG_TREE_MODEL | O.FLAG_SITE_MODEL
    want = cpu.gradient(flags, include_root_freqs=0)
    lnl_cpu = cpu.logP()
    cpu.close()
    dev = O.Reference(spec)
    model = L.refh_model_handle(dev.h)
    assert G.phb_physher_attach(model, 0) == 0
    assert rel_err(dev.logP(), lnl_cpu) < RTOL
    n = L.refh_initialize_gradient(dev.h, flags, 0)
    assert n == want.size and n > dev.N, "the request carries site-model entries"
    L.refh_mark_dirty(dev.h)
    g = np.ctypeslib.as_array(G.phb_physher_gradient(model), shape=(n,)).copy()
    assert grad_err(g[:dev.N], want[:dev.N]) < RTOL
    for k in range(dev.N, n):
        assert rel_err(g[k], want[k]) < 1e-9, (case, k, g[k], want[k])
    G.phb_physher_detach(model)
    dev.close()


def test_use_upper_on_a_time_tree_stays_on_the_device(libs):
    """tlk->use_upper on the JC69 time tree (r1: exit(2)): a changed clock rate or node height moves several branch lengths, the glue
    pushes them all and lnL comes from the device's resident partials."""
    L, G = libs
    _bind_round2(L, G)
    cpu = O.Reference(_spec())
    cpu.set_include_jacobian(False)
    cpu.set_clock_rate(0.003)
    want = cpu.logP()
    cpu.close()
    dev = O.Reference(_spec())
    dev.set_include_jacobian(False)
    model = L.refh_model_handle(dev.h)
    assert G.phb_physher_attach(model, 0) == 0
    G.phb_physher_update_uppers(model)  # switches use_upper on, partials resident
    evals = G.phb_physher_evaluations(model)
    dev.set_clock_rate(0.003)
    got = L.refh_plain_logP(dev.h)
    assert rel_err(got, want) < RTOL and G.phb_physher_evaluations(model) > evals
    G.phb_physher_detach(model)
    dev.close()


# ---------------------------------------------------------------------------------------------------------------------------------
# phycpp: the reference's C++ wrapper (src/phycpp/physher.cpp, what torchtree-physher binds), UNMODIFIED, on the device path
# ---------------------------------------------------------------------------------------------------------------------------------

def _phycpp(name):
    path = os.path.join(os.path.dirname(O.REF_SO), name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not built (make -C oracle phycpp, where /root/reference exists)")
    lib = C.CDLL(path)
    dp, cpp = C.POINTER(C.c_double), C.POINTER(C.c_char_p)
    lib.phycpp_unrooted_gtr.argtypes = [C.c_char_p, C.c_int, cpp, cpp, dp, dp, C.c_double, C.c_int, C.c_int, dp, dp, C.c_int]
    lib.phycpp_unrooted_gtr.restype = C.c_int
    lib.phycpp_time_jc69.argtypes = [C.c_char_p, C.c_int, cpp, dp, cpp, C.c_double, dp, dp, C.c_int]
    lib.phycpp_time_jc69.restype = C.c_int
    return lib


def _strs(items):
    return (C.c_char_p * len(items))(*[s.encode() for s in items])


def _phycpp_unrooted(lib, T=13, sites=350, seed=461, cats=4, tipstates=False):
    from physher_b200 import synthetic as syn

    topo = syn.random_topology(T, seed)
    bl = syn.random_branch_lengths(topo, seed + 1)
    pat = syn.random_patterns(T, sites, 4, 0.3, seed + 2, unknown_frac=0.02)
    names = [f"t{i}" for i in range(T)]
    seqs = syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)
    rates = np.array([0.05, 0.3, 0.1, 0.15, 0.3, 0.1])
    freqs = np.array([0.1, 0.2, 0.3, 0.4])
    lnl, grad = C.c_double(0.0), np.zeros(2 * T + 4)
    dp = C.POINTER(C.c_double)
    n = lib.phycpp_unrooted_gtr(syn.to_newick(topo, bl, names).encode(), T, _strs(names), _strs(seqs), rates.ctypes.data_as(dp),
                                freqs.ctypes.data_as(dp), 0.5, cats, int(tipstates), C.byref(lnl), grad.ctypes.data_as(dp), grad.size)
    return lnl.value, grad[:n].copy()


def test_phycpp_cpu_harness_matches_the_c_reference():
    """The harness itself (CPU link): TreeLikelihoodInterface::LogLikelihood of the C++ wrapper is the C model's logP."""
    cpu = _phycpp("libphycpp_cpu.so")
    lnl, g = _phycpp_unrooted(cpu)
    assert np.isfinite(lnl) and g.size == 2 * 13 - 1 - 2 and np.isfinite(g).all() and np.abs(g).max() > 1.0


test_phycpp_cpu_harness_matches_the_c_reference.pytestmark = []  # runs without a GPU


@pytest.mark.parametrize("cats,tipstates", [(4, False), (1, True)])
def test_phycpp_tree_likelihood_interface_on_the_device(libs, cats, tipstates):
    """TreeLikelihoodInterface (physher.cpp:560-665) of the UNMODIFIED C++ wrapper, linked with --wrap so that its constructor attaches the
    device backend and Gradient() -- which calls TreeLikelihood_gradient directly -- lands in phb_physher_gradient
    (integration/phycpp_wrap.c): LogLikelihood / Gradient against the plainly linked wrapper on the CPU."""
    L, G = libs
    cpu = _phycpp("libphycpp_cpu.so")
    dev = _phycpp("libphycpp_b200.so")
    want_l, want_g = _phycpp_unrooted(cpu, cats=cats, tipstates=tipstates)
    before = phb_launches()
    got_l, got_g = _phycpp_unrooted(dev, cats=cats, tipstates=tipstates)
    assert phb_launches() > before, "the wrapped phycpp did not launch anything on the device"
    assert rel_err(got_l, want_l) < RTOL
    # phycpp requests the reference's default gradient (include_root_freqs = true): the drop-in reproduces that form too
    assert grad_err(got_g, want_g) < RTOL


def test_phycpp_time_tree_kat_on_the_device(libs):
    """C1 through phycpp: ReparameterizedTimeTreeModelInterface + JC69Interface + StrictClockModelInterface + TreeLikelihoodInterface
    on the fluA fixture -- the path torchtree-physher takes for examples/fluA -- against the reference's known answers."""
    L, G = libs
    kat = json.load(open(os.path.join(GOLDEN, "c1_kat.json")))
    fx = json.load(open(os.path.join(GOLDEN, "c1_jc69_time.json")))
    spec = fx["model"]
    seqs = spec["sitepattern"]["alignment"]["sequences"]
    names = list(seqs.keys())
    dates = np.array([float(spec["tree"]["dates"][n]) for n in names])
    dp = C.POINTER(C.c_double)
    out = {}
    for tag, libname in (("cpu", "libphycpp_cpu.so"), ("dev", "libphycpp_b200.so")):
        lib = _phycpp(libname)
        lnl, grad = C.c_double(0.0), np.zeros(200)
        n = lib.phycpp_time_jc69(spec["tree"]["newick"].encode(), len(names), _strs(names), dates.ctypes.data_as(dp), _strs([seqs[k] for k in names]),
                                 float(spec["branchmodel"]["rate"]["value"]), C.byref(lnl), grad.ctypes.data_as(dp), grad.size)
        out[tag] = (lnl.value, grad[:n].copy())
    assert abs(out["cpu"][0] - kat["logP"]) < 1e-6, "the harness rebuilt the fixture"
    assert rel_err(out["dev"][0], out["cpu"][0]) < RTOL and grad_err(out["dev"][1], out["cpu"][1]) < RTOL
