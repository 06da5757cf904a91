"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref) in this container.

TEST INFRASTRUCTURE ONLY.  Run where /root/reference exists:

    make -C oracle ref && python oracle/make_golden.py

Every fixture holds the plain-array inputs of the hot path exactly as the reference sees them
(node ids, pattern order, weights, eigen system or closed-form matrices, category rates) and the
reference's outputs for them (lnL, per-pattern lnL, branch gradients in the reference's variants).
The known-answer values of /root/reference/tests/test_tree_likelihood.c are stored verbatim in the
C1 fixture as a reference-independent pin.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from physher_b200 import synthetic as syn  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DATA = "/root/reference/tests/data"


def save_case(name: str, ref: O.Reference, extra: dict | None = None, scaled: bool = True, partials: bool = False):
    """Evaluate the reference in every variant the parity tests need and store inputs + outputs."""
    pb = ref.problem()
    d = dict(
        left=pb.left, right=pb.right, parent=pb.parent, root=np.int32(pb.root), nstate=np.int32(pb.nstate),
        tip_states=pb.tip_states, weights=pb.weights, freqs=pb.freqs, rates=pb.rates, props=pb.props, bl=pb.bl,
        use_tip_states=np.int32(pb.use_tip_states), unrooted=np.int32(pb.unrooted),
        time_elapsed=pb.meta["time_elapsed"],
    )
    if not pb.use_tip_states:
        d["tip_partials"] = pb.tip_partials.astype(np.uint8)  # 0/1 vectors (datatype.c get_partials)
        assert np.array_equal(d["tip_partials"].astype(np.float64), pb.tip_partials)
    if pb.evec is not None:
        d.update(evec=pb.evec, eval=pb.eval, ivec=pb.ivec)
    else:
        d.update(P_override=pb.P_override, dP_override=pb.dP_override)
    Pm, dPm = ref.matrices()
    d["ref_matrices"], d["ref_dmatrices"] = Pm, dPm
    d["ref_lnl"] = np.float64(ref.logP())
    d["ref_pattern_lnl"] = ref.pattern_lnl()
    assert not ref.rescaling(), "reference switched rescaling on by itself"
    if not ref.time_mode:
        # exact form: include_root_freqs = false (SURVEY 8c caveat 3); default form: the reference's own default
        d["ref_grad_exact"] = ref.gradient(O.FLAG_TREE_MODEL, include_root_freqs=0)
        d["ref_grad_default"] = ref.gradient(O.FLAG_TREE_MODEL, include_root_freqs=-1)
    if partials:
        low = np.zeros((ref.N, ref.C, ref.P, ref.S))
        up = np.zeros_like(low)
        ref.gradient(O.FLAG_TREE_MODEL, include_root_freqs=0)
        for n in range(ref.N):
            p = ref.partials(n)
            if p is not None:
                low[n] = p
            if n != ref.root:
                up[n] = ref.partials(ref.N + n)
        d["ref_lower"], d["ref_upper"] = low, up
    if scaled:
        ref.use_rescaling(True)
        d["ref_lnl_scaled"] = np.float64(ref.logP())
        sf = ref.scaling_factors(ref.root)
        d["ref_root_scaling"] = sf
        if not ref.time_mode:
            d["ref_grad_scaled_compat"] = ref.gradient(O.FLAG_TREE_MODEL, include_root_freqs=0)
        ref.use_rescaling(False)
    if extra:
        d.update(extra)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: T={ref.T} P={ref.P} S={ref.S} C={ref.C} lnL={float(d['ref_lnl'])!r} -> {os.path.getsize(path)/1024:.0f} KiB")


def synthetic_nucleotide(name, T, sites, model_spec, categories, seed, tipstates, mu=0.3, unknown=0.02, topo=None, scaled=True,
                         bl_range=(0.01, 0.1), partials=False):
    topo = topo or syn.random_topology(T, seed)
    bl = syn.random_branch_lengths(topo, seed + 1, *bl_range)
    pat = syn.random_patterns(T, sites, 4, mu, seed + 2, unknown_frac=unknown)
    names = [f"t{i}" for i in range(T)]
    seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)))
    spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, model_spec, categories=categories, alpha=0.5,
                                 tipstates=tipstates)
    ref = O.Reference(spec)
    save_case(name, ref, scaled=scaled, partials=partials)
    ref.close()


def read_fasta(path):
    seqs, name = {}, None
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs[name] = ""
        elif name:
            seqs[name] += line
    return seqs


def write_c1_dropin_fixture():
    """tests/golden/c1_jc69_time.json: the reference's own fixture tests/data/jc69-time.json with the alignment inlined (the GPU
    box has no /root/reference), and tests/golden/c1_kat.json: the known answers of tests/test_tree_likelihood.c, parsed
    from that file, plus the reference's own gradient vectors for the same requests."""
    import re

    doc = json.load(open(os.path.join(REF_DATA, "jc69-time.json")))
    aln = doc["model"]["sitepattern"]["alignment"]
    aln["sequences"] = read_fasta(os.path.join(REF_DATA, aln.pop("file")))
    doc["model"]["tree"].pop("_file", None)
    json.dump(doc, open(os.path.join(GOLDEN, "c1_jc69_time.json"), "w"), indent=0, separators=(",", ":"))
    src = open("/root/reference/tests/test_tree_likelihood.c").read()

    def scalar(name):
        return float(re.search(r"double\s+" + name + r"\s*=\s*([-0-9.eE+]+)\s*;", src).group(1))

    def vector(name):
        body = re.search(r"double\s+" + name + r"\[\d+\]\s*=\s*\{(.*?)\}", src, re.S).group(1)
        return [float(x) for x in body.replace("\n", " ").split(",") if x.strip()]

    kat = {
        "source": "tests/test_tree_likelihood.c:28-116 (test_treelikelihood_time)",
        "logP": scalar("expected_logP"), "rate_grad": scalar("expected_rate_grad"),
        "ratio_grad": vector("expected_ratio_grad"), "root_height_grad": scalar("expected_root_height_grad"),
        "logP_jacobian": scalar("expected_logP_jacobian"), "ratio_jac_grad": vector("expected_ratios_jac_grad"),
        "root_height_jac_grad": scalar("expected_root_height_jac_grad"),
    }
    # the reference's own TreeLikelihood_gradient output for the requests the drop-in test makes
    ref = O.Reference(doc["model"])
    ref.set_include_jacobian(False)
    kat["ref_gradient_tree_branch"] = [float(x) for x in ref.gradient(O.FLAG_TREE_MODEL | O.FLAG_BRANCH_MODEL)]
    kat["ref_logP"] = ref.logP()
    ref.close()
    json.dump(kat, open(os.path.join(GOLDEN, "c1_kat.json"), "w"), indent=1)
    print("c1 drop-in fixture:", len(aln["sequences"]), "sequences,", len(kat["ratio_grad"]), "ratio gradients")


def write_c1_time_tree_fixture():
    """tests/golden/c1_time_tree.npz: inputs of the time-tree chain of the reference's own fixture (tip dates as heights, ratios,
    root height, clock rate) and the reference's outputs for the fixture's own values and for perturbed samples."""
    doc = json.load(open(os.path.join(GOLDEN, "c1_jc69_time.json")))
    ref = O.Reference(doc["model"])
    pb = ref.problem()
    th, ratios, rates = ref.time_tree()
    rng = np.random.default_rng(20261017)
    B = 6
    samples = np.tile(ratios, (B, 1))
    samples[1:, :-1] = np.clip(ratios[None, :-1] * rng.lognormal(0, 0.15, size=(B - 1, ratios.shape[0] - 1)), 0.02, 0.98)
    samples[1:, -1] = ratios[-1] * rng.lognormal(0, 0.05, size=B - 1)
    rate_samples = rates[0] * np.concatenate([[1.0], rng.lognormal(0, 0.2, size=B - 1)])
    out = dict(tip_heights=th, ratios=samples, rates=rate_samples, ref_lnl=np.zeros(B), ref_lnl_jacobian=np.zeros(B),
               ref_grad=np.zeros((B, ref.T)), ref_grad_jacobian=np.zeros((B, ref.T)))
    for b in range(B):
        ref.set_ratios(samples[b])
        ref.set_clock_rate(rate_samples[b])
        ref.set_include_jacobian(False)
        out["ref_lnl"][b] = ref.logP()
        out["ref_grad"][b] = ref.gradient(O.FLAG_TREE_MODEL | O.FLAG_BRANCH_MODEL)
        ref.set_include_jacobian(True)
        out["ref_lnl_jacobian"][b] = ref.logP()
        out["ref_grad_jacobian"][b] = ref.gradient(O.FLAG_TREE_MODEL | O.FLAG_BRANCH_MODEL)
    ref.close()
    np.savez_compressed(os.path.join(GOLDEN, "c1_time_tree.npz"), **out)
    print("c1 time tree fixture:", B, "samples, lnL", out["ref_lnl"])


def write_sitepattern_fixture():
    """tests/golden/sitepatterns.npz: encoded alignments and the reference's SitePattern (patterns in ITS order, weights) for the
    reference's own fluA alignment and for a synthetic alignment with ambiguity codes that makes its hash table grow five times."""
    out = {}
    doc = json.load(open(os.path.join(GOLDEN, "c1_jc69_time.json")))
    ref = O.Reference(doc["model"])
    pat, w, names = ref.patterns_raw()
    seqs = doc["model"]["sitepattern"]["alignment"]["sequences"]
    out.update(fluA_alignment=O.encode_nucleotides([seqs[n] for n in names]), fluA_patterns=pat, fluA_weights=w)
    ref.close()
    rng = np.random.default_rng(5)
    T, n = 11, 30000
    chars = np.array(list("ACGTRYN-?"))
    base = rng.integers(0, 4, size=(1, n))
    alnc = np.where(rng.random((T, n)) < 0.08, rng.integers(0, 9, size=(T, n)), base)
    seqs = {f"t{i}": "".join(chars[alnc[i]]) for i in range(T)}
    topo = syn.random_topology(T, 1)
    spec = O.treelikelihood_spec(syn.to_newick(topo, syn.random_branch_lengths(topo, 2), list(seqs)), seqs, O.nucleotide_model_spec("jc69"))
    ref = O.Reference(spec)
    pat, w, names = ref.patterns_raw()
    out.update(synth_alignment=O.encode_nucleotides([seqs[n_] for n_ in names]), synth_patterns=pat, synth_weights=w)
    ref.close()
    np.savez_compressed(os.path.join(GOLDEN, "sitepatterns.npz"), **out)
    print("sitepattern fixture:", out["fluA_patterns"].shape, out["synth_patterns"].shape)


def write_dlnl_dq_fixture():
    """tests/golden/dlnl_dq_gtr_g4.npz: a GTR+G4 problem, the reference's dP/d theta matrices (m->dPdp) for its five free rate
    parameters and the reference's own calculate_dlnl_dQ values (include_root_freqs false / true)."""
    T, sites = 12, 400
    topo = syn.random_topology(T, 100)
    bl = syn.random_branch_lengths(topo, 101)
    pat = syn.random_patterns(T, sites, 4, 0.3, 102, unknown_frac=0.02)
    names = [f"t{i}" for i in range(T)]
    seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)))
    gtr = O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1])
    ref = O.Reference(O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, gtr, categories=4, alpha=0.5, tipstates=True))
    pb = ref.problem()
    K = 5
    d = dict(left=pb.left, right=pb.right, parent=pb.parent, root=np.int32(pb.root), nstate=np.int32(4), tip_states=pb.tip_states,
             weights=pb.weights, freqs=pb.freqs, rates=pb.rates, props=pb.props, bl=pb.bl, evec=pb.evec, eval=pb.eval, ivec=pb.ivec,
             use_tip_states=np.int32(1), unrooted=np.int32(1), time_elapsed=pb.meta["time_elapsed"],
             dPdp=np.stack([ref.dPdp(k) for k in range(K)]),
             ref_dlnl_dq=np.array([ref.dlnl_dQ(k, 0) for k in range(K)]),
             ref_dlnl_dq_root_freqs=np.array([ref.dlnl_dQ(k, 1) for k in range(K)]),
             ref_lnl=np.float64(ref.logP()))
    ref.close()
    np.savez_compressed(os.path.join(GOLDEN, "dlnl_dq_gtr_g4.npz"), **d)
    print("dlnl_dQ fixture:", d["dPdp"].shape, d["ref_dlnl_dq"])


def write_branch_fixture():
    """tests/golden/branch_gtr_g4.npz: a GTR+G4 problem and the reference's own single-branch lnL, d lnL/dt, d2 lnL/dt2
    (_calculate_uppper, calculate_dldt_uppper, d2lnldt2_uppper) for a tip, two internal nodes and the root's left child at
    0.5x, 1x and 2.5x their branch length; the same for an HKY problem with one rate category and tip partials."""
    out = {}
    for tag, model, cats, tipstates, seed in (("g4", O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1]), 4, True, 300),
                                              ("c1", O.nucleotide_model_spec("hky", [0.3, 0.2, 0.2, 0.3], kappa=3.0), 1, False, 310)):
        T, sites = 10, 300
        topo = syn.random_topology(T, seed)
        bl = syn.random_branch_lengths(topo, seed + 1)
        pat = syn.random_patterns(T, sites, 4, 0.3, seed + 2, unknown_frac=0.03)
        names = [f"t{i}" for i in range(T)]
        seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.NUCLEOTIDES)))
        ref = O.Reference(O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, model, categories=cats, alpha=0.5, tipstates=tipstates))
        pb = ref.problem()
        root = int(pb.root)
        nodes = [0, 3, T, T + 3, int(pb.left[root])]
        nodes = [n for n in dict.fromkeys(nodes) if n != root and n != int(pb.right[root])]
        factors = np.array([0.5, 1.0, 2.5])
        vals = np.array([[ref.branch_derivatives(n, pb.bl[n] * f) for f in factors] for n in nodes])
        d = dict(left=pb.left, right=pb.right, parent=pb.parent, root=np.int32(root), nstate=np.int32(4), weights=pb.weights, freqs=pb.freqs,
                 rates=pb.rates, props=pb.props, bl=pb.bl, evec=pb.evec, eval=pb.eval, ivec=pb.ivec, use_tip_states=np.int32(tipstates),
                 nodes=np.array(nodes, dtype=np.int32), factors=factors, ref_branch=vals, ref_lnl=np.float64(ref.logP()))
        if tipstates:
            d["tip_states"] = pb.tip_states
        else:
            d["tip_partials"] = pb.tip_partials
        out.update({f"{tag}_{k}": v for k, v in d.items()})
        print("branch fixture", tag, "nodes", nodes, "lnL", d["ref_lnl"], "first", vals[0, 1])
        ref.close()
    np.savez_compressed(os.path.join(GOLDEN, "branch_derivatives.npz"), **out)


def main():
    if "--only-branch" in sys.argv:
        write_branch_fixture()
        return
    if "--only-dq" in sys.argv:
        write_dlnl_dq_fixture()
        return
    if "--only-patterns" in sys.argv:
        write_sitepattern_fixture()
        return
    if "--only-dropin" in sys.argv:
        write_c1_dropin_fixture()
        write_c1_time_tree_fixture()
        return
    write_c1_dropin_fixture()
    write_c1_time_tree_fixture()
    write_sitepattern_fixture()
    write_dlnl_dq_fixture()
    write_branch_fixture()
    cwd = os.getcwd()
    os.chdir(REF_DATA)  # fixtures reference fluA.fa / tiny.fa by relative path

    # --- C1: JC69 strict clock on fluA, the reference's own known-answer test -------------------
    kat = dict(
        kat_lnl=np.float64(-4777.616349713985),  # tests/test_tree_likelihood.c:28
        kat_clock_grad=np.float64(328017.6732813406),  # :39
        kat_lnl_jacobian=np.float64(-4786.867701371271),  # :90
    )
    for tipstates in (True, False):
        spec = json.load(open("jc69-time.json"))["model"]
        spec["tipstates"] = tipstates
        ref = O.Reference(spec)
        ref.set_include_jacobian(False)
        clock = ref.gradient(O.FLAG_BRANCH_MODEL)
        extra = dict(kat, ref_clock_grad=np.float64(clock[0]))
        save_case("c1_jc69_fluA_" + ("tipstates" if tipstates else "tippartials"), ref, extra=extra, scaled=True)
        ref.close()

    # --- tiny.fa, NJ start tree, GTR+G4 non-uniform pi -----------------------
    spec = json.load(open("jc69.json"))["model"]
    spec["sitemodel"]["substitutionmodel"] = O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1])
    spec["sitemodel"]["distribution"] = {"distribution": "gamma", "categories": 4,
                                         "parameters": {"alpha": {"id": "alpha", "type": "parameter", "value": 0.5, "lower": 0}}}
    spec["tipstates"] = False
    ref = O.Reference(spec)
    save_case("tiny_gtr_g4", ref, partials=True)
    ref.close()
    os.chdir(cwd)

    gtr = O.nucleotide_model_spec("gtr", [0.1, 0.2, 0.3, 0.4], [0.05, 0.3, 0.1, 0.15, 0.3, 0.1])
    hky = O.nucleotide_model_spec("hky", [0.1, 0.2, 0.3, 0.4], kappa=3.0)
    jc = O.nucleotide_model_spec("jc69")
    synthetic_nucleotide("synth_gtr_g4_tippartials", 12, 400, gtr, 4, 100, tipstates=False)
    synthetic_nucleotide("synth_gtr_g4_tipstates", 12, 400, gtr, 4, 100, tipstates=True)
    synthetic_nucleotide("synth_hky_g4_tipstates", 17, 300, hky, 4, 200, tipstates=True)
    synthetic_nucleotide("synth_jc69_c1_tipstates", 9, 250, jc, 1, 300, tipstates=True)
    synthetic_nucleotide("synth_gtr_c1_tippartials", 9, 250, gtr, 1, 310, tipstates=False)
    synthetic_nucleotide("synth_gtr_g4_caterpillar", 16, 200, gtr, 4, 400, tipstates=True, topo=syn.caterpillar_topology(16))
    synthetic_nucleotide("synth_gtr_g4_balanced", 16, 200, gtr, 4, 500, tipstates=True, topo=syn.balanced_topology(16))
    # deep tree with long branches: per-pattern likelihood underflows 1e-40 at inner nodes => rescaling really triggers
    synthetic_nucleotide("synth_jc69_c1_deep_scaled", 120, 60, jc, 1, 600, tipstates=True, mu=0.75, unknown=0.0,
                         topo=syn.caterpillar_topology(120), bl_range=(0.5, 1.5))
    synthetic_nucleotide("synth_gtr_g4_deep_scaled", 120, 60, gtr, 4, 700, tipstates=False, mu=0.75, unknown=0.0,
                         topo=syn.caterpillar_topology(120), bl_range=(0.5, 1.5))

    # --- 20-state LG+G4 ---------------------------------------------------------------------------
    T, sites = 8, 160
    topo = syn.random_topology(T, 800)
    bl = syn.random_branch_lengths(topo, 801)
    pat = syn.random_patterns(T, sites, 20, 0.35, 802, unknown_frac=0.02)
    names = [f"t{i}" for i in range(T)]
    seqs = dict(zip(names, syn.sequences_from_patterns(pat, syn.AMINO_ACIDS)))
    # explicit frequencies: new_LG_with_parameters(NULL) builds a 0-dimensional simplex (lg.c:45-49)
    aa_freqs = np.random.default_rng(803).dirichlet(np.full(20, 10.0))
    lg = {"id": "sm", "type": "substitutionmodel", "model": "lg", "datatype": "aa",
          "frequencies": {"id": "freqs", "type": "Simplex", "values": [float(x) for x in aa_freqs]}}
    for tipstates in (False, True):
        spec = O.treelikelihood_spec(syn.to_newick(topo, bl, names), seqs, lg, categories=4, alpha=0.5, tipstates=tipstates, datatype="aa")
        ref = O.Reference(spec)
        save_case("synth_lg_g4_" + ("tipstates" if tipstates else "tippartials"), ref)
        ref.close()

    # --- 61-state GY94 (C API + generic kernels, see ref_harness.c) -------------------------------
    T, sites = 6, 110
    topo = syn.random_topology(T, 900)
    bl = syn.random_branch_lengths(topo, 901)
    pat = syn.random_patterns(T, sites, 61, 0.3, 902)
    bases = "TCAG"
    codons = [a + b + c for a in bases for b in bases for c in bases]
    aa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
    sense = [c for c, a in zip(codons, aa) if a != "*"]
    names = [f"t{i}" for i in range(T)]
    seqs = {n: "".join(sense[s] for s in row) for n, row in zip(names, pat)}
    ref = O.Reference(codon=dict(newick=syn.to_newick(topo, bl, names), sequences=seqs, kappa=2.5, omega=0.3))
    save_case("synth_gy94_tippartials", ref, scaled=False)
    ref.close()


if __name__ == "__main__":
    main()
