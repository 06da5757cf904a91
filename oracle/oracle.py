"""ctypes wrappers over the CPU oracle (oracle/phb_oracle.c) and the compiled reference harness.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by physher_b200/.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_harness.so")
REFERENCE_ROOT = "/root/reference"


def build(ref: bool = True) -> None:
    """Compile the restatement and, when /root/reference is present, the reference itself."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "phyc")):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])
        if os.path.exists(os.path.join(os.path.dirname(HERE), "physher_b200", "libphysher_b200.so")):
            # reference-side binding of INTEGRATION.md and the reference's C++ wrapper linked onto it (tests/test_glue_dropin.py)
            subprocess.check_call(["make", "-s", "-C", HERE, "glue", "phycpp"])


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_uint8)


class _Problem(C.Structure):
    _fields_ = [
        ("T", C.c_int), ("N", C.c_int), ("S", C.c_int), ("C", C.c_int), ("P", C.c_int), ("root", C.c_int),
        ("left", _ip), ("right", _ip), ("parent", _ip),
        ("tip_mode", C.c_int), ("tip_states", _bp), ("tip_partials", _dp), ("weights", _dp),
        ("evec", _dp), ("eval", _dp), ("ivec", _dp), ("P_override", _dp), ("dP_override", _dp),
        ("freqs", _dp), ("rates", _dp), ("props", _dp), ("bl", _dp),
        ("scale", C.c_int), ("scaling_threshold", C.c_double),
        ("include_root_freqs", C.c_int), ("compat_scaled_gradient", C.c_int), ("unrooted", C.c_int),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("lnl", C.c_double), ("pattern_lnl", _dp), ("grad", _dp), ("cat_grad", _dp),
        ("lower", _dp), ("upper", _dp), ("matrices", _dp), ("dmatrices", _dp), ("scaling", _dp),
    ]


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


@dataclass
class Problem:
    """Plain-array description of one tree-likelihood evaluation (shared by oracle and GPU tests)."""

    left: np.ndarray
    right: np.ndarray
    parent: np.ndarray
    root: int
    nstate: int
    tip_states: np.ndarray | None  # uint8 [T][P]
    weights: np.ndarray  # [P]
    freqs: np.ndarray
    rates: np.ndarray
    props: np.ndarray
    bl: np.ndarray
    evec: np.ndarray | None = None
    eval: np.ndarray | None = None
    ivec: np.ndarray | None = None
    tip_partials: np.ndarray | None = None  # [T][P][S]
    use_tip_states: bool = True
    P_override: np.ndarray | None = None
    dP_override: np.ndarray | None = None
    scale: bool = False
    scaling_threshold: float = 1e-40
    include_root_freqs: bool = False
    compat_scaled_gradient: bool = False
    unrooted: bool = True
    meta: dict = field(default_factory=dict)

    @property
    def ntips(self):
        return (self.left.shape[0] + 1) // 2

    @property
    def nnodes(self):
        return self.left.shape[0]

    @property
    def npatterns(self):
        return self.weights.shape[0]

    @property
    def ncat(self):
        return self.rates.shape[0]


_lib = None


def _oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _lib = C.CDLL(ORACLE_SO)
        _lib.oracle_evaluate.argtypes = [C.POINTER(_Problem), C.POINTER(_Result)]
        _lib.oracle_evaluate.restype = C.c_int
        _lib.oracle_p_t.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
        _lib.oracle_dp_dt.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
    return _lib


def evaluate(pb: Problem, gradient: bool = True, partials: bool = False, matrices: bool = False) -> dict:
    """Run the C restatement. Returns dict(lnl, pattern_lnl, grad, cat_grad[, lower, upper, matrices, dmatrices, scaling])."""
    lib = _oracle()
    N, S, Cc, P = pb.nnodes, pb.nstate, pb.ncat, pb.npatterns
    keep = []

    def c64(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return a

    def c32(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return a

    cp = _Problem()
    cp.T, cp.N, cp.S, cp.C, cp.P, cp.root = pb.ntips, N, S, Cc, P, int(pb.root)
    cp.left, cp.right, cp.parent = _i(c32(pb.left)), _i(c32(pb.right)), _i(c32(pb.parent))
    cp.tip_mode = 0 if pb.use_tip_states else 1
    if pb.tip_states is not None:
        ts = np.ascontiguousarray(pb.tip_states, dtype=np.uint8)
        keep.append(ts)
        cp.tip_states = ts.ctypes.data_as(_bp)
    if pb.tip_partials is not None:
        cp.tip_partials = _d(c64(pb.tip_partials))
    cp.weights = _d(c64(pb.weights))
    if pb.evec is not None:
        cp.evec, cp.eval, cp.ivec = _d(c64(pb.evec)), _d(c64(pb.eval)), _d(c64(pb.ivec))
    if pb.P_override is not None:
        cp.P_override = _d(c64(pb.P_override))
    if pb.dP_override is not None:
        cp.dP_override = _d(c64(pb.dP_override))
    cp.freqs, cp.rates, cp.props, cp.bl = _d(c64(pb.freqs)), _d(c64(pb.rates)), _d(c64(pb.props)), _d(c64(pb.bl))
    cp.scale = int(pb.scale)
    cp.scaling_threshold = pb.scaling_threshold
    cp.include_root_freqs = int(pb.include_root_freqs)
    cp.compat_scaled_gradient = int(pb.compat_scaled_gradient)
    cp.unrooted = int(pb.unrooted)

    out = {"pattern_lnl": np.zeros(P)}
    res = _Result()
    res.pattern_lnl = _d(out["pattern_lnl"])
    if gradient:
        out["grad"] = np.zeros(N)
        out["cat_grad"] = np.zeros((N, Cc))
        res.grad, res.cat_grad = _d(out["grad"]), _d(out["cat_grad"])
    if partials:
        out["lower"] = np.zeros((N, Cc, P, S))
        out["scaling"] = np.zeros((N, P))
        res.lower, res.scaling = _d(out["lower"]), _d(out["scaling"])
        if gradient:
            out["upper"] = np.zeros((N, Cc, P, S))
            res.upper = _d(out["upper"])
    if matrices:
        out["matrices"] = np.zeros((N, Cc, S, S))
        out["dmatrices"] = np.zeros((N, Cc, S, S))
        res.matrices, res.dmatrices = _d(out["matrices"]), _d(out["dmatrices"])
    rc = lib.oracle_evaluate(C.byref(cp), C.byref(res))
    if rc != 0:
        raise MemoryError("oracle_evaluate failed")
    out["lnl"] = res.lnl
    return out


# DataType.encoding of the nucleotide data type: the NUCLEOTIDE_STATES table (datatype.c:74-91), case-insensitive;
# letters without a meaning and '?' -> 16, every other character (gaps included) -> 17
NUCLEOTIDE_CODES = dict(A=0, C=1, G=2, T=3, U=3, R=5, Y=6, M=7, W=8, S=9, K=10, B=11, D=12, H=13, V=14, N=15)


def encode_nucleotides(sequences) -> np.ndarray:
    def code(ch):
        u = ch.upper()
        if u in NUCLEOTIDE_CODES:
            return NUCLEOTIDE_CODES[u]
        return 16 if (u.isalpha() and u.isascii()) or ch == "?" else 17

    return np.array([[code(ch) for ch in s] for s in sequences], dtype=np.uint8)


def compress_patterns(alignment, hashtable_size=100):
    """new_SitePattern2 restated (incl. the reference's pattern order): uint8 [T][nsites] -> (patterns [T][P], weights [P], site_to_pattern)."""
    lib = _oracle()
    a = np.ascontiguousarray(alignment, dtype=np.uint8)
    T, n = a.shape
    pat = np.zeros(T * n, np.uint8)
    w = np.zeros(n)
    smap = np.zeros(n, np.int32)
    lib.oracle_compress_patterns.argtypes = [C.c_int, C.c_size_t, _bp, C.c_uint, _bp, _dp, _ip]
    lib.oracle_compress_patterns.restype = C.c_long
    P = lib.oracle_compress_patterns(T, n, a.ctypes.data_as(_bp), hashtable_size, pat.ctypes.data_as(_bp), _d(w), _i(smap))
    if P < 0:
        raise MemoryError("oracle_compress_patterns failed")
    return pat[: T * P].reshape(T, P).copy(), w[:P].copy(), smap


def matrix_gradient(pb: Problem, M) -> np.ndarray:
    """Node sweep of calculate_dlnl_dQ (treelikelihood.c:2337-2583) from the restatement's own lower / upper partials:
    out[k] = sum_n sum_p w_p / L_p sum_c prop_c sum_i f_i U_n[c,p,i] (M_k[n,c] L_n[c,p])_i, unscaled form."""
    assert not pb.scale
    res = evaluate(pb, partials=True)
    N, Cc, P, S, T = pb.nnodes, pb.ncat, pb.npatterns, pb.nstate, pb.ntips
    lower = res["lower"].copy()
    for t in range(T):  # tips: indicator vectors, ones for unknown states
        if pb.use_tip_states:
            tp = np.ones((P, S))
            known = pb.tip_states[t] < S
            tp[known] = np.eye(S)[pb.tip_states[t][known]]
        else:
            tp = pb.tip_partials[t]
        lower[t] = tp[None]
    fq = np.ones(S) if pb.include_root_freqs else pb.freqs
    wl = pb.weights / np.exp(res["pattern_lnl"])
    skip = {int(pb.root)} | ({int(pb.right[pb.root])} if pb.unrooted else set())
    out = np.zeros(len(M))
    for k, Mk in enumerate(M):
        tot = 0.0
        for n in range(N):
            if n in skip:
                continue
            ml = np.einsum("cij,cpj->cpi", Mk[n], lower[n])
            tot += float(np.einsum("c,p,i,cpi,cpi->", pb.props, wl, fq, res["upper"][n], ml))
        out[k] = tot
    return out


def branch_derivatives(pb: Problem, node: int, bls) -> np.ndarray:
    """lnL, d lnL/dt and d2 lnL/dt2 [len(bls)][3] of the branch above `node` at candidate lengths, every other branch as in pb.

    Restates _calculate_uppper (treelikelihood.c:2662-2677: calculate_branch_partials + integrate_partials +
    node_log_likelihoods on the upper partial of `node`, its lower partial and a fresh P(t)), calculate_dldt_uppper (:2195-2262,
    matrix r dP/dt) and d2lnldt2_uppper (:2267-2335, matrix r^2 d2P/dt2, :2331 for the second derivative of the log).  Upper and
    lower partials come from the C restatement (oracle_evaluate, update_upper_partials with include_root_freqs = false as
    SingleTreeLikelihood_update_uppers does, :1535).  Unscaled problems only."""
    from dataclasses import replace
    assert not pb.scale and pb.evec is not None
    res = evaluate(replace(pb, include_root_freqs=False), gradient=True, partials=True)
    S, Cc, P, T = pb.nstate, pb.ncat, pb.npatterns, pb.ntips
    U = res["upper"][node]  # [C][P][S]
    if node >= T:
        L = res["lower"][node]
    elif pb.use_tip_states:  # a known state selects a column, an unknown one the row sum (treelikelihoodX.c:878-1001)
        st = pb.tip_states[node].astype(int)
        one = np.where(st[:, None] < S, np.eye(S + 1)[np.minimum(st, S)][:, :S], 1.0)
        L = np.broadcast_to(one, (Cc, P, S))
    else:
        L = np.broadcast_to(pb.tip_partials[node], (Cc, P, S))
    lib = _oracle()
    ev, la, iv = (np.ascontiguousarray(a, dtype=np.float64) for a in (pb.evec, pb.eval, pb.ivec))
    out = np.zeros((len(bls), 3))
    for k, t in enumerate(bls):
        lk = np.zeros(P), np.zeros(P), np.zeros(P)
        for c in range(Cc):
            r = float(pb.rates[c])
            M0, M1 = np.zeros((S, S)), np.zeros((S, S))
            lib.oracle_p_t(S, _d(ev), _d(la), _d(iv), float(t) * r, _d(M0))
            lib.oracle_dp_dt(S, _d(ev), _d(la), _d(iv), float(t) * r, _d(M1))
            M2 = (ev * (la * la * np.exp(la * float(t) * r))[None, :]) @ iv  # d2p_d2t, substmodel.c:801-828
            prop = 1.0 if Cc == 1 else float(pb.props[c])
            for q, M in enumerate((M0, M1 * r, M2 * r * r)):
                lk[q][:] += prop * np.einsum("pi,ij,pj->p", U[c] * pb.freqs[None, :], M, L[c])
        w = pb.weights
        out[k] = [np.sum(w * np.log(lk[0])), np.sum(w * lk[1] / lk[0]), np.sum(w * (lk[2] * lk[0] - lk[1] ** 2) / lk[0] ** 2)]
    return out


def time_evaluate(pb: Problem, tip_heights, ratios, rates, include_jacobian=False) -> dict:
    """Time-tree chain around `evaluate` (naive reference forms): ratios [T-1] (root entry = root height), rates [1] or [N]
    -> dict(lnl, log_jacobian, heights, bl, grad_ratios, grad_rates)."""
    import copy

    lib = _oracle()
    T, N = pb.ntips, pb.nnodes
    left, right, parent = (np.ascontiguousarray(a, dtype=np.int32) for a in (pb.left, pb.right, pb.parent))
    th = np.ascontiguousarray(tip_heights, dtype=np.float64)
    r = np.ascontiguousarray(ratios, dtype=np.float64)
    c = np.ascontiguousarray(rates, dtype=np.float64)
    lowers, heights, bl, lj = np.zeros(N), np.zeros(N), np.zeros(N), C.c_double(0)
    lib.oracle_time_forward.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_double)]
    lib.oracle_time_backward.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _dp, _dp, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp]
    lib.oracle_time_forward(T, N, int(pb.root), _i(left), _i(right), _i(parent), _d(th), _d(r), _d(c), c.shape[0], _d(lowers), _d(heights),
                            _d(bl), C.byref(lj))
    q = copy.copy(pb)
    q.bl, q.unrooted = bl, False
    out = evaluate(q)
    gr, gc = np.zeros(T - 1), np.zeros(c.shape[0])
    lib.oracle_time_backward(T, N, int(pb.root), _i(left), _i(right), _i(parent), _d(r), _d(c), c.shape[0], _d(lowers), _d(heights),
                             _d(np.ascontiguousarray(out["grad"])), int(include_jacobian), _d(gr), _d(gc))
    return dict(lnl=out["lnl"], log_jacobian=lj.value, heights=heights, bl=bl, lowers=lowers, grad_ratios=gr, grad_rates=gc, grad=out["grad"])


# ---------------------------------------------------------------------------------------------
# compiled reference (oracle/_ref)
# ---------------------------------------------------------------------------------------------

FLAG_TREE_MODEL = 1 << 0  # treelikelihood.h:32-38
FLAG_SITE_MODEL = 1 << 1
FLAG_SUBSTITUTION_MODEL = 1 << 2
FLAG_BRANCH_MODEL = 1 << 6

_ref = None


def reference_available() -> bool:
    return os.path.exists(REF_SO)


def _reflib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(REF_SO)
        L.refh_create.argtypes = [C.c_char_p]
        L.refh_create.restype = C.c_void_p
        L.refh_free.argtypes = [C.c_void_p]
        L.refh_create_codon.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_double, C.c_double]
        L.refh_create_codon.restype = C.c_void_p
        L.refh_dims.argtypes = [C.c_void_p, _ip]
        L.refh_topology.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.refh_branch_lengths.argtypes = [C.c_void_p, _dp, _dp]
        L.refh_set_branch_lengths.argtypes = [C.c_void_p, _dp]
        L.refh_tip_states.argtypes = [C.c_void_p, _bp]
        L.refh_tip_partials.argtypes = [C.c_void_p, _dp]
        L.refh_weights.argtypes = [C.c_void_p, _dp]
        L.refh_model.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.refh_model.restype = C.c_int
        L.refh_sitemodel.argtypes = [C.c_void_p, _dp, _dp]
        L.refh_matrices.argtypes = [C.c_void_p, _dp, _dp]
        L.refh_use_rescaling.argtypes = [C.c_void_p, C.c_int]
        L.refh_rescaling.argtypes = [C.c_void_p]
        L.refh_rescaling.restype = C.c_int
        L.refh_enable_sse.argtypes = [C.c_void_p, C.c_int]
        L.refh_set_include_jacobian.argtypes = [C.c_void_p, C.c_int]
        L.refh_use_generic_kernels.argtypes = [C.c_void_p]
        L.refh_logP.argtypes = [C.c_void_p]
        L.refh_logP.restype = C.c_double
        L.refh_pattern_lnl.argtypes = [C.c_void_p, _dp]
        L.refh_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.c_int]
        L.refh_gradient.restype = C.c_int
        L.refh_partials.argtypes = [C.c_void_p, C.c_int, _dp]
        L.refh_partials.restype = C.c_int
        L.refh_scaling_factors.argtypes = [C.c_void_p, C.c_int, _dp]
        L.refh_scaling_factors.restype = C.c_int
        L.refh_time_logP.argtypes = [C.c_void_p, C.c_int, _dp]
        L.refh_time_logP.restype = C.c_double
        L.refh_time_gradient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.refh_time_gradient.restype = C.c_double
        L.refh_patterns_raw.argtypes = [C.c_void_p, _bp, _dp]
        L.refh_patterns_raw.restype = C.c_int
        L.refh_pattern_name.argtypes = [C.c_void_p, C.c_int]
        L.refh_pattern_name.restype = C.c_char_p
        L.refh_dPdp.argtypes = [C.c_void_p, C.c_int, _dp]
        L.refh_dlnl_dQ.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refh_dlnl_dQ.restype = C.c_double
        L.refh_branch_derivatives.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp]
        L.refh_branch_derivatives.restype = C.c_int
        L.refh_time_tree.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.refh_time_tree.restype = C.c_int
        L.refh_set_ratios.argtypes = [C.c_void_p, _dp]
        L.refh_set_clock_rate.argtypes = [C.c_void_p, C.c_double]
        _ref = L
    return _ref


class Reference:
    """One `treelikelihood` Model of the unmodified reference, built from a JSON dict."""

    def __init__(self, spec: dict | None = None, codon: dict | None = None):
        self.L = _reflib()
        if codon is not None:
            names = list(codon["sequences"].keys())
            n = len(names)
            a_names = (C.c_char_p * n)(*[s.encode() for s in names])
            a_seqs = (C.c_char_p * n)(*[codon["sequences"][s].encode() for s in names])
            self.h = self.L.refh_create_codon(codon["newick"].encode(), n, a_names, a_seqs,
                                              float(codon["kappa"]), float(codon["omega"]))
        else:
            text = json.dumps({"model": spec}).encode()
            self.h = self.L.refh_create(text)
        dims = np.zeros(10, np.int32)
        self.L.refh_dims(self.h, _i(dims))
        (self.T, self.N, self.S, self.C, self.P, self.root, self.time_mode, _scale, self.use_tip_states, self.root_right) = [int(x) for x in dims]

    def close(self):
        if self.h:
            self.L.refh_free(self.h)
            self.h = None

    def logP(self) -> float:
        return float(self.L.refh_logP(self.h))

    def pattern_lnl(self):
        out = np.zeros(self.P)
        self.L.refh_pattern_lnl(self.h, _d(out))
        return out

    def gradient(self, flags=FLAG_TREE_MODEL, include_root_freqs=-1):
        cap = 4 * self.N + 64
        out = np.zeros(cap)
        n = self.L.refh_gradient(self.h, flags, include_root_freqs, _d(out), cap)
        return out[:n].copy()

    def use_rescaling(self, use: bool):
        self.L.refh_use_rescaling(self.h, int(use))

    def rescaling(self) -> bool:
        return bool(self.L.refh_rescaling(self.h))

    def enable_sse(self, v: bool):
        self.L.refh_enable_sse(self.h, int(v))

    def set_include_jacobian(self, v: bool):
        self.L.refh_set_include_jacobian(self.h, int(v))

    def use_generic_kernels(self):
        self.L.refh_use_generic_kernels(self.h)

    def set_branch_lengths(self, bl):
        bl = np.ascontiguousarray(bl, dtype=np.float64)
        self.L.refh_set_branch_lengths(self.h, _d(bl))

    def partials(self, idx):
        out = np.zeros((self.C, self.P, self.S))
        ok = self.L.refh_partials(self.h, idx, _d(out))
        return out if ok else None

    def scaling_factors(self, idx):
        out = np.zeros(self.P)
        ok = self.L.refh_scaling_factors(self.h, idx, _d(out))
        return out if ok else None

    def time_logP(self, iters):
        last = C.c_double(0)
        return float(self.L.refh_time_logP(self.h, iters, C.byref(last))), last.value

    def time_gradient(self, iters, flags=FLAG_TREE_MODEL, include_root_freqs=-1):
        return float(self.L.refh_time_gradient(self.h, flags, include_root_freqs, iters))

    def patterns_raw(self):
        """sp->patterns [sequences][P] in ALIGNMENT order, sp->weights and sp->names exactly as new_SitePattern2 left them."""
        pat = np.zeros((self.T, self.P), np.uint8)
        w = np.zeros(self.P)
        n = self.L.refh_patterns_raw(self.h, pat.ctypes.data_as(_bp), _d(w))
        names = [self.L.refh_pattern_name(self.h, i).decode() for i in range(n)]
        return pat, w, names

    def dPdp(self, index):
        out = np.zeros((self.N, self.C, self.S, self.S))
        self.L.refh_dPdp(self.h, int(index), _d(out))
        return out

    def dlnl_dQ(self, index, include_root_freqs=0):
        return float(self.L.refh_dlnl_dQ(self.h, int(index), int(include_root_freqs)))

    def branch_derivatives(self, node, bl):
        """(lnL, d lnL/dt, d2 lnL/dt2) of the branch above `node` at length `bl` from the reference's own _calculate_uppper,
        calculate_dldt_uppper and d2lnldt2_uppper (treelikelihood.c:2592-2686, 2195-2335)."""
        out = np.zeros(3)
        if _reflib().refh_branch_derivatives(self.h, int(node), float(bl), _d(out)) != 0:
            raise ValueError(f"node {node}: the reference's upper-likelihood functions exclude the root and its right child")
        return out

    def time_tree(self):
        """(tip_heights[T], ratios[T-1] with the root height in the root's entry, rates[1 or N]) as the reference holds them."""
        th, ratios, rates = np.zeros(self.T), np.zeros(self.T - 1), np.zeros(self.N)
        nr = self.L.refh_time_tree(self.h, _d(th), _d(ratios), _d(rates))
        return th, ratios, rates[:nr].copy() if nr == 1 else rates

    def set_ratios(self, ratios):
        self.L.refh_set_ratios(self.h, _d(np.ascontiguousarray(ratios, dtype=np.float64)))

    def set_clock_rate(self, rate):
        self.L.refh_set_clock_rate(self.h, float(rate))

    def problem(self, **kw) -> Problem:
        """Export every input of the hot path as plain arrays (ids and pattern order as the reference has them)."""
        N, S, Cc, P, T = self.N, self.S, self.C, self.P, self.T
        left, right, parent = (np.zeros(N, np.int32) for _ in range(3))
        self.L.refh_topology(self.h, _i(left), _i(right), _i(parent))
        bl, dt = np.zeros(N), np.zeros(N)
        self.L.refh_branch_lengths(self.h, _d(bl), _d(dt))
        ts = np.zeros((T, P), np.uint8)
        self.L.refh_tip_states(self.h, ts.ctypes.data_as(_bp))
        tp = np.zeros((T, P, S))
        self.L.refh_tip_partials(self.h, _d(tp))
        w = np.zeros(P)
        self.L.refh_weights(self.h, _d(w))
        evec, ev, ivec, freqs = np.zeros((S, S)), np.zeros(S), np.zeros((S, S)), np.zeros(S)
        has_eigen = self.L.refh_model(self.h, _d(evec), _d(ev), _d(ivec), _d(freqs))
        rates, props = np.zeros(Cc), np.zeros(Cc)
        self.L.refh_sitemodel(self.h, _d(rates), _d(props))
        pb = Problem(left=left, right=right, parent=parent, root=self.root, nstate=S, tip_states=ts, weights=w,
                     freqs=freqs, rates=rates, props=props, bl=bl, tip_partials=tp,
                     use_tip_states=bool(self.use_tip_states), unrooted=not self.time_mode)
        if has_eigen:
            pb.evec, pb.eval, pb.ivec = evec, ev, ivec
        else:
            Pm, dPm = self.matrices()
            pb.P_override, pb.dP_override = Pm, dPm
        pb.meta["time_elapsed"] = dt
        for k, v in kw.items():
            setattr(pb, k, v)
        return pb

    def matrices(self):
        Pm = np.zeros((self.N, self.C, self.S, self.S))
        dPm = np.zeros_like(Pm)
        self.L.refh_matrices(self.h, _d(Pm), _d(dPm))
        return Pm, dPm


def treelikelihood_spec(newick: str, sequences: dict, model: dict, categories: int = 1, alpha: float = 0.5,
                        tipstates: bool = False, sse: bool = True, datatype: str = "nucleotide",
                        time_tree: dict | None = None) -> dict:
    """JSON for the reference's `treelikelihood` object (treelikelihood.c:819-942), inline data only."""
    sitemodel = {"id": "sitemodel", "type": "sitemodel", "substitutionmodel": model}
    if categories > 1:
        sitemodel["distribution"] = {
            "distribution": "gamma", "categories": categories,
            "parameters": {"shape": {"id": "alpha", "type": "parameter", "value": alpha, "lower": 0}},
        }
    tree = {"id": "tree", "type": "tree", "parameters": "tree.distances", "newick": newick}
    spec = {
        "id": "treelikelihood", "type": "treelikelihood", "tipstates": tipstates, "sse": sse,
        "sitepattern": {"id": "patterns", "type": "sitepattern", "datatype": datatype,
                        "alignment": {"id": "seqs", "type": "alignment", "sequences": sequences}},
        "sitemodel": sitemodel, "tree": tree,
    }
    if time_tree:
        tree.update(time_tree["tree"])
        spec["branchmodel"] = time_tree["branchmodel"]
    return spec


def nucleotide_model_spec(name: str, freqs=None, rates=None, kappa=None) -> dict:
    freqs = [0.25] * 4 if freqs is None else list(map(float, freqs))
    m = {"id": "sm", "type": "substitutionmodel", "model": name, "datatype": "nucleotide",
         "frequencies": {"id": "freqs", "type": "Simplex", "values": freqs}}
    if name == "gtr":
        m["rates"] = {"id": "gtr_rates", "type": "Simplex", "values": list(map(float, rates))}
    if name == "hky":
        m["rates"] = {"kappa": {"id": "kappa", "type": "parameter", "value": float(kappa), "lower": 0}}
    return m
