"""Host-side model inputs for the tree-likelihood path: rate matrices, eigen systems, Γ categories.

In physher these are produced by host C code that stays host C in a drop-in deployment
(substmodel.c:1092 `update_eigen_system`, eigen.c:115-264, sitemodel.c:573-780) and are handed to
the tree likelihood as plain arrays.  This module builds the same inputs with numpy for the
tests and the benchmark, where the reference library is not linked.  Nothing here runs on the
hot path: one call per model change, S <= 64.

Conventions match the reference: Q[i][j] = rate i -> j, rows sum to 0, normalised so that
-sum_i pi_i Q_ii = 1 (substmodel.c `normalize_Q`); P(t) = evec . diag(exp(eval t)) . ivec.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class SubstitutionModel:
    """Eigen system of a reversible substitution model (what `SubstitutionModel.eigendcmp` holds)."""

    name: str
    nstate: int
    freqs: np.ndarray  # [S]
    evec: np.ndarray  # [S,S]
    eval: np.ndarray  # [S]
    ivec: np.ndarray  # [S,S]
    Q: np.ndarray | None = None

    def p_t(self, t: float) -> np.ndarray:
        """|V exp(Λt) V⁻¹| as substmodel.c:518-557 (fabs on every entry)."""
        return np.abs((self.evec * np.exp(self.eval * t)) @ self.ivec)

    def dp_dt(self, t: float) -> np.ndarray:
        """V Λ exp(Λt) V⁻¹ as substmodel.c:695-723."""
        return (self.evec * (self.eval * np.exp(self.eval * t))) @ self.ivec


def reversible_from_exchangeabilities(name: str, R: np.ndarray, freqs: np.ndarray) -> SubstitutionModel:
    """Q = R·diag(π) off-diagonal, normalised; eigen system through the symmetrised matrix.

    For reversible Q, B = Π^{1/2} Q Π^{-1/2} is symmetric, so Q = Π^{-1/2} W Λ Wᵀ Π^{1/2}
    with orthonormal W (the standard route; the reference uses a general Hessenberg/QR solver,
    eigen.c:115-264, which gives the same P(t) up to rounding).
    """
    freqs = np.asarray(freqs, dtype=np.float64)
    S = freqs.shape[0]
    R = np.asarray(R, dtype=np.float64)
    Q = R * freqs[None, :]
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    Q /= -(freqs * np.diag(Q)).sum()
    sq = np.sqrt(freqs)
    B = (sq[:, None] * Q) / sq[None, :]
    B = 0.5 * (B + B.T)
    lam, W = np.linalg.eigh(B)
    evec = W / sq[:, None]
    ivec = W.T * sq[None, :]
    return SubstitutionModel(name, S, freqs, np.ascontiguousarray(evec), lam, np.ascontiguousarray(ivec), Q)


def jc69() -> SubstitutionModel:
    """JC69 (jc69.c:43-56: off-diagonal 1/3, diagonal -1)."""
    return reversible_from_exchangeabilities("JC69", np.ones((4, 4)), np.full(4, 0.25))


def hky(kappa: float, freqs) -> SubstitutionModel:
    """HKY85, states ordered A,C,G,T; transitions A<->G and C<->T scaled by kappa (hky.c)."""
    R = np.ones((4, 4))
    R[0, 2] = R[2, 0] = kappa
    R[1, 3] = R[3, 1] = kappa
    return reversible_from_exchangeabilities("HKY", R, np.asarray(freqs, dtype=np.float64))


def gtr(rates, freqs) -> SubstitutionModel:
    """GTR with exchangeabilities (AC, AG, AT, CG, CT, GT) (gtr.c)."""
    a, b, c, d, e, f = [float(x) for x in rates]
    R = np.array([[0, a, b, c], [a, 0, d, e], [b, d, 0, f], [c, e, f, 0]], dtype=np.float64)
    return reversible_from_exchangeabilities("GTR", R, np.asarray(freqs, dtype=np.float64))


def random_reversible(nstate: int, seed: int, name: str | None = None) -> SubstitutionModel:
    """A seeded random reversible model with `nstate` states (20: amino-acid-shaped, 61: codon-shaped).

    Used for synthetic benchmark inputs when the reference's empirical matrices are not linked;
    kernel cost depends only on the state count.
    """
    rng = np.random.default_rng(seed)
    R = rng.gamma(shape=0.7, scale=1.0, size=(nstate, nstate)) + 1e-3
    R = 0.5 * (R + R.T)
    freqs = rng.dirichlet(np.full(nstate, 8.0))
    return reversible_from_exchangeabilities(name or f"REV{nstate}", R, freqs)


# universal genetic code, codon order TCAG^3 with the three stops removed (geneticcode.c)
_BASES = "TCAG"
_AA = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"


def gy94(kappa: float, omega: float, freqs=None) -> SubstitutionModel:
    """Goldman-Yang 94 codon model, 61 sense codons of the universal code (gy94.c).

    q_ij = π_j · (κ if transition) · (ω if non-synonymous) for codons differing at one position.
    """
    codons = [a + b + c for a in _BASES for b in _BASES for c in _BASES]
    sense = [i for i in range(64) if _AA[i] != "*"]
    S = len(sense)
    freqs = np.full(S, 1.0 / S) if freqs is None else np.asarray(freqs, dtype=np.float64)
    purines = set("AG")
    R = np.zeros((S, S))
    for x, ci in enumerate(sense):
        for y, cj in enumerate(sense):
            if x == y:
                continue
            diff = [p for p in range(3) if codons[ci][p] != codons[cj][p]]
            if len(diff) != 1:
                continue
            b1, b2 = codons[ci][diff[0]], codons[cj][diff[0]]
            r = 1.0
            if (b1 in purines) == (b2 in purines):
                r *= kappa
            if _AA[ci] != _AA[cj]:
                r *= omega
            R[x, y] = r
    return reversible_from_exchangeabilities("GY94", R, freqs)


def _gammainc_lower_reg(a: float, x: float) -> float:
    """Regularised lower incomplete gamma P(a, x) (series / continued fraction)."""
    if x <= 0:
        return 0.0
    gln = math.lgamma(a)
    if x < a + 1.0:
        ap, s, d = a, 1.0 / a, 1.0 / a
        for _ in range(10000):
            ap += 1.0
            d *= x / ap
            s += d
            if abs(d) < abs(s) * 1e-17:
                break
        return s * math.exp(-x + a * math.log(x) - gln)
    b = x + 1.0 - a
    c = 1.0 / 1e-300
    d = 1.0 / b
    h = d
    for i in range(1, 10000):
        an = -i * (i - a)
        b += 2.0
        d = an * d + b
        d = 1e-300 if abs(d) < 1e-300 else d
        c = b + an / c
        c = 1e-300 if abs(c) < 1e-300 else c
        d = 1.0 / d
        de = d * c
        h *= de
        if abs(de - 1.0) < 1e-17:
            break
    return 1.0 - math.exp(-x + a * math.log(x) - gln) * h


def _gamma_quantile(p: float, a: float) -> float:
    """Quantile of Gamma(shape a, rate a) (mean 1) by bisection on the regularised incomplete gamma."""
    lo, hi = 0.0, 1.0
    while _gammainc_lower_reg(a, hi * a) < p:
        hi *= 2.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if _gammainc_lower_reg(a, mid * a) < p:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def discrete_gamma(alpha: float, ncat: int):
    """Median-quantile discrete Γ rates, mean-normalised, equal proportions (sitemodel.c:573-780,
    QUADRATURE_QUANTILE_MEDIAN).  Returns (rates[C], proportions[C])."""
    if ncat == 1:
        return np.ones(1), np.ones(1)
    q = np.array([_gamma_quantile((2.0 * i + 1.0) / (2.0 * ncat), alpha) for i in range(ncat)])
    q *= ncat / q.sum()
    return q, np.full(ncat, 1.0 / ncat)
