"""CPU restatement (test infrastructure, never shipped) of the reference's Dirichlet priors and Bayesian score:
src/DiscreteBayesNet/dirichlet_priors.jl (UniformPrior / BDeuPrior) and
src/DiscreteBayesNet/structure_scoring.jl:17-94 (bayesian_score_component, bayesian_score_components,
bayesian_score).  Works on the sufficient statistics N_i (r x q, column-major) that `statistics` returns, which is
exactly what the reference computes inline with `sparse(...)` before taking loggamma."""
import math

import numpy as np


def uniform_alpha(r: int, q: int, alpha: float = 1.0) -> np.ndarray:
    """UniformPrior(α): every pseudo-count is α (dirichlet_priors.jl: get(::UniformPrior, ...) = fill(α, r, q))."""
    return np.full((r, q), float(alpha))


def bdeu_alpha(r: int, q: int, x: float = 1.0) -> np.ndarray:
    """BDeuPrior(x): α_ijk = x / (r_i q_i) (dirichlet_priors.jl: get(::BDeuPrior, ...))."""
    return np.full((r, q), float(x) / (r * q))


def bayesian_score_component(N: np.ndarray, alpha: np.ndarray) -> float:
    """structure_scoring.jl:17-42:
       Σ lnΓ(α + N) − Σ lnΓ(α) + Σ_j lnΓ(Σ_k α_kj) − Σ_j lnΓ(Σ_k α_kj + Σ_k N_kj)."""
    N = np.asarray(N, np.float64)
    a = np.asarray(alpha, np.float64)
    lg = np.vectorize(math.lgamma)
    return float(lg(a + N).sum() - lg(a).sum() + lg(a.sum(axis=0)).sum() - lg(a.sum(axis=0) + N.sum(axis=0)).sum())


def split_counts(card, par_ptr, par_idx, cpt_off, counts):
    """flat counts -> list of r x q matrices (column-major), one per node."""
    out = []
    for i in range(len(card)):
        q = 1
        for k in range(par_ptr[i], par_ptr[i + 1]):
            q *= int(card[par_idx[k]])
        out.append(np.asarray(counts[cpt_off[i]:cpt_off[i + 1]]).reshape((int(card[i]), q), order="F"))
    return out
