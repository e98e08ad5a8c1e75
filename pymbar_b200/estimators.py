"""K x K post-processing of the converged weights, from G = W^T W alone.

pymbar's uncertainty path materialises the N x K weight matrix on the host (``np.exp(self.Log_W_nk)``,
mbar.py:455, :1849) only to reduce it to W^T W.  The Hessian pass already produces that K x K matrix
on the GPU (``DeviceProblem.weight_moments``), so the covariance Theta, the free-energy uncertainties,
the overlap matrix and the effective sample numbers follow without the N x K array ever existing.
The K x K linear algebra below is numpy (eigh / pinv on a K x K matrix, as in the reference).
"""
from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger(__name__)


def asymptotic_covariance(G, N_k, method="svd-ew", tol=1.0e-10):
    """Theta of mbar.py:1756-1864 from G = W^T W.

    'svd-ew' (the reference default, :1837-1858): with G = V S^2 V^T,
    Theta = V S pinv(I - S V^T N V S) S V^T.  'approximate' (:1811-1816): Theta = G."""
    G = np.asarray(G, dtype=np.float64)
    N_k = np.asarray(N_k, dtype=np.float64)
    K = G.shape[0]
    if method in (None, "bootstrap", "svd", "svd-ew"):
        S2, V = np.linalg.eigh(G)
        S2[S2 < 0.0] = 0.0
        Sigma = np.diag(np.sqrt(S2))
        inner = np.identity(K) - Sigma @ V.T @ np.diag(N_k) @ V @ Sigma
        return V @ Sigma @ np.linalg.pinv(inner, rcond=tol) @ Sigma @ V.T
    if method == "approximate":
        return G.copy()
    from .utils import ParameterError

    raise ParameterError(f"Method {method} unrecognized.")


def error_of_differences(cov, warning_cutoff=1.0e-10):
    """Standard deviation of every pairwise difference from a covariance matrix (mbar.py:1687-1715):
    var(x_i - x_j) = C_ii + C_jj - 2 C_ij.  Round-off negatives above -|warning_cutoff| are clipped to zero;
    anything more negative is reported and left in place (its square root is NaN, as in the reference)."""
    var = np.diag(cov)
    spread = np.add.outer(var, var) - 2.0 * np.asarray(cov)
    negative = spread < 0.0
    if negative.any():
        worst = float(spread.min())
        if worst < -abs(warning_cutoff):
            logger.warning("A squared uncertainty is negative. Largest Magnitude = {0:f}".format(abs(worst)))
        else:
            spread = np.where(negative, 0.0, spread)
    return np.sqrt(spread)


def free_energy_differences(f_k, G, N_k, uncertainty_method=None, warning_cutoff=1.0e-10, return_theta=False):
    """compute_free_energy_differences (mbar.py:620-760): Delta_f[i, j] = f_j - f_i and its uncertainty."""
    f_k = np.asarray(f_k, dtype=np.float64)
    out = {"Delta_f": f_k - np.vstack(f_k)}
    Theta = asymptotic_covariance(G, N_k, method=uncertainty_method)
    out["dDelta_f"] = error_of_differences(Theta, warning_cutoff=warning_cutoff)
    if return_theta:
        out["Theta"] = Theta
    return out


def overlap(G, N_k):
    """compute_overlap (mbar.py:563-617): O = N_k * (W^T W), its eigenvalues, 1 - second largest."""
    O = np.asarray(N_k, dtype=np.float64) * np.asarray(G)
    eig = np.sort(np.linalg.eigvals(O))[::-1]
    return {"scalar": 1 - eig[1], "eigenvalues": eig, "matrix": O}


def effective_sample_number(G):
    """compute_effective_sample_number (mbar.py:496-561): N_eff_k = 1 / sum_n W_nk^2 = 1 / G_kk."""
    return 1.0 / np.diag(np.asarray(G))
