"""Expectations and perturbed free energies without the N x K weight matrix.

pymbar computes <A> and the free energies of new states by appending columns to the host-resident
``Log_W_nk`` (mbar.py:886-940): one column per new state l (energies u_ln) and one per observable
(log A_n added to that state's log-weights), then takes the asymptotic covariance of the augmented
N x (K + NL + S) matrix.  Every appended column is an *unsampled state* of an augmented MBAR problem:

    new state l            energies  u_ln
    observable i at state l  energies  u_ln - log(A_in - A_min_i + logfactor_i)

so the existing kernels do all of it: the all-state self-consistent update returns the appended states'
free energies (``log_C_a`` / ``f_k[sa]`` of mbar.py:924-940) and ``weight_moments`` returns the
(K + NL + S)^2 second moments that Theta needs.  The K x K algebra on top is the reference's.
"""
from __future__ import annotations

import numpy as np

from . import estimators as est
from .problem import DeviceProblem
from .utils import ParameterError


def expectations_inner(u_kn, N_k, f_k, A_n, u_ln, state_map, uncertainty_method=None, return_theta=False,
                       device=0):
    """MBAR.compute_expectations_inner (mbar.py:766-1012) for the analytical (non-bootstrap) methods.

    Parameters follow the reference: A_n [I, N] observables, u_ln [L, N] energies of the states of
    interest, state_map either a 1-D list of states (free energies only) or [2, S] rows
    (state index into u_ln, observable index into A_n).  Returns the same dictionary keys:
    'observables', 'f', 'Theta', 'Amin'."""
    logfactor = 4.0 * np.finfo(np.float64).eps
    u_kn = np.asarray(u_kn, dtype=np.float64)
    N_k = np.asarray(N_k)
    f_k = np.asarray(f_k, dtype=np.float64)
    K, N = u_kn.shape
    state_map = np.asarray(state_map)
    if state_map.ndim < 2:
        state_list = state_map.astype(int).copy()
        state_map = np.zeros((0, 0), int)
        S = 0
    else:
        state_map = state_map.astype(int)
        state_list = state_map[0, :]
        S = state_map.shape[1]
    u_ln = np.asarray(u_ln, dtype=np.float64)
    if u_ln.ndim == 1:
        u_ln = u_ln.reshape(1, -1)
    A_n = np.array(A_n, dtype=np.float64)
    if A_n.ndim == 1:
        A_n = A_n.reshape(1, -1)

    L_list = np.unique(state_list)
    NL = len(L_list)
    if S > 0:
        A_list = np.unique(state_map[1, :])
        A_min = np.zeros(len(A_list))
    else:
        A_list = np.zeros(0, dtype=int)
        A_min = np.zeros(0)
    logfactors = np.zeros(len(A_list))
    for i in A_list:
        A_min[i] = np.min(A_n[i, :])
        logfactors[i] = np.abs(logfactor * A_min[i])
        A_n[i, :] = A_n[i, :] - (A_min[i] - logfactors[i])

    # augmented problem: K original states, then the NL states of interest, then the S observables
    msize = K + NL + S
    aug = np.empty((msize, N), dtype=np.float64)
    aug[:K] = u_kn
    for l in L_list:
        aug[K + l] = u_ln[l]
    with np.errstate(divide="ignore"):
        for s in range(S):
            aug[K + NL + s] = u_ln[state_map[0, s]] - np.log(A_n[state_map[1, s]])
    N_aug = np.zeros(msize)
    N_aug[:K] = N_k
    f_aug = np.zeros(msize)
    f_aug[:K] = f_k
    result = {}
    with DeviceProblem(aug, N_aug, device=device) as p:
        f_new = p.self_consistent_update(f_aug)            # appended rows: -logsumexp_n(-v_an - L_n)
        f_aug[K:] = f_new[K:]
        if return_theta:
            _, G = p.weight_moments(f_aug)
    # observable estimates: exp(f_l - f_s) + the constant removed for positivity (mbar.py:943-953)
    if S > 0:
        A_i = np.exp(f_aug[K + state_map[0, :]] - f_aug[K + NL + np.arange(S)])
        A_i += A_min[state_map[1, :]] - logfactors[state_map[1, :]]
        result["observables"] = A_i
    result["f"] = f_aug[K + state_list]
    if return_theta:
        Theta_ij = est.asymptotic_covariance(G, N_aug, method=uncertainty_method)
        si = K + NL + np.arange(S) if S > 0 else np.zeros(0, dtype=int)
        li = K + state_list
        idx = np.concatenate((si, li)).astype(int)
        result["Theta"] = Theta_ij[np.ix_(idx, idx)]
        if S > 0:
            result["Amin"] = A_min[state_map[1, np.arange(S)]] - logfactors[state_map[1, np.arange(S)]]
    return result


def compute_expectations(u_kn, N_k, f_k, A_n, u_ln=None, output="averages", state_dependent=False,
                         compute_uncertainty=True, uncertainty_method=None, warning_cutoff=1.0e-10,
                         return_theta=False, device=0):
    """MBAR.compute_expectations (mbar.py:1124-1312) with 2-D inputs: A_n [N] (or [K, N] when
    state_dependent), optional u_ln [L, N] for states other than the sampled set."""
    if uncertainty_method == "bootstrap":
        raise ParameterError("bootstrap uncertainties are served by pymbar_b200.bootstrap, not here")
    u_kn = np.asarray(u_kn, dtype=np.float64)
    u = u_kn if u_ln is None else np.asarray(u_ln, dtype=np.float64)
    if u.ndim == 1:
        u = u.reshape(1, -1)
    Ks = u.shape[0]
    state_map = np.zeros((2, Ks), int)
    state_map[0] = np.arange(Ks)
    state_map[1] = np.arange(Ks) if state_dependent else 0
    inner = expectations_inner(u_kn, N_k, f_k, A_n, u, state_map, uncertainty_method=uncertainty_method,
                               return_theta=compute_uncertainty or return_theta, device=device)
    out = {}
    if compute_uncertainty or return_theta:
        diag = np.ones(2 * Ks)
        diag[:Ks] = diag[Ks:] = inner["observables"] - inner["Amin"]
        Theta = np.diag(diag) @ inner["Theta"] @ np.diag(diag)
        covA = Theta[:Ks, :Ks] + Theta[Ks:, Ks:] - Theta[:Ks, Ks:] - Theta[Ks:, :Ks]
    if output == "averages":
        out["mu"] = inner["observables"]
        if compute_uncertainty:
            out["sigma"] = np.sqrt(covA.diagonal())
    elif output == "differences":
        A = inner["observables"]
        out["mu"] = A - np.vstack(A)
        if compute_uncertainty:
            out["sigma"] = est.error_of_differences(covA, warning_cutoff=warning_cutoff)
    else:
        raise ParameterError(f"output={output!r} must be 'averages' or 'differences'")
    if return_theta:
        out["Theta"] = Theta
    return out


def compute_perturbed_free_energies(u_kn, N_k, f_k, u_ln, compute_uncertainty=True, uncertainty_method=None,
                                    warning_cutoff=1.0e-10, device=0):
    """MBAR.compute_perturbed_free_energies (mbar.py:1442-1521): free energies of L new states."""
    u_ln = np.asarray(u_ln, dtype=np.float64)
    if u_ln.ndim == 1:
        u_ln = u_ln.reshape(1, -1)
    L = u_ln.shape[0]
    inner = expectations_inner(u_kn, N_k, f_k, np.array([0.0]), u_ln, np.arange(L),
                               uncertainty_method=uncertainty_method, return_theta=compute_uncertainty,
                               device=device)
    f = inner["f"]
    out = {"Delta_f": f - np.vstack(f)}
    if compute_uncertainty:
        out["dDelta_f"] = est.error_of_differences(inner["Theta"], warning_cutoff=warning_cutoff)
    return out
