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
from .utils import ParameterError


def _as_rows(a):
    a = np.array(a, dtype=np.float64)
    return a.reshape(1, -1) if a.ndim == 1 else a


def expectations_inner(u_kn, N_k, f_k, A_n, u_ln, state_map, uncertainty_method=None, return_theta=False,
                       device=0, problem=None):
    """MBAR.compute_expectations_inner (mbar.py:766-1012) for the analytical (non-bootstrap) methods.

    Same contract as the reference: A_n [I, N] observables, u_ln [L, N] energies of the states of interest,
    state_map either a 1-D list of states (free energies only) or a [2, S] table whose columns are
    (row of u_ln, row of A_n).  Returns the keys 'observables', 'f', 'Theta', 'Amin'.

    The appended rows form an augmented problem on top of the RESIDENT u_kn (`DeviceProblem.augmented`): only
    those rows are uploaded.  `problem` may name the resident DeviceProblem of (u_kn, N_k); otherwise the
    residency cache of `mbar_solvers` provides it."""
    from . import mbar_solvers as ms

    u_kn = np.asarray(u_kn)
    f_k = np.asarray(f_k, dtype=np.float64)
    K = u_kn.shape[0]
    energies = _as_rows(u_ln)
    obs = _as_rows(A_n)
    table = np.asarray(state_map)
    if table.ndim >= 2:
        of_state, of_obs = table[0].astype(int), table[1].astype(int)
    else:
        of_state, of_obs = table.astype(int), np.zeros(0, dtype=int)
    n_pairs = len(of_obs)
    wanted_states = np.unique(of_state)                  # rows of u_ln that become appended states
    n_states = len(wanted_states)

    # observables must be positive to live in the exponent: shift each one used by its minimum (plus a guard
    # of a few ulp so the smallest value does not become log 0), and give the shift back at the end
    guard = 4.0 * np.finfo(np.float64).eps
    floor = {}
    for i in np.unique(of_obs):
        lo = obs[i].min()
        floor[i] = lo - abs(guard * lo)
    # appended rows: the states of interest, then one row per (state, observable) pair whose "energy" is
    # u_l - log(A_i - floor_i): its unsampled-state free energy is -log sum_n (A_i - floor_i) e^{-u_l} / D_n
    extra = np.empty((n_states + n_pairs, u_kn.shape[1]))
    for pos, l in enumerate(wanted_states):
        extra[pos] = energies[l]
    with np.errstate(divide="ignore"):
        for s in range(n_pairs):
            extra[n_states + s] = energies[of_state[s]] - np.log(obs[of_obs[s]] - floor[of_obs[s]])
    row_of_state = {int(l): K + pos for pos, l in enumerate(wanted_states)}
    rows_l = np.array([row_of_state[int(l)] for l in of_state], dtype=int)
    rows_s = K + n_states + np.arange(n_pairs)

    f_aug = np.concatenate([f_k, np.zeros(n_states + n_pairs)])
    N_aug = np.concatenate([np.asarray(N_k, dtype=np.float64), np.zeros(n_states + n_pairs)])

    def run(base):
        with base.augmented(extra) as q:
            f_new = q.self_consistent_update(f_aug)        # appended rows: -logsumexp_n(-v_an - L_n)
            f_aug[K:] = f_new[K:]
            return q.weight_moments(f_aug)[1] if return_theta else None

    if problem is not None:
        G = run(problem)
    else:
        with ms._borrow(u_kn, N_aug[:K]) as base:
            G = run(base)

    out = {}
    if n_pairs:
        shift = np.array([floor[i] for i in of_obs])
        out["observables"] = np.exp(f_aug[rows_l] - f_aug[rows_s]) + shift      # mbar.py:943-953
    out["f"] = f_aug[rows_l]
    if return_theta:
        Theta = est.asymptotic_covariance(G, N_aug, method=uncertainty_method)
        pick = np.concatenate([rows_s, rows_l]).astype(int)        # observables first, then their states
        out["Theta"] = Theta[np.ix_(pick, pick)]
        if n_pairs:
            out["Amin"] = shift
    return out


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
