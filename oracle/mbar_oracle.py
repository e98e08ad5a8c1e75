"""numpy restatement of the pymbar MBAR solve path — TEST INFRASTRUCTURE, NOT PRODUCT.

Every function cites the reference lines it restates (paths relative to
/root/reference/pymbar/).  Arithmetic follows the reference's numpy branch
(JAX absent): the same two ``scipy.special.logsumexp`` calls per pass, the same
``numpy.linalg.lstsq`` Newton solve, the same convergence rule.  scipy is the
reference's own (unpinned) third-party dependency (pyproject.toml:27-31;
scipy 1.18.1 / numpy 2.3.5 in this image), so calling it here *is* the
reference's arithmetic rather than a re-derivation of it.

Parity status: PINNED.  tests/test_oracle_golden.py checks this module against
  * the reference's only literal golden vector for the path
    (examples/harmonic-oscillators/harmonic-oscillators.py_output.txt:34-36), and
  * fixtures produced by running the unmodified reference in the build
    container (oracle/make_golden.py -> tests/golden/*.npz).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this module.
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.optimize
from scipy.special import logsumexp

# mbar_solvers.py:102-117 — protocol constants accepted by the drivers.
DEFAULT_SOLVER_PROTOCOL = (
    dict(method="hybr", continuation=True),
    dict(method="adaptive", options=dict(min_sc_iter=0)),
)
ROBUST_SOLVER_PROTOCOL = (
    dict(method="adaptive", options=dict(maxiter=1000)),
    dict(method="L-BFGS-B", options=dict(maxiter=1000)),
)
BOOTSTRAP_SOLVER_PROTOCOL = (dict(method="adaptive", options=dict(min_sc_iter=0)),)

# mbar_solvers.py:120-139
MINIMIZE_METHODS = ("L-BFGS-B", "dogleg", "CG", "BFGS", "Newton-CG", "TNC",
                    "trust-ncg", "trust-krylov", "trust-exact", "SLSQP")
MINIMIZE_NO_HESSIAN = ("L-BFGS-B", "BFGS", "CG", "TNC", "SLSQP")
ROOT_METHODS = ("hybr", "lm")


class OracleParameterError(Exception):
    """utils.py:401-409 ParameterError stand-in."""


def _as_inputs(u_kn, N_k, f_k):
    # mbar_solvers.py:174-203 validate_inputs (dtype/shape contract only).
    u_kn = np.ascontiguousarray(u_kn, dtype=np.float64)
    if u_kn.ndim != 2:
        raise ValueError("u_kn must be 2-D [K, N]")
    K = u_kn.shape[0]
    N_k = np.ascontiguousarray(N_k, dtype=np.float64)
    f_k = np.ascontiguousarray(f_k, dtype=np.float64)
    if N_k.shape != (K,) or f_k.shape != (K,):
        raise ValueError("N_k and f_k must have shape (K,)")
    return u_kn, N_k, f_k


def log_denominator_n(u_kn, N_k, f_k):
    """L_n = log sum_k N_k exp(f_k - u_kn)   (mbar_solvers.py:238, :290, :335, :403, :447)."""
    return logsumexp(f_k - u_kn.T, b=N_k, axis=1)


def self_consistent_update(u_kn, N_k, f_k, states_with_samples=None):
    """Eq. C3.  mbar_solvers.py:231-257."""
    if states_with_samples is not None:
        u_kn, N_k, f_k = u_kn[states_with_samples], N_k[states_with_samples], f_k[states_with_samples]
    L_n = log_denominator_n(u_kn, N_k, f_k)
    return -1.0 * logsumexp(-L_n - u_kn, axis=1)


def mbar_gradient(u_kn, N_k, f_k):
    """Eq. C6.  mbar_solvers.py:284-292."""
    L_n = log_denominator_n(u_kn, N_k, f_k)
    log_num_k = logsumexp(-L_n - u_kn, axis=1)
    return -1 * N_k * (1.0 - np.exp(f_k + log_num_k))


def mbar_objective(u_kn, N_k, f_k):
    """mbar_solvers.py:327-338."""
    return np.sum(log_denominator_n(u_kn, N_k, f_k)) - np.dot(N_k, f_k)


def mbar_objective_and_gradient(u_kn, N_k, f_k):
    """mbar_solvers.py:341-355."""
    L_n = log_denominator_n(u_kn, N_k, f_k)
    log_num_k = logsumexp(-L_n - u_kn, axis=1)
    grad = -1 * N_k * (1.0 - np.exp(f_k + log_num_k))
    return np.sum(L_n) - np.dot(N_k, f_k), grad


def mbar_log_W_nk(u_kn, N_k, f_k):
    """Eq. 9, [N, K].  mbar_solvers.py:439-449."""
    L_n = log_denominator_n(u_kn, N_k, f_k)
    return f_k - u_kn.T - L_n[:, np.newaxis]


def mbar_W_nk(u_kn, N_k, f_k):
    """mbar_solvers.py:476-483."""
    return np.exp(mbar_log_W_nk(u_kn, N_k, f_k))


def mbar_hessian(u_kn, N_k, f_k):
    """Eq. C9.  mbar_solvers.py:395-411."""
    W = mbar_W_nk(u_kn, N_k, f_k)
    H = np.dot(W.T, W)
    H *= N_k
    H *= N_k[:, np.newaxis]
    H -= np.diag(W.sum(0) * N_k)
    return -1.0 * H


def precondition_u_kn(u_kn, N_k, f_k):
    """mbar_solvers.py:697-707."""
    u_kn = u_kn - u_kn.min(0)
    u_kn += logsumexp(f_k - u_kn.T, b=N_k, axis=1) - np.dot(N_k, f_k) / N_k.sum()
    return u_kn


def adaptive(u_kn, N_k, f_k, tol=1.0e-8, options=None):
    """Newton / self-consistent adaptive loop.  mbar_solvers.py:510-667 (numpy branch :580-594)."""
    options = {} if options is None else options
    gamma = options.setdefault("gamma", 1.0)
    maxiter = options.setdefault("maxiter", 10000)
    min_sc_iter = options.setdefault("min_sc_iter", 2)
    nr_iter = sci_iter = 0
    success, message = False, "Did not converge."
    g = mbar_gradient(u_kn, N_k, f_k)
    history = []
    for _iteration in range(maxiter):
        H = mbar_hessian(u_kn, N_k, f_k)
        step = np.linalg.lstsq(H, g, rcond=-1)[0]
        step -= step[0]
        f_nr = f_k - gamma * step
        f_sci = self_consistent_update(u_kn, N_k, f_k)
        f_sci = f_sci - f_sci[0]
        g_sci = mbar_gradient(u_kn, N_k, f_sci)
        g_nr = mbar_gradient(u_kn, N_k, f_nr)
        gn_sci, gn_nr = np.dot(g_sci, g_sci), np.dot(g_nr, g_nr)
        f_old = f_k
        if gn_sci < gn_nr or sci_iter < min_sc_iter:       # :607
            f_k, g, sci_iter = f_sci, g_sci, sci_iter + 1
            history.append("sci")
        else:
            f_k, g, nr_iter = f_nr, g_nr, nr_iter + 1
            history.append("nr")
        div = np.abs(f_k[1:])                               # :627-640
        div[div < min(1e-8, tol)] = 1.0
        max_delta = np.max(np.abs(f_k[1:] - f_old[1:]) / div)
        max_diff = np.max(np.abs(f_sci[1:] - f_nr[1:]) / div)
        if np.isnan(max_delta) or (max_delta < tol and max_diff < np.sqrt(tol)):
            success = True
            message = "Convergence achieved by change in f with respect to previous guess."
            break
    return dict(success=success, message=message, x=f_k, nr_iter=nr_iter, sci_iter=sci_iter,
                history=history)


def solve_mbar_once(u_kn, N_k, f_k, method="adaptive", tol=1e-12, continuation=None, options=None):
    """mbar_solvers.py:738-883."""
    u_kn, N_k, f_k = _as_inputs(u_kn, N_k, f_k)
    f_k = f_k - f_k[0]
    u_kn = precondition_u_kn(u_kn, N_k, f_k)
    pad = lambda x: np.pad(x, (1, 0), mode="constant")
    grad = lambda x: mbar_gradient(u_kn, N_k, pad(x))[1:]
    hess = lambda x: mbar_hessian(u_kn, N_k, pad(x))[1:][:, 1:]

    def obj_and_grad(x):
        o, g = mbar_objective_and_gradient(u_kn, N_k, pad(x))
        return np.array(o), np.array(g[1:])

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if method in MINIMIZE_METHODS:
            results = scipy.optimize.minimize(
                obj_and_grad, f_k[1:], jac=True,
                hess=None if method in MINIMIZE_NO_HESSIAN else hess,
                method=method, tol=tol, options=options)
            f_k = pad(results["x"])
        elif method == "adaptive":
            results = adaptive(u_kn, N_k, f_k, tol=tol, options=options)
            f_k = results["x"]
        elif method in ROOT_METHODS:
            results = scipy.optimize.root(grad, f_k[1:], jac=hess, method=method, tol=tol,
                                          options=options)
            f_k = pad(results["x"])
        else:
            raise OracleParameterError(f"Method {method} for solution of free energies not recognized")
    return f_k, results


def solve_mbar(u_kn, N_k, f_k, solver_protocol=None):
    """mbar_solvers.py:886-974 (protocol chain; best gradient norm wins on failure)."""
    protocol = DEFAULT_SOLVER_PROTOCOL if solver_protocol is None else solver_protocol
    fks, gnorms, all_results = [], [], []
    for stage in protocol:
        stage = {k: (dict(v) if isinstance(v, dict) else v) for k, v in stage.items()}
        f_res, results = solve_mbar_once(u_kn, N_k, f_k, **stage)
        fks.append(f_res)
        gnorms.append(np.linalg.norm(mbar_gradient(u_kn, np.asarray(N_k, float), f_res)))
        all_results.append(results)
        if results["success"]:
            break
        if stage.get("continuation"):
            f_k = f_res
    if not results["success"]:
        f_res = fks[int(np.argmin(gnorms))]
    return f_res, all_results


def solve_mbar_for_all_states(u_kn, N_k, f_k, states_with_samples, solver_protocol):
    """mbar_solvers.py:977-1017."""
    f_k = np.array(f_k, dtype=np.float64)
    if len(states_with_samples) == 1:
        f_nonzero = np.array([0.0])
    else:
        f_nonzero, _ = solve_mbar(u_kn[states_with_samples], N_k[states_with_samples],
                                  f_k[states_with_samples], solver_protocol=solver_protocol)
    f_k[states_with_samples] = np.array(f_nonzero)
    f_k = self_consistent_update(u_kn, np.asarray(N_k, float), f_k)
    f_k -= f_k[0]
    return f_k


def mbar_f_k(u_kn, N_k, solver_protocol=None, initial_f_k=None):
    """What MBAR.__init__ computes for f_k (mbar.py:234-243, :325-365, :370-415) with the
    default zero initialisation and no bootstraps."""
    u_kn = np.array(u_kn, dtype=np.float64)
    N_k = np.array(N_k, dtype=np.int64)
    K = u_kn.shape[0]
    if u_kn.shape[1] != N_k.sum():
        raise OracleParameterError("The sum of all N_k must equal the total number of samples")
    sws = np.where(N_k != 0)[0]
    f_k = np.zeros(K) if initial_f_k is None else np.array(initial_f_k, float) - initial_f_k[0]
    protocol = DEFAULT_SOLVER_PROTOCOL if solver_protocol in (None, "default") else solver_protocol
    return solve_mbar_for_all_states(u_kn, N_k, f_k, sws, protocol)


# ---------------------------------------------------------------------------------------------
# Single-pass restatement of the same quantities (the algebra the CUDA kernels use):
# one exp per entry, no second logsumexp.  Used by tests to cross-check the kernel's *formulation*
# against the two-logsumexp reference arithmetic above; never used by the product.
# ---------------------------------------------------------------------------------------------
def single_pass_sums(u_kn, N_k, f_k):
    """Return (S_k, L_n) with S_k = sum_n W_nk and L_n the log-denominator, for N_k > 0 rows."""
    with np.errstate(divide="ignore"):
        c = f_k + np.log(N_k)
    a = c[:, None] - u_kn
    m = a.max(axis=0)
    e = np.exp(a - m)
    D = e.sum(axis=0)
    S = (e / D).sum(axis=1) / N_k
    return S, m + np.log(D)
