"""What `pymbar_b200.install()` does to the `pymbar.MBAR` class itself.

Rebinding `pymbar.mbar_solvers` moves the solve to the GPU, but `MBAR.__init__` then still asks for the full
N x K `Log_W_nk` on the host (mbar.py:455; 20.5 GB at K=256, N=1e7) and every estimator reduces that matrix on
the CPU.  With the facade installed:

* `MBAR.Log_W_nk` becomes LAZY: inside `MBAR.__init__` the backend's `mbar_log_W_nk` hands back a ticket
  instead of the array; the first read of `mbar.Log_W_nk` (property) downloads it, so code that really wants
  the matrix still gets a plain writable ndarray, and code that never looks at it never pays for it;
* `compute_effective_sample_number` (mbar.py:496-561), `compute_overlap` (:563-617) and
  `compute_free_energy_differences` (:620-760) are answered from the K x K second moments
  G = W^T W that the Hessian kernels produce on the device (`weight_moments`), through
  `pymbar_b200.estimators`;
* `compute_expectations_inner` (:766-1012) — the one routine behind compute_expectations,
  compute_multiple_expectations, compute_perturbed_free_energies and compute_entropy_and_enthalpy — is
  answered by the augmented-problem formulation of `pymbar_b200.expectations`.

* `_initialize_with_bar` (:1936-1988, `MBAR(initialize="BAR")`) uses `pymbar_b200.initialize` (samples grouped by
  state once, Brent on Bennett's equation).

Anything outside what the device path implements (bootstrap uncertainties, `uncertainty_method="svd"`) calls
the original method, which then reads `self.Log_W_nk` and materialises it.  `uninstall()` restores the class.
"""
from __future__ import annotations

import threading

import numpy as np

from . import estimators as est

_TLS = threading.local()
_SAVED = {}
STATS = {"tickets": 0, "redeemed": 0, "moments": 0, "expectations": 0}


class LogWeightTicket:
    """Stands for `mbar_log_W_nk(u_kn, N_k, f_k)` until somebody needs the numbers."""

    __slots__ = ("u_kn", "N_k", "f_k")

    def __init__(self, u_kn, N_k, f_k):
        self.u_kn, self.N_k, self.f_k = u_kn, np.array(N_k), np.array(f_k, dtype=np.float64)
        STATS["tickets"] += 1

    def redeem(self):
        from . import mbar_solvers as ms

        STATS["redeemed"] += 1
        return ms.mbar_log_W_nk(self.u_kn, self.N_k, self.f_k)


def deferring():
    return getattr(_TLS, "defer", 0) > 0


def _moments(mbar):
    """(S_k, G = W^T W) of the converged weights, all K states, computed on the device once per MBAR object."""
    cached = mbar.__dict__.get("_b200_moments")
    if cached is not None and np.array_equal(cached[2], mbar.f_k):
        return cached[0], cached[1]
    from . import mbar_solvers as ms

    STATS["moments"] += 1
    with ms._borrow(mbar.u_kn, np.asarray(mbar.N_k, dtype=np.float64)) as p:
        S, G = p.weight_moments(np.asarray(mbar.f_k, dtype=np.float64))
    mbar.__dict__["_b200_moments"] = (S, G, np.array(mbar.f_k))
    return S, G


def _check_normalised(S, tolerance=1.0e-4):
    # utils.check_w_normalized (utils.py:340-388): column sums of W must be 1
    from . import utils as u

    bad = np.flatnonzero(np.abs(S - 1.0) > tolerance)
    if bad.size:
        raise u.ParameterError(f"Warning: Should have \\sum_n W_nk = 1.  Actual column sum for state "
                               f"{int(bad[0]):d} was {S[bad[0]]:f}. {bad.size:d} other columns have similar problems")


def install_on(MBAR):
    """Patch the class object `MBAR` (pymbar.mbar.MBAR)."""
    if MBAR in _SAVED:
        return
    saved = {name: MBAR.__dict__.get(name) for name in
             ("__init__", "Log_W_nk", "compute_effective_sample_number", "compute_overlap",
              "compute_free_energy_differences", "compute_expectations_inner", "_initialize_with_bar")}
    _SAVED[MBAR] = saved
    orig_init = saved["__init__"]
    orig_fed = saved["compute_free_energy_differences"]
    orig_inner = saved["compute_expectations_inner"]

    def __init__(self, *args, **kwargs):
        _TLS.defer = getattr(_TLS, "defer", 0) + 1
        try:
            orig_init(self, *args, **kwargs)
        finally:
            _TLS.defer -= 1

    def _get_logw(self):
        v = self.__dict__.get("_b200_logw")
        if isinstance(v, LogWeightTicket):
            v = v.redeem()
            self.__dict__["_b200_logw"] = v
        return v

    def _set_logw(self, value):
        self.__dict__["_b200_logw"] = value

    def compute_effective_sample_number(self, verbose=False):
        _, G = _moments(self)
        n_eff = est.effective_sample_number(G)
        if verbose:
            import logging

            log = logging.getLogger("pymbar.mbar")
            for k in range(self.K):
                log.info("Effective number of sample in state {:d} is {:10.3f}".format(k, n_eff[k]))
                log.info("Efficiency for state {:d} is {:6f}/{:d} = {:10.4f}".format(k, n_eff[k], self.N,
                                                                                     n_eff[k] / self.N))
        return n_eff

    def compute_overlap(self):
        _, G = _moments(self)
        return est.overlap(G, self.N_k)

    def compute_free_energy_differences(self, compute_uncertainty=True, uncertainty_method=None,
                                        warning_cutoff=1.0e-10, return_theta=False):
        device_ok = uncertainty_method in (None, "svd-ew", "approximate")
        if not device_ok and (compute_uncertainty or return_theta):
            return orig_fed(self, compute_uncertainty=compute_uncertainty, uncertainty_method=uncertainty_method,
                            warning_cutoff=warning_cutoff, return_theta=return_theta)
        Delta = np.array(self.f_k - np.vstack(self.f_k))
        self._zerosamestates(Delta)
        out = {"Delta_f": Delta}
        if compute_uncertainty or return_theta:
            S, G = _moments(self)
            _check_normalised(S)
            Theta = est.asymptotic_covariance(G, self.N_k, method=uncertainty_method)
            if compute_uncertainty:
                d = np.array(est.error_of_differences(Theta, warning_cutoff=warning_cutoff))
                self._zerosamestates(d)
                out["dDelta_f"] = d
            if return_theta:
                out["Theta"] = Theta
        return out

    def compute_expectations_inner(self, A_n, u_ln, state_map, uncertainty_method=None, warning_cutoff=1.0e-10,
                                   return_theta=False):
        if uncertainty_method not in (None, "svd-ew", "approximate"):
            return orig_inner(self, A_n, u_ln, state_map, uncertainty_method=uncertainty_method,
                              warning_cutoff=warning_cutoff, return_theta=return_theta)
        from . import expectations as ex
        from . import mbar_solvers as ms

        STATS["expectations"] += 1
        return ex.expectations_inner(self.u_kn, self.N_k, self.f_k, A_n, u_ln, state_map,
                                     uncertainty_method=uncertainty_method, return_theta=return_theta,
                                     device=ms._DEVICE)

    def _initialize_with_bar(self, u_kn, f_k_init=None):
        from . import initialize as init

        return init.initialize_with_bar(u_kn, self.N_k, self.x_kindices, f_k_init)

    MBAR._initialize_with_bar = _initialize_with_bar
    MBAR.__init__ = __init__
    MBAR.Log_W_nk = property(_get_logw, _set_logw, doc="log weights [N, K] (mbar.py:455), downloaded on first use")
    MBAR.compute_effective_sample_number = compute_effective_sample_number
    MBAR.compute_overlap = compute_overlap
    MBAR.compute_free_energy_differences = compute_free_energy_differences
    MBAR.compute_expectations_inner = compute_expectations_inner


def uninstall_from(MBAR):
    saved = _SAVED.pop(MBAR, None)
    if saved is None:
        return
    for name, value in saved.items():
        if value is None:
            if name in MBAR.__dict__:
                delattr(MBAR, name)
        else:
            setattr(MBAR, name, value)
