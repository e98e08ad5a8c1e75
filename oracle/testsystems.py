"""Input generators for parity tests — TEST INFRASTRUCTURE, NOT PRODUCT.

Restates the two samplers the reference's solver tests draw their u_kn from, so that tests on
the GPU box (where /root/reference does not exist) can regenerate the *same* inputs the golden
fixtures were produced from (legacy ``np.random.seed`` streams are stable across numpy versions):

* harmonic oscillators  — testsystems/harmonic_oscillators.py:154-192
* exponential distributions — testsystems/exponential_distributions.py:148-183
* evenly spaced fixtures — utils_for_testing.py:62-84 (``oscillators`` / ``exponentials``)

make_golden.py asserts, against the real reference, that these reproduce its u_kn bit-for-bit.
"""
from __future__ import annotations

import numpy as np


def harmonic_u_kn(O_k, K_k, N_k, seed=None, beta=1.0):
    """u_kn[l, n] = beta/2 K_l (x_n - O_l)^2, x_n ~ Normal(O_k, (beta K_k)^-1/2) in block order."""
    O_k = np.asarray(O_k, np.float64)
    K_k = np.asarray(K_k, np.float64)
    N_k = np.asarray(N_k, int)
    np.random.seed(seed)
    xs = []
    for k, n in enumerate(N_k):
        sigma = (beta * K_k[k]) ** -0.5
        xs.append(np.random.normal(loc=O_k[k], scale=sigma, size=n))
    x_n = np.concatenate(xs) if xs else np.zeros(0)
    u_kn = np.empty((len(O_k), x_n.size), np.float64)
    for l in range(len(O_k)):
        u_kn[l] = beta * 0.5 * K_k[l] * (x_n - O_k[l]) ** 2.0
    return x_n, u_kn, N_k


def exponential_u_kn(rates, N_k, seed=None, beta=1.0):
    """u_kn[l, n] = beta * rate_l * x_n, x_n ~ Exponential(scale 1/rate_k) in block order."""
    rates = np.asarray(rates, np.float64)
    N_k = np.asarray(N_k, np.int32)
    np.random.seed(seed)
    xs = [np.random.exponential(scale=rates[k] ** -1.0, size=n) for k, n in enumerate(N_k)]
    x_n = np.concatenate(xs) if xs else np.zeros(0)
    u_kn = np.empty((len(rates), x_n.size), np.float64)
    for l in range(len(rates)):
        u_kn[l] = beta * rates[l] * x_n
    return x_n, u_kn, N_k


def oscillators(n_states, n_samples, seed=None):
    """utils_for_testing.py:62-72 with an explicit seed."""
    O_k = np.linspace(1, 5, n_states)
    k_k = np.linspace(1, 3, n_states)
    N_k = (np.ones(n_states) * n_samples).astype("int")
    _, u_kn, N_k = harmonic_u_kn(O_k, k_k, N_k, seed=seed)
    return u_kn, N_k


def exponentials(n_states, n_samples, seed=None):
    """utils_for_testing.py:75-84 with an explicit seed."""
    rates = np.linspace(1, 3, n_states)
    N_k = (np.ones(n_states) * n_samples).astype("int")
    _, u_kn, N_k = exponential_u_kn(rates, N_k, seed=seed)
    return u_kn, N_k


def harmonic_analytical_f_k(K_k, beta=1.0):
    """testsystems/harmonic_oscillators.py:92-96."""
    fe = -0.5 * np.log(2 * np.pi / (beta * np.asarray(K_k, np.float64)))
    return fe - fe[0]


# The golden example of examples/harmonic-oscillators/harmonic-oscillators.py:86-89,99-100,137-140.
GOLDEN_EXAMPLE = dict(
    K_k=[25, 16, 9, 4, 1, 1],
    O_k=[0, 1, 2, 3, 4, 5],
    N_k=[10000, 10000, 10000, 10000, 0, 10000],
    seed=0,
    # harmonic-oscillators.py_output.txt:34-36 "Final dimensionless free energies"
    f_k_printed=[0.0, -0.22821647, -0.49856217, -0.89211081, -1.57434696, -1.57231022],
)
