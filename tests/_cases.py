"""Shared helpers: load golden fixtures and regenerate their inputs (oracle-side, tests only)."""
import ast
import hashlib
import os

import numpy as np

from oracle import testsystems as ots

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMALL = ["small_osc_8x40", "small_exp_6x50", "small_empty_state", "small_empty_first"]
MEDIUM = ["osc_50x100", "osc_100x100", "osc_200x50", "exp_200x50"]
C1 = ["c1_harmonic_5x1000"]     # BASELINE.json configs[0]: HarmonicOscillatorsTestCase() defaults, K=5, N=5000
ALL = ["golden_example"] + SMALL + MEDIUM + C1


def _regen(spec):
    kind = spec[0]
    if kind == "harmonic":
        _, O, K, N, seed = spec
        return ots.harmonic_u_kn(O, K, N, seed=seed)[1]
    if kind == "exponential":
        _, rates, N, seed = spec
        return ots.exponential_u_kn(rates, N, seed=seed)[1]
    if kind == "osc":
        return ots.oscillators(spec[1], spec[2], seed=spec[3])[0]
    if kind == "exp":
        return ots.exponentials(spec[1], spec[2], seed=spec[3])[0]
    raise ValueError(kind)


def load(name):
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    if "u_kn" not in z:
        spec = ast.literal_eval(str(z["regen"]).replace("np.float64", ""))
        z["u_kn"] = _regen(spec)
    sha = hashlib.sha256(np.ascontiguousarray(z["u_kn"]).tobytes()).hexdigest()
    assert sha == str(z["u_sha"]), f"{name}: regenerated input differs from the fixture's input"
    z["N_k"] = z["N_k"].astype(np.int64)
    return z
