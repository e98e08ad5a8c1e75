"""Exception and validation surface shared with pymbar.utils (utils.py:117-232, :401-422)."""
import warnings

import numpy as np


class ParameterError(Exception):
    """pymbar.utils.ParameterError (utils.py:401)."""


class ConvergenceError(Exception):
    """pymbar.utils.ConvergenceError (utils.py:408)."""


class TypeCastPerformanceWarning(RuntimeWarning):
    """pymbar.utils.TypeCastPerformanceWarning (utils.py:36)."""


def ensure_type(val, dtype, ndim, name, shape=None, warn_on_cast=True):
    """Same contract as pymbar.utils.ensure_type (utils.py:117-232) for the cases the solver path
    uses: ndarray only, cast to dtype with a warning, ndim / shape checks, C-contiguous result."""
    if not isinstance(val, np.ndarray):
        raise TypeError(f"{name} must be numpy array.  You supplied type {type(val)}")
    dtype = np.dtype(np.float64 if dtype == "float" else dtype)
    if val.dtype != dtype:
        if warn_on_cast:
            warnings.warn(f"Casting {name} dtype={val.dtype} to {dtype} ", TypeCastPerformanceWarning)
        val = val.astype(dtype)
    if val.ndim != ndim:
        raise ValueError(f"{name} must be ndim {ndim}. You supplied {val.ndim}")
    val = np.ascontiguousarray(val)
    if shape is not None:
        if len(shape) != val.ndim or any(s is not None and s != t for s, t in zip(shape, val.shape)):
            raise ValueError(f"{name} must be shape {tuple(shape)}. You supplied  {val.shape}")
    return val
