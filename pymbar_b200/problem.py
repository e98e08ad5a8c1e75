"""DeviceProblem: one (u_kn, N_k) data set resident in HBM on one GPU (one shard of the samples).

Thin object wrapper over the C ABI (include/mbar_b200.h).  All math happens in libmbar_b200.so;
this file only marshals numpy buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SolveResult, check


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, K=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if K is not None and a.shape != (K,):
        raise ValueError(f"expected shape ({K},), got {a.shape}")
    return a


class PinnedArray:
    """float64 ndarray backed by cudaHostAlloc memory (mbar_b200_host_alloc)."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape))
        self._ptr = C.c_void_p()
        check(_lib.load().mbar_b200_host_alloc(C.byref(self._ptr), n * 8))
        buf = (C.c_double * n).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64).reshape(self.shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            _lib.load().mbar_b200_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceProblem:
    """u_kn [K, N_local] + global N_k [K] on one B200.

    Parameters
    ----------
    u_kn : ndarray [K, N_local] float64 (any strides with unit inner stride), or None to allocate
        only (fill later with :meth:`synthesize` / :meth:`upload`).
    N_k : array [K] — GLOBAL sample counts (all ranks), int or float; zeros mark unsampled states.
    device : CUDA device ordinal.
    N_local : required when u_kn is None.
    """

    def __init__(self, u_kn, N_k, device=0, N_local=None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        N_k = _f64(N_k)
        if N_k.ndim != 1:
            raise ValueError("N_k must be 1-D")
        self.K = int(N_k.shape[0])
        self.N_k = N_k
        if u_kn is not None:
            if u_kn.ndim != 2 or u_kn.shape[0] != self.K:
                raise ValueError(f"u_kn must be [K={self.K}, N], got {u_kn.shape}")
            N_local = u_kn.shape[1]
        if N_local is None:
            raise ValueError("N_local is required when u_kn is None")
        self.N = int(N_local)
        self.device = int(device)
        check(self._lib.mbar_b200_create(C.byref(self._h), self.device, self.K, self.N, _dptr(N_k)))
        if u_kn is not None:
            self.upload(u_kn)

    # ---- lifetime -------------------------------------------------------------------------
    def close(self):
        if self._h is not None and self._h.value:
            self._lib.mbar_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- data -------------------------------------------------------------------------------
    def upload(self, u_kn):
        u = np.asarray(u_kn)
        if u.dtype != np.float64 or u.ndim != 2 or u.strides[1] != 8 or u.strides[0] % 8 or u.strides[0] < 8 * u.shape[1]:
            u = np.ascontiguousarray(u, dtype=np.float64)
        if u.shape != (self.K, self.N):
            raise ValueError(f"u_kn must be [{self.K}, {self.N}], got {u.shape}")
        check(self._lib.mbar_b200_upload_u_kn(self._h, C.c_void_p(u.ctypes.data), u.strides[0] // 8))

    def augmented(self, u_extra):
        """New DeviceProblem = these samples + the rows of `u_extra` [E, N] as unsampled states (the resident
        tiles are copied on the device, only the new rows are uploaded)."""
        u = np.asarray(u_extra, dtype=np.float64)
        if u.ndim == 1:
            u = u.reshape(1, -1)
        if u.ndim != 2 or u.shape[1] != self.N:
            raise ValueError(f"u_extra must be [E, {self.N}], got {u.shape}")
        if u.strides[1] != 8 or u.strides[0] % 8 or u.strides[0] < 8 * u.shape[1]:
            u = np.ascontiguousarray(u)
        h = C.c_void_p()
        check(self._lib.mbar_b200_create_augmented(self._h, u.shape[0], C.c_void_p(u.ctypes.data),
                                                   u.strides[0] // 8, C.byref(h)))
        q = DeviceProblem.__new__(DeviceProblem)
        q._lib, q._h = self._lib, h
        q.K, q.N, q.device = self.K + u.shape[0], self.N, self.device
        q.N_k = np.concatenate([self.N_k, np.zeros(u.shape[0])])
        return q

    def upload_device_ptr(self, ptr, ld):
        check(self._lib.mbar_b200_upload_u_kn_dev(self._h, C.c_void_p(int(ptr)), int(ld)))

    def synthesize(self, O_k, k_k, seed=0, n_offset=0, N_global=None):
        O_k, k_k = _f64(O_k, self.K), _f64(k_k, self.K)
        spec = _lib.Synth(int(seed), int(n_offset), int(N_global if N_global is not None else self.N),
                          _dptr(O_k), _dptr(k_k))
        check(self._lib.mbar_b200_synthesize(self._h, C.byref(spec)))

    def set_sample_weights(self, w):
        """Per-sample multiplicities (bootstrap replicate without copying u_kn); None restores w = 1."""
        if w is None:
            check(self._lib.mbar_b200_set_sample_weights(self._h, None))
            return
        w = _f64(w)
        if w.shape != (self.N,):
            raise ValueError(f"weights must have shape ({self.N},)")
        check(self._lib.mbar_b200_set_sample_weights(self._h, C.c_void_p(w.ctypes.data)))

    def download(self, n0=0, n=None, out=None):
        n = self.N - n0 if n is None else n
        if out is None:
            out = np.empty((self.K, n), np.float64)
        check(self._lib.mbar_b200_download_u_kn(self._h, int(n0), int(n), C.c_void_p(out.ctypes.data),
                                                out.strides[0] // 8))
        return out

    def set_kernel(self, which):
        code = {"auto": 0, "fused": 1, "generic": 2}.get(which, which)
        check(self._lib.mbar_b200_set_pass_kernel(self._h, int(code)))

    def counters(self):
        v = [C.c_int64(0) for _ in range(4)]
        check(self._lib.mbar_b200_get_counters(self._h, *[C.byref(x) for x in v]))
        return dict(launches=v[0].value, passes=v[1].value, h2d_bytes=v[2].value, d2h_bytes=v[3].value)

    def last_pass_ms(self):
        ms = C.c_double(0)
        check(self._lib.mbar_b200_last_pass_ms(self._h, C.byref(ms)))
        return ms.value

    def last_loop_ms(self):
        tot, ker, it = C.c_double(0), C.c_double(0), C.c_int32(0)
        check(self._lib.mbar_b200_last_loop_ms(self._h, C.byref(tot), C.byref(ker), C.byref(it)))
        return dict(total_ms=tot.value, kernel_ms_sum=ker.value, iters=it.value)

    # ---- communicator -------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        check(_lib.load().mbar_b200_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), _lib.UNIQUE_ID_BYTES)
        check(self._lib.mbar_b200_comm_init(self._h, int(nranks), int(rank), buf))

    def peer_export(self):
        buf = C.create_string_buffer(64)
        check(self._lib.mbar_b200_peer_export(self._h, buf))
        return buf.raw

    def peer_attach(self, nranks, rank, handles):
        blob = b"".join(bytes(h) for h in handles)
        buf = C.create_string_buffer(blob, len(blob))
        check(self._lib.mbar_b200_peer_attach(self._h, int(nranks), int(rank), buf))

    # ---- the pass and the reference primitives ----------------------------------------------------
    def _bad(self, f):
        """The reference propagates NaN through its primitives; the C ABI
        reports ERR_RANGE instead.  Mirror the reference for optimiser trial points that left the
        representable range: non-finite (or > 1e6 in magnitude) f_k on a sampled state -> NaNs."""
        c = f[self.N_k > 0]
        return not np.all(np.isfinite(c)) or np.max(np.abs(c)) > 9.0e5

    def streaming_pass(self, f_k, want_G=False):
        f = _f64(f_k, self.K)
        S = np.empty(self.K)
        sumL = C.c_double(0)
        G = np.empty((self.K, self.K)) if want_G else None
        check(self._lib.mbar_b200_pass(self._h, _dptr(f), _dptr(S), C.byref(sumL),
                                       _dptr(G) if want_G else None))
        return S, sumL.value, G

    def pass_multi(self, f_mk):
        """One call, one synchronisation, M (1 or 2) candidate vectors: returns (S[M, K], sumL[M])."""
        f = np.ascontiguousarray(f_mk, dtype=np.float64)
        if f.ndim != 2 or f.shape[1] != self.K or not 1 <= f.shape[0] <= 2:
            raise ValueError(f"expected [M in (1, 2), {self.K}], got {f.shape}")
        S = np.empty_like(f)
        sumL = np.empty(f.shape[0])
        check(self._lib.mbar_b200_pass_multi(self._h, f.shape[0], _dptr(f), _dptr(S), _dptr(sumL)))
        return S, sumL

    def self_consistent_update(self, f_k):
        f = _f64(f_k, self.K)
        if self._bad(f):
            return np.full(self.K, np.nan)
        out = np.empty(self.K)
        check(self._lib.mbar_b200_self_consistent_update(self._h, _dptr(f), _dptr(out)))
        return out

    def gradient(self, f_k):
        f = _f64(f_k, self.K)
        if self._bad(f):
            return np.full(self.K, np.nan)
        out = np.empty(self.K)
        check(self._lib.mbar_b200_gradient(self._h, _dptr(f), _dptr(out)))
        return out

    def objective_and_gradient(self, f_k):
        f = _f64(f_k, self.K)
        if self._bad(f):
            return np.nan, np.full(self.K, np.nan)
        g = np.empty(self.K)
        obj = C.c_double(0)
        check(self._lib.mbar_b200_objective_and_gradient(self._h, _dptr(f), C.byref(obj), _dptr(g)))
        return obj.value, g

    def objective(self, f_k):
        f = _f64(f_k, self.K)
        if self._bad(f):
            return np.nan
        obj = C.c_double(0)
        check(self._lib.mbar_b200_objective_and_gradient(self._h, _dptr(f), C.byref(obj), None))
        return obj.value

    def hessian(self, f_k):
        f = _f64(f_k, self.K)
        if self._bad(f):
            return np.full((self.K, self.K), np.nan)
        H = np.empty((self.K, self.K))
        check(self._lib.mbar_b200_hessian(self._h, _dptr(f), _dptr(H)))
        return H

    def weight_moments(self, f_k):
        """(S_k, G = W^T W) over ALL states — the K x K input of pymbar_b200.estimators."""
        f = _f64(f_k, self.K)
        S = np.empty(self.K)
        G = np.empty((self.K, self.K))
        check(self._lib.mbar_b200_weight_moments(self._h, _dptr(f), _dptr(S), _dptr(G)))
        return S, G

    def log_W_nk(self, f_k, exponentiate=False, out=None, rows=None, row0=0):
        """[N, K] log weights (mbar_solvers.py:439-473); `rows` / `row0` (multiple of 32) fetch a block of rows."""
        f = _f64(f_k, self.K)
        if self._bad(f):
            n = self.N if rows is None else int(rows)
            return np.full((n, self.K), np.nan)
        if rows is None and row0 == 0:
            if out is None:
                out = np.empty((self.N, self.K), np.float64)
            check(self._lib.mbar_b200_log_W_nk(self._h, _dptr(f), C.c_void_p(out.ctypes.data),
                                               out.strides[0] // 8, int(bool(exponentiate))))
            return out
        rows = self.N - row0 if rows is None else int(rows)
        if out is None:
            out = np.empty((rows, self.K), np.float64)
        check(self._lib.mbar_b200_log_W_nk_rows(self._h, _dptr(f), int(row0), rows, C.c_void_p(out.ctypes.data),
                                                out.strides[0] // 8, int(bool(exponentiate))))
        return out

    def log_denominator(self, f_k):
        f = _f64(f_k, self.K)
        if self._bad(f):                 # the reference propagates NaN
            return np.full(self.N, np.nan)
        out = np.empty(self.N)
        check(self._lib.mbar_b200_log_denominator(self._h, _dptr(f), _dptr(out)))
        return out

    # ---- native loops ---------------------------------------------------------------------------
    @staticmethod
    def _result(r):
        return {name: getattr(r, name) for name, _ in SolveResult._fields_}

    def solve_sci(self, f_k, tol=1e-12, maxiter=10000):
        f = _f64(f_k, self.K).copy()
        r = SolveResult()
        check(self._lib.mbar_b200_solve_sci(self._h, _dptr(f), float(tol), int(maxiter), C.byref(r)))
        return f, self._result(r)

    def solve_adaptive(self, f_k, tol=1e-12, maxiter=10000, min_sc_iter=2, gamma=1.0):
        f = _f64(f_k, self.K).copy()
        r = SolveResult()
        check(self._lib.mbar_b200_solve_adaptive(self._h, _dptr(f), float(tol), int(maxiter), int(min_sc_iter),
                                                 float(gamma), C.byref(r)))
        return f, self._result(r)

    def set_loop_mode(self, mode="device", batch=0):
        """'device' (default): solver iterations stay on the GPU, polled every `batch` iterations;
        'stepped': one host round trip per pass (round-1 behaviour, also the robust fallback)."""
        code = {"device": 0, "stepped": 1}.get(mode, mode)
        check(self._lib.mbar_b200_set_loop_mode(self._h, int(code), int(batch)))

    def loop_stats(self):
        polls, mode, batch = C.c_int64(0), C.c_int32(0), C.c_int32(0)
        check(self._lib.mbar_b200_get_loop_stats(self._h, C.byref(polls), C.byref(mode), C.byref(batch)))
        cap, lau = C.c_int64(0), C.c_int64(0)
        check(self._lib.mbar_b200_get_graph_stats(self._h, C.byref(cap), C.byref(lau)))
        return dict(polls=polls.value, mode="stepped" if mode.value else "device", batch=batch.value,
                    graph_captures=cap.value, graph_launches=lau.value)

    def last_kernels(self):
        a, b = C.create_string_buffer(256), C.create_string_buffer(256)
        check(self._lib.mbar_b200_last_kernels(self._h, a, b, 256))
        return dict(pass_kernel=a.value.decode(), hessian_kernel=b.value.decode())

    def last_hessian_ms(self):
        w, h = C.c_double(0), C.c_double(0)
        check(self._lib.mbar_b200_last_hessian_ms(self._h, C.byref(w), C.byref(h)))
        return dict(weights_ms=w.value, hessian_ms=h.value)

    def sci_iterate(self, f_k, iters):
        f = _f64(f_k, self.K).copy()
        check(self._lib.mbar_b200_sci_iterate(self._h, _dptr(f), int(iters)))
        return f


def measure_fp64_peak(device=0):
    """(DMMA TFLOP/s, DFMA TFLOP/s) of this GPU from register-only loops (mbar_b200_measure_fp64_peak)."""
    a, b = C.c_double(0), C.c_double(0)
    check(_lib.load().mbar_b200_measure_fp64_peak(int(device), C.byref(a), C.byref(b)))
    return a.value, b.value


def gpu_numa_node(device=0):
    """NUMA node of the GPU's PCI function (-1: the host exposes none)."""
    n = C.c_int(-1)
    check(_lib.load().mbar_b200_gpu_numa_node(int(device), C.byref(n)))
    return n.value
