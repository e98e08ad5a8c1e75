"""World-size-2 (gloo, CPU) check of the multi-GPU decomposition the library relies on: with the
global f_k and N_k replicated, per-shard partial sums S_k, sum L and W^T W add up to the full-data
quantities, so one sum all-reduce per pass is the only exchange.  Also exercises
bench.py's rank plumbing helpers."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from oracle import mbar_oracle as orc
    from tests import _cases

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = _cases.load("small_empty_state")
    u, N = z["u_kn"], z["N_k"].astype(float)
    s = N > 0
    f = z["f_rand"]
    Ntot = u.shape[1]
    lo, hi = Ntot * rank // world, Ntot * (rank + 1) // world
    us = u[s][:, lo:hi]
    # per-shard partials with GLOBAL N_k and f_k
    with np.errstate(divide="ignore"):
        a = (f[s] + np.log(N[s]))[:, None] - us
    m = a.max(0)
    e = np.exp(a - m)
    D = e.sum(0)
    NW = e / D
    part = np.concatenate([NW.sum(1), [np.sum(m + np.log(D))], (NW @ NW.T).ravel()])
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    tot = t.numpy()
    k = s.sum()
    S = tot[:k] / N[s]
    ok = True
    S_ref, L_ref = orc.single_pass_sums(u[s], N[s], f[s])
    ok &= np.allclose(S, S_ref, rtol=1e-12)
    ok &= np.isclose(tot[k], L_ref.sum(), rtol=1e-13)
    H = np.diag(N[s] * S) - tot[k + 1:].reshape(k, k)
    ok &= np.allclose(H, orc.mbar_hessian(u[s], N[s], f[s]), rtol=1e-10, atol=1e-10)
    ok &= np.allclose(f[s] - np.log(S), orc.self_consistent_update(u[s], N[s], f[s]), atol=1e-12)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_partials_sum_over_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_rank_helpers():
    sys.path.insert(0, ROOT)
    import bench

    N_k = bench.global_N_k(256, 2 * 10_000_000)
    assert N_k.sum() == 2e7 and N_k.min() == 78125
    N_k = bench.global_N_k(7, 100)
    assert N_k.sum() == 100 and np.all(N_k[:-1] == 14)
    O, k = bench.workload_params(256)
    assert O[0] == 1 and O[-1] == 5 and k[0] == 1 and k[-1] == 3


def _exchange_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pymbar_b200.sharded import Exchange

    ex = Exchange()
    got = ex.broadcast(b"id-from-rank-0" if rank == 0 else None)
    gathered = ex.all_gather(bytes([rank]) * 4)
    q.put((rank, got == b"id-from-rank-0" and gathered == [bytes([r]) * 4 for r in range(world)]
           and ex.rank == rank and ex.world == world))
    dist.destroy_process_group()


def test_rendezvous_transport_of_the_sharded_entry_points():
    """pymbar_b200.sharded.Exchange (what carries the NCCL id and the cudaIpc handles) under a 2-rank gloo group"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


class _GlooShard:
    """CPU stand-in for a sample-sharded DeviceProblem (test infrastructure): per-shard partial sums with the
    GLOBAL N_k and f_k, one sum all-reduce per pass over a gloo group — the algebra the library implements with
    NCCL / peer memory.  It lets the mirror's sharded entry points (pymbar_b200/sharded.py) and the protocol chain
    above them run on two CPU ranks."""

    def __init__(self, u_local, N_k, exchange=None, device=None, peer=True):
        self.u = np.array(u_local, dtype=np.float64)
        self.N_k = np.asarray(N_k, dtype=np.float64)
        self.K = len(self.N_k)
        self.s = self.N_k > 0

    def __enter__(self): return self
    def __exit__(self, *a): pass
    def close(self): pass

    def _allreduce(self, x, op=dist.ReduceOp.SUM):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        dist.all_reduce(t, op=op)
        return t.numpy()

    def _pass(self, f, want_G=False):
        f = np.asarray(f, float)
        with np.errstate(divide="ignore"):
            a = (f[self.s] + np.log(self.N_k[self.s]))[:, None] - self.u[self.s]
        m = a.max(0)
        e = np.exp(a - m)
        D = e.sum(0)
        L = m + np.log(D)
        NW = e / D
        parts = [NW.sum(1), [L.sum()]] + ([(NW @ NW.T).ravel()] if want_G else [])
        tot = self._allreduce(np.concatenate(parts))
        k = int(self.s.sum())
        S = np.zeros(self.K)
        S[self.s] = tot[:k] / self.N_k[self.s]
        G = tot[k + 1:].reshape(k, k) if want_G else None
        return S, tot[k], G, L

    def streaming_pass(self, f, want_G=False):
        S, sumL, G, _ = self._pass(f, want_G)
        return S, sumL, G

    def gradient(self, f):
        S, _, _, _ = self._pass(f)
        return self.N_k * (S - 1.0) * self.s

    def objective_and_gradient(self, f):
        S, sumL, _, _ = self._pass(f)
        return float(sumL - self.N_k @ np.asarray(f, float)), self.N_k * (S - 1.0) * self.s

    def hessian(self, f):
        S, _, G, _ = self._pass(f, want_G=True)
        H = np.zeros((self.K, self.K))
        H[np.ix_(self.s, self.s)] = np.diag(self.N_k[self.s] * S[self.s]) - G
        return H

    def self_consistent_update(self, f):
        f = np.asarray(f, float)
        _, _, _, L = self._pass(f)
        t = -self.u - L[None, :]                       # log of the un-normalised weights of every state
        mx = self._allreduce(t.max(1), op=dist.ReduceOp.MAX)
        sums = self._allreduce(np.exp(t - mx[:, None]).sum(1))
        return -(mx + np.log(sums))

    def solve_adaptive(self, f, tol=1e-12, maxiter=10000, min_sc_iter=2, gamma=1.0):
        f = np.array(f, float)
        act = np.flatnonzero(self.s)
        f[act] -= f[act[0]]
        nr = sci = 0
        for it in range(maxiter):
            g = self.gradient(f)
            H = self.hessian(f)[np.ix_(act[1:], act[1:])]
            f_nr = f.copy()
            f_nr[act[1:]] -= gamma * np.linalg.solve(H, g[act[1:]])
            f_sci = f.copy()
            f_sci[act] = self.self_consistent_update(f)[act]
            f_sci[act] -= f_sci[act[0]]
            gs, gn = np.linalg.norm(self.gradient(f_sci)), np.linalg.norm(self.gradient(f_nr))
            new = f_sci if (gs < gn or sci < min_sc_iter) else f_nr
            nr, sci = nr + (new is f_nr), sci + (new is f_sci)
            div = np.abs(new[act[1:]])
            div[div < min(1e-8, tol)] = 1.0
            delta = np.max(np.abs(new - f)[act[1:]] / div)
            diff = np.max(np.abs(f_sci - f_nr)[act[1:]] / div)
            f = new
            if delta < tol and diff < np.sqrt(tol):
                return f, dict(success=1, iterations=it + 1, nr_iterations=nr, sci_iterations=sci, max_delta=delta)
        return f, dict(success=0, iterations=maxiter, nr_iterations=nr, sci_iterations=sci, max_delta=delta)


def _sharded_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PYMBAR_B200_CACHE="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pymbar_b200 import mbar_solvers as ms
    from pymbar_b200 import sharded as sh
    from tests import _cases

    sh.ShardedProblem = _GlooShard                      # the device is replaced, the mirror's logic is not
    ok, worst = True, 0.0
    for name in ("small_empty_state", "small_osc_8x40", "osc_50x100"):
        z = _cases.load(name)
        u, N_k = z["u_kn"], z["N_k"]
        lo, hi = u.shape[1] * rank // world, u.shape[1] * (rank + 1) // world
        sws = np.where(N_k != 0)[0]
        for pname, proto in (("default", ms.DEFAULT_SOLVER_PROTOCOL), ("robust", ms.ROBUST_SOLVER_PROTOCOL)):
            proto = tuple({k: (dict(v) if isinstance(v, dict) else v) for k, v in st.items()} for st in proto)
            f = sh.solve_mbar_for_all_states(np.ascontiguousarray(u[:, lo:hi]), N_k, np.zeros(len(N_k)), sws, proto)
            err = float(np.max(np.abs(f - z[f"fk_{pname}"])))
            worst = max(worst, err)
            ok &= err < 1e-8
            same = [None] * world
            dist.all_gather_object(same, f.tobytes())
            ok &= all(b == same[0] for b in same)          # identical decisions and results on every rank
    q.put((rank, bool(ok), worst))
    dist.destroy_process_group()


def test_sharded_mirror_entry_points_on_two_cpu_ranks():
    """pymbar_b200.sharded.solve_mbar_for_all_states (reference signature, this rank's columns only) through the
    protocol chain of the mirror — scipy hybr / adaptive / L-BFGS-B running redundantly on all-reduced primitives —
    on a 2-rank gloo group: reference f_k to 1e-8 and bit-identical across ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 150
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)], res
