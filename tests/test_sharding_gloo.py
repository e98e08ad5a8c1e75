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
