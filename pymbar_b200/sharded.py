"""Sample-sharded (multi-GPU) entry points with the reference's signatures.

One process per GPU; rank r passes ITS columns of u_kn (any contiguous split of the samples) together with the
GLOBAL N_k.  Every reduction over samples is followed by one exchange of the K+2 partial sums (NCCL all-reduce,
or the in-kernel peer-memory exchange for the device-resident loops), so every rank computes bit-identical
K-vectors and takes identical solver decisions: scipy's root / minimize stages run redundantly and
deterministically on all ranks, no broadcast is needed.

    import torch.distributed as dist, pymbar_b200.sharded as sh
    dist.init_process_group("nccl")                       # torchrun: one rank per GPU
    f_k = sh.solve_mbar_for_all_states(u_kn[:, lo:hi], N_k, f0, states_with_samples, protocol)
    logW_local = sh.mbar_log_W_nk(u_kn[:, lo:hi], N_k, f_k)   # rows [lo, hi) of the reference's [N, K] matrix

The rendezvous needs some transport for a 128-byte NCCL id and the 64-byte cudaIpc handles; `Exchange` wraps
torch.distributed (any backend; gloo works for the bytes) but anything with broadcast / all_gather of Python
objects can be substituted (mpi4py, a file, ...).
"""
from __future__ import annotations

import os

import numpy as np

from . import mbar_solvers as ms
from .problem import DeviceProblem


class Exchange:
    """Byte transport for the rendezvous, on torch.distributed."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def broadcast(self, obj):
        box = [obj if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0, group=self.group)
        return box[0]

    def all_gather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out


def attach(problem, exchange=None, peer=True):
    """Join `problem` (this rank's shard) to the communicator of all ranks; peer=True also maps the peer-memory
    inboxes used by the in-kernel exchange of the device-resident loops (same node, NVLink / NVSwitch)."""
    ex = exchange if exchange is not None else Exchange()
    if ex.world == 1:
        return problem
    uid = ex.broadcast(DeviceProblem.comm_unique_id() if ex.rank == 0 else None)
    problem.comm_init(ex.world, ex.rank, uid)
    if peer and not os.environ.get("MBAR_B200_NO_PEER"):
        problem.peer_attach(ex.world, ex.rank, ex.all_gather(problem.peer_export()))
    return problem


def local_device():
    return int(os.environ.get("PYMBAR_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


class ShardedProblem(DeviceProblem):
    """DeviceProblem of this rank's sample slice, attached to all ranks."""

    def __init__(self, u_kn_local, N_k_global, exchange=None, device=None, peer=True):
        super().__init__(u_kn_local, N_k_global, device=local_device() if device is None else device)
        attach(self, exchange, peer=peer)


def solve_mbar_for_all_states(u_kn_local, N_k, f_k, states_with_samples, solver_protocol, exchange=None):
    """mbar_solvers.solve_mbar_for_all_states (mbar_solvers.py:977-1017) on sharded samples: same arguments,
    except that `u_kn_local` holds only this rank's columns.  Returns the same f_k on every rank."""
    u, N_f, f = ms._prep(u_kn_local, N_k, f_k)
    f = np.array(f, dtype=np.float64)
    with ShardedProblem(u, N_f, exchange) as p:
        if len(states_with_samples) > 1:
            f, _ = ms._solve_protocol_on(p, f, solver_protocol)
        else:
            f[np.asarray(states_with_samples)] = 0.0
        f = p.self_consistent_update(f)
    f -= f[0]
    return f


def solve_mbar(u_kn_local, N_k, f_k, solver_protocol=None, exchange=None):
    """mbar_solvers.solve_mbar (mbar_solvers.py:886-974) on sharded samples."""
    u, N_f, f = ms._prep(u_kn_local, N_k, f_k)
    with ShardedProblem(u, N_f, exchange) as p:
        return ms._solve_protocol_on(p, f, solver_protocol)


def mbar_log_W_nk(u_kn_local, N_k, f_k, exchange=None):
    """This rank's rows of mbar_log_W_nk (mbar_solvers.py:439-473): [N_local, K].  (The log weights of a sample
    depend only on its own column and on the global f_k, N_k: no exchange is involved.)"""
    u, N_f, f = ms._prep(u_kn_local, N_k, f_k)
    with DeviceProblem(u, N_f, device=local_device()) as p:
        return p.log_W_nk(f)


def weight_moments(u_kn_local, N_k, f_k, exchange=None):
    """(S_k, G = W^T W) over ALL samples of all ranks — the input of pymbar_b200.estimators."""
    u, N_f, f = ms._prep(u_kn_local, N_k, f_k)
    with ShardedProblem(u, N_f, exchange, peer=False) as p:
        return p.weight_moments(f)
