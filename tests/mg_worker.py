"""Worker for the multi-GPU parity test (launched by torchrun, one rank per GPU): every rank owns a
contiguous slice of the samples of a golden case; all-reduced results must equal the CPU oracle on
the full data set."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from oracle import mbar_oracle as orc
    from pymbar_b200 import DeviceProblem
    from tests import _cases

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    worst = 0.0
    for name in ("small_empty_state", "osc_50x100", "osc_200x50"):
        z = _cases.load(name)
        u, N = z["u_kn"], z["N_k"].astype(float)
        s = N > 0
        Ntot = u.shape[1]
        lo, hi = Ntot * rank // world, Ntot * (rank + 1) // world
        p = DeviceProblem(np.ascontiguousarray(u[:, lo:hi]), N, device=local)
        uid = [DeviceProblem.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        p.comm_init(world, rank, uid[0])
        f = z["f_rand"]
        errs = {
            "sci": np.max(np.abs(p.self_consistent_update(f) - z["rand_sci"])),
            "grad": np.max(np.abs(p.gradient(f)[s] - z["rand_grad"]) / N.max()),
            "obj": abs(p.objective(f) - z["rand_obj"]) / abs(z["rand_obj"]),
            "hess": np.max(np.abs(p.hessian(f)[np.ix_(s, s)] - orc.mbar_hessian(u[s], N[s], f[s]))),
        }
        fk, r = p.solve_adaptive(np.zeros(len(N)), tol=1e-12, min_sc_iter=0)
        ref = np.zeros(len(N))
        ref[s] = z["adaptive_x"]
        errs["adaptive"] = np.max(np.abs(fk[s] - ref[s]))
        f5 = p.sci_iterate(np.zeros(len(N)), 5)          # NCCL all-reduce + epilogue kernel
        handles = [None] * world
        dist.all_gather_object(handles, p.peer_export())
        p.peer_attach(world, rank, handles)
        f5p = p.sci_iterate(np.zeros(len(N)), 5)         # in-kernel exchange over peer memory
        f5p = p.sci_iterate(f5p, 0)
        errs_peer = np.max(np.abs(f5p[s] - f5[s]))
        # device-resident loops over the peer exchange (every pass of the adaptive iteration exchanges in-kernel,
        # the K x K Hessian goes through NCCL) and host-stepped loops must agree with the NCCL-only run above
        fk_p, r_p = p.solve_adaptive(np.zeros(len(N)), tol=1e-12, min_sc_iter=0)
        errs["adaptive_peer"] = np.max(np.abs(fk_p[s] - ref[s]))
        assert r_p["success"] and r["success"], (r, r_p)
        p.set_loop_mode("stepped")
        fk_s, r_s = p.solve_adaptive(np.zeros(len(N)), tol=1e-12, min_sc_iter=0)
        p.set_loop_mode("device")
        errs["adaptive_stepped_vs_device"] = np.max(np.abs(fk_s[s] - fk_p[s]))
        fs_p, rs_p = p.solve_sci(np.zeros(len(N)), tol=1e-11, maxiter=20000)
        errs["sci_loop_peer"] = np.max(np.abs(fs_p[s] - ref[s])) * 1e-2       # SCI stops ~1e-9 from the optimum
        gb = [None] * world
        dist.all_gather_object(gb, (fk_p.tobytes(), fs_p.tobytes(), fk.tobytes()))
        assert all(b == gb[0] for b in gb), "ranks disagree after the device-resident loops"
        fh = np.zeros(len(N))
        for _ in range(5):
            nxt = orc.self_consistent_update(u[s], N[s], fh[s])
            fh[s] = nxt - nxt[0]
        errs["sci_iterate"] = np.max(np.abs(f5[s] - fh[s]))
        errs["sci_iterate_peer_vs_nccl"] = errs_peer
        # every rank must hold bit-identical f after the in-kernel exchange (rank-ordered sum)
        g = [None] * world
        dist.all_gather_object(g, f5p.tobytes())
        assert all(b == g[0] for b in g), "ranks disagree after the peer exchange"
        bad = {k: v for k, v in errs.items() if not v < 1e-8}
        worst = max(worst, max(errs.values()))
        assert not bad, (name, rank, bad)
        p.close()
    # the mirror's sharded entry points (pymbar_b200.sharded): reference signatures, this rank's columns only
    from pymbar_b200 import mbar_solvers as ms
    from pymbar_b200 import sharded as sh

    for name in ("small_empty_state", "osc_50x100"):
        z = _cases.load(name)
        u, N_k = z["u_kn"], z["N_k"]
        Ntot = u.shape[1]
        lo, hi = Ntot * rank // world, Ntot * (rank + 1) // world
        sws = np.where(N_k != 0)[0]
        for proto_name, proto in (("default", ms.DEFAULT_SOLVER_PROTOCOL), ("robust", ms.ROBUST_SOLVER_PROTOCOL)):
            proto = tuple({k: (dict(v) if isinstance(v, dict) else v) for k, v in st.items()} for st in proto)
            f = sh.solve_mbar_for_all_states(np.ascontiguousarray(u[:, lo:hi]), N_k, np.zeros(len(N_k)), sws, proto)
            err = np.max(np.abs(f - z[f"fk_{proto_name}"]))
            assert err < 1e-8, (name, proto_name, err)
            worst = max(worst, err)
            gb = [None] * world
            dist.all_gather_object(gb, f.tobytes())
            assert all(b == gb[0] for b in gb), "sharded solve: ranks disagree"
        lw = sh.mbar_log_W_nk(np.ascontiguousarray(u[:, lo:hi]), N_k, z["fk_default"])
        ref_lw = orc.mbar_log_W_nk(u, N_k.astype(float), z["fk_default"])[lo:hi]
        assert np.max(np.abs(lw - ref_lw)) < 1e-9
        S, G = sh.weight_moments(np.ascontiguousarray(u[:, lo:hi]), N_k, z["fk_default"])
        W = orc.mbar_W_nk(u, N_k.astype(float), z["fk_default"])
        assert np.max(np.abs(G - W.T @ W)) < 1e-9 * np.max(W.T @ W) + 1e-14
    dist.barrier()
    if rank == 0:
        print(f"MG_OK world={world} worst_err={worst:.3e} (primitives, NCCL + in-kernel peer exchange, "
              f"device-resident and host-stepped loops; bit-identical f on all ranks)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
