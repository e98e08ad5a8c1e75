#!/usr/bin/env python
"""bench.py — MBAR self-consistent iteration throughput on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric / configs[2] shape): synthetic harmonic-oscillator u_kn, K = 256
states, N = 1e7 samples PER GPU (20.48 GB of fp64 in HBM, far larger than the 126 MB L2), equal N_k,
Philox-keyed generation on device.  A step = one full self-consistent iteration (Eq. C3):
one fused streaming pass over u_kn + (N>1) one NCCL all-reduce of the K+2 partials + the K-vector
update, device resident, no host round trip between steps.  Samples shard over ranks (weak
scaling: each rank owns 1e7 samples of a 1e7*N-sample problem), no other data-path collective.

Reported: value = K*N_total*iterations / second (entries/s, whole job), `iter_per_s`, achieved HBM
GB/s, `roofline` for the fused pass kernel (algorithmic bytes 8*K*N_local per launch over its
CUDA-event launch duration, against MEASURED_PEAKS.json), `e2e` through the reference-facing call
`pymbar_b200.mbar_solvers.self_consistent_update(u_kn_host, N_k, f_k)` with the u_kn upload inside
every timed step, and `cpu_baseline` = the numpy oracle port of the reference timed on this host.

`--impl reference`: the reference's own CPU algorithm (oracle numpy port — the reference is pure
Python and /root/reference does not exist on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_STATES = 256
N_PER_GPU = 10_000_000
METRIC = "MBAR self-consistent iteration throughput at K=256, N=1e7 samples per GPU (u_kn entries/s; iter/s and HBM GB/s alongside)"
UNIT = "entries/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_params(K):
    O_k = np.linspace(1.0, 5.0, K)      # utils_for_testing.py:64-66 spacing
    k_k = np.linspace(1.0, 3.0, K)
    return O_k, k_k


def global_N_k(K, N_total):
    N_k = np.full(K, N_total // K, dtype=np.float64)
    N_k[-1] += N_total - N_k.sum()
    return N_k


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(K, N_local):
    """dram read+write bytes per launch of the fused kernel from the committed ncu capture, scaled
    to this launch's size when the capture was taken at a smaller N of the same K."""
    path = os.path.join(ROOT, "profiles", "fused_pass_c3_r1.json")
    try:
        t = json.load(open(path))
        if t["K"] == K:
            return float(t["dram_bytes_per_launch"]) * (N_local / t["N"])
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle = numpy port of the reference; test infrastructure used here only as the baseline)
# ---------------------------------------------------------------------------------------------
def cpu_sample(K, n_sample, seed=0):
    from oracle import testsystems as ots  # noqa: F401  (same functional form, numpy RNG on the host)

    O_k, k_k = workload_params(K)
    rng = np.random.default_rng(seed)
    per = max(1, n_sample // K)
    x = np.concatenate([rng.normal(O_k[k], k_k[k] ** -0.5, per) for k in range(K)])
    u_kn = 0.5 * k_k[:, None] * (x[None, :] - O_k[:, None]) ** 2
    return u_kn, np.full(K, float(per))


def time_cpu_reference(K, budget_s, steps, warmup):
    """Oracle self_consistent_update on a bounded sample sized to `budget_s` seconds in total."""
    from oracle import mbar_oracle as orc

    u, N_k = cpu_sample(K, 20 * K)
    f = np.zeros(K)
    t0 = time.perf_counter()
    orc.self_consistent_update(u, N_k, f)
    rate = u.size / max(time.perf_counter() - t0, 1e-6)              # entries/s, rough
    entries = rate * budget_s / max(1, steps + warmup)
    n_sample = int(min(max(entries / K, 40 * K), 2_000_000))
    u, N_k = cpu_sample(K, n_sample)
    for _ in range(warmup):
        f = orc.self_consistent_update(u, N_k, f)
        f -= f[0]
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        f = orc.self_consistent_update(u, N_k, f)
        f -= f[0]
        times.append(time.perf_counter() - t0)
    per_step = float(np.mean(times))
    return u.size / per_step, per_step, u.shape[1]


def time_c_port(K, n_sample=400_000):
    """The C/pthreads restatement of the same two sweeps with every host thread: context beside the
    single-threaded numpy port (which is what the reference's own implementation is)."""
    try:
        from oracle import c_oracle

        u, N_k = cpu_sample(K, n_sample)
        f = np.zeros(K)
        c_oracle.self_consistent_update(u, N_k, f)
        t0 = time.perf_counter()
        for _ in range(3):
            c_oracle.self_consistent_update(u, N_k, f)
        dt = (time.perf_counter() - t0) / 3
        return {"value": u.size / dt, "unit": UNIT, "threads": c_oracle.threads(),
                "what": "oracle/mbar_oracle.c (pthreads), same arithmetic, all host threads"}
    except Exception as exc:  # pragma: no cover
        return {"unavailable": str(exc)[:200]}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    K = K_STATES
    value, per_step, n_sample = time_cpu_reference(K, budget_s=90.0, steps=args.steps, warmup=args.warmup)
    try:
        from threadpoolctl import threadpool_info
        blas = max([t.get("num_threads", 1) for t in threadpool_info()] or [1])
    except Exception:
        blas = None
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "iter_per_s_at_full_size": value / (K * N_PER_GPU),
        "config": {"workload": "MBAR self-consistent iteration, K=256, N=1e7 per GPU (reference timed on a bounded sample)",
                   "K": K, "N_sample": n_sample, "flush": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "host_cpus": os.cpu_count(), "blas_threads": blas,
                         "c_port_all_threads": time_c_port(K),
                         "sample": f"self_consistent_update (numpy oracle port of mbar_solvers.py:231-242, "
                                   f"scipy.special.logsumexp x2, single-threaded like the reference) on K={K}, "
                                   f"N={n_sample} of the same harmonic family; throughput is flat in N (BASELINE.md §2)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pymbar_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from pymbar_b200 import DeviceProblem, PinnedArray
    from pymbar_b200 import mbar_solvers as ms

    K, N_local = K_STATES, args.n_per_gpu
    N_total = N_local * world
    N_k = global_N_k(K, N_total)
    O_k, k_k = workload_params(K)

    prob = DeviceProblem(None, N_k, device=local, N_local=N_local)
    prob.synthesize(O_k, k_k, seed=args.seed, n_offset=rank * N_local, N_global=N_total)
    if distributed:
        uid = [DeviceProblem.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        prob.comm_init(world, rank, uid[0])
        if not os.environ.get("MBAR_B200_NO_PEER"):
            # in-kernel exchange of the partial sums over peer memory (one kernel per iteration)
            handles = [None] * world
            dist.all_gather_object(handles, prob.peer_export())
            prob.peer_attach(world, rank, handles)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    f = np.zeros(K)
    f = prob.sci_iterate(f, args.warmup)                   # W untimed steps (also warms NCCL)
    barrier()
    c0 = prob.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    f = prob.sci_iterate(f, args.steps)                    # exactly K timed steps, device resident
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    loop = prob.last_loop_ms()                             # CUDA events on the launching stream
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    c1 = prob.counters()

    dev_ms = torch.tensor([loop["total_ms"], loop["kernel_ms_sum"]], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)      # max over ranks, device-timed
    total_ms, kernel_ms = dev_ms.tolist()
    ms_per_step = total_ms / args.steps
    value = K * N_total * args.steps / (total_ms * 1e-3)
    kern_ms_per_launch = kernel_ms / args.steps
    peak, peak_src = measured_peak()
    achieved = 8.0 * K * N_local / (kern_ms_per_launch * 1e-3) / 1e9

    # a real solve on the same data (not part of the timed region): adaptive Newton/self-consistent
    # solver of mbar_solvers.py:510-667 from f = 0 to tol 1e-12 (C3: "Newton-Raphson with K x K Hessian")
    barrier()
    t0 = time.perf_counter()
    f_solved, solve_info = prob.solve_adaptive(np.zeros(K), tol=1e-12, maxiter=100, min_sc_iter=0)
    solve_wall = time.perf_counter() - t0
    adaptive_solve = {k: solve_info[k] for k in ("success", "iterations", "nr_iterations", "sci_iterations",
                                                 "passes", "hessian_passes", "gnorm", "device_ms")}
    adaptive_solve["wall_s"] = solve_wall

    # parity spot check of the timed state against the CPU oracle on a slice (not timed)
    parity = None
    if rank == 0:
        from oracle import mbar_oracle as orc

        sl = prob.download(0, 4096)
        Sdev = None
        try:
            p2 = DeviceProblem(sl, N_k, device=local)
            Sdev, _, _ = p2.streaming_pass(f)
            S_ref, _ = orc.single_pass_sums(sl, N_k, f)
            parity = float(np.max(np.abs(Sdev - S_ref) / S_ref))
            p2.close()
        except Exception as exc:  # pragma: no cover
            parity = f"failed: {exc}"

    # ---- e2e: reference-facing call with HOST buffers, upload inside every step --------------------
    e2e = None
    if args.e2e_steps > 0:
        os.environ["PYMBAR_B200_CACHE"] = "0"              # every call uploads u_kn (no residency)
        os.environ["PYMBAR_B200_DEVICE"] = str(local)
        ms._DEVICE = local
        pin = PinnedArray((K, N_local))
        prob.download(0, N_local, out=pin.array)           # setup: host copy of this rank's shard
        h2d0 = 8 * K * N_local + 8 * K                      # u_kn + c vector per step
        fh = np.zeros(K)
        # free the resident copy so two 20 GB problems never coexist with staging buffers
        Nk_local_view = N_k
        times = []
        for i in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            out = ms.self_consistent_update(pin.array, Nk_local_view, fh)
            dt = time.perf_counter() - t0
            if i > 0:
                times.append(dt)
            fh = out - out[0]
        t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": K * N_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d0,
               "d2h_bytes_per_step": 8 * (2 * K + 2), "s_per_step": e2e_s, "steps": args.e2e_steps,
               "call": "pymbar_b200.mbar_solvers.self_consistent_update(u_kn_host[pinned], N_k, f_k), "
                       "PYMBAR_B200_CACHE=0 (create + upload + pass + destroy per call)",
               "note": "N>1: each rank times its own shard's call (no cross-rank reduction in this leg)"
                       if distributed else "single GPU"}
        pin.free()

    if rank == 0:
        cpu_val, cpu_step, n_sample = time_cpu_reference(K, budget_s=args.cpu_budget, steps=2, warmup=1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "iter_per_s": 1e3 / ms_per_step,
            "hbm_gbs_per_gpu": 8.0 * K * N_local / (ms_per_step * 1e-3) / 1e9,
            "config": {"workload": "MBAR self-consistent iteration (Eq. C3), synthetic harmonic u_kn, "
                                   "K=256, N=1e7 per GPU, fp64, device-resident, samples sharded over GPUs",
                       "K": K, "N_per_gpu": N_local, "N_total": N_total, "seed": args.seed,
                       "parallelism": f"sample-sharded x{world}; per iteration the {K + 2} partial sums are exchanged "
                                      + ("inside the pass kernel over NVLink peer memory (rank-ordered sum)"
                                         if (distributed and not os.environ.get("MBAR_B200_NO_PEER")) else
                                         "by one NCCL all-reduce" if distributed else "n/a (1 GPU)"),
                       "l2": "inputs (20.48 GB per GPU) far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the launching stream around the K-step loop, max over ranks",
                       "wall_s_rank0": wall},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(K, N_local),
                         "kernel": "pass_fused_kernel<R=32, FULL, 8 warps, batch 8, MODE 3 (LDS table + multiplicative constant), CL=1>",
                         "algorithmic_bytes_per_launch": 8 * K * N_local,
                         "launch_ms": kern_ms_per_launch, "peak_source": peak_src,
                         "how": "per-launch cudaEvent pairs recorded around the kernel inside the timed loop"},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
                             "host_cpus": os.cpu_count(),
                             "c_port_all_threads": time_c_port(K),
                             "sample": f"numpy oracle self_consistent_update on K={K}, N={n_sample} of the same "
                                       f"family ({cpu_step:.2f} s/step); reference is single-threaded numpy"},
            "e2e": e2e,
            "adaptive_solve": adaptive_solve,
            "gpu_launches": c1["launches"] - c0["launches"],
            "clocks": clocks,
            "parity_S_rel_err_vs_oracle_slice": parity,
        }
        print(json.dumps(line), flush=True)
    prob.close()
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = env_int("WORLD_SIZE", 1)
    if args.gpus > 1 and world == 1 and args.impl == "ours":
        # convenience: relaunch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
