#!/usr/bin/env python
"""Summarise ncu artefacts brought back from the GPU box into small committed files under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
    python tools/summarize_ncu.py full gpurun_out/fused_c3_r1.ncu-rep profiles/fused_pass_c3_r1.md [K N]
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    agg = OrderedDict()
    for r in rows[hdr + 1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0][:90]
        t = float(r[iv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    unit = rows[hdr + 1][h.index("Metric Unit")]
    with open(dst, "w") as f:
        f.write(f"# ncu launch list — `{src}`\n\nEvery launch with its device time (cold-cache, serialised: "
                f"compare SHARES, not absolutes).  Unit: {unit}.\n\n| kernel | launches | total | share |\n|---|---:|---:|---:|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {t:,.0f} | {100 * t / tot:.1f}% |\n")
    print(open(dst).read())


def full(rep, dst, K=None, N=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    stalls = sorted(((float(v), h.split("issue_stalled_")[1].split("_per_")[0]) for h, (v, u) in d.items()
                     if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and v), reverse=True)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full — `{rep}`\n\nkernel: `{d.get('Kernel Name', ('?',))[0]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in KEYS:
            if k in d:
                f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
        f.write("\nWarp stall reasons (warps per issue-active cycle):\n\n| reason | ratio |\n|---|---:|\n")
        for v, h in stalls[:9]:
            f.write(f"| {h} | {v:.3f} |\n")
    if K and N:
        rd = float(d["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
        json.dump({"K": K, "N": N, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                   "algorithmic_bytes": 8.0 * K * N, "source": rep}, open(dst.replace(".md", ".json"), "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], *(int(float(a)) for a in sys.argv[4:6]))
