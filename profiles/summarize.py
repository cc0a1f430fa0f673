"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv           > profiles/rNN_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_fft_job.ncu-rep   > profiles/rNN_fft_job_kernel.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) > vi:
            agg[r[ki][:90]][0] += 1
            agg[r[ki][:90]][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("# gpu__time_duration.sum per kernel (cold-cache, serialised: compare SHARES)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:10.3f} ms {c:5d} launches {100 * t / tot:6.2f}%  {n}")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    for D in rows[2:]:
        print("# kernel:", D[H.index("Kernel Name")][:100])
        for k in KEYS:
            if k in H:
                i = H.index(k)
                print(f"{k:80s} {D[i]:>22s} {U[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    H, D = rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(H)}
    stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
    tot, ops, samples, inst = collections.Counter(), collections.Counter(), 0, 0
    for r in D:
        if len(r) < len(H):
            continue
        samples += int(r[ci["# Samples"]] or 0)
        ie = int(r[ci["Instructions Executed"]] or 0)
        inst += ie
        for st in stalls:
            tot[st] += int(r[ci[st]] or 0)
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
        ops[m.group(2).split(".")[0] if m else "?"] += ie
    print(f"# warp stall sampling: {samples} samples over {len(D)} SASS instructions, {inst} warp-instructions executed")
    for k, v in tot.most_common(8):
        print(f"{k:30s} {100 * v / max(samples, 1):6.2f}%")
    print("# opcode mix (share of executed warp-instructions)")
    for k, v in ops.most_common(16):
        print(f"{k:12s} {100 * v / max(inst, 1):6.2f}%")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
