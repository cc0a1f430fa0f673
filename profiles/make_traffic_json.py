"""profiles/r02_traffic.json from one `ncu --set full` capture of ONE launch of the bench step.

    python profiles/make_traffic_json.py gpurun_out/<capture>.ncu-rep <sets_per_launch> "<command that was profiled>"

bench.py quotes these numbers in `roofline` ONLY while `kernel_source_hash` equals the hash of the CUDA sources
the library is built from (pde_b200/csrc/build.py:source_hash) -- a capture of another kernel revision reads null.
Executed FP64 work: thread-level DFMA/DMUL/DADD counts of the capture (2 flops per DFMA, 1 per DMUL/DADD);
shared-memory efficiency: wavefronts actually issued by LDS/STS against the ideal count (settles whether the
`l1tex__data_bank_conflicts_pipe_lsu_mem_shared` counter means real conflicts or the 128-bit multi-wavefront
accounting).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pde_b200.csrc.build import source_hash  # noqa: E402

path, sets, cmd = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, D = rows[0], rows[2]


def val(name, default=None):
    if name not in H:
        return default
    return float(D[H.index(name)].replace(",", ""))


cycles = val("sm__cycles_elapsed.avg")
n_sm = 148
per_cycle = lambda op: val(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed", 0.0)  # noqa: E731
dfma, dmul, dadd = (per_cycle(o) * cycles for o in ("dfma", "dmul", "dadd"))
out = {
    "kernel": D[H.index("Kernel Name")][:60],
    "kernel_source_hash": source_hash(),
    "sets_per_launch": sets, "maturities": 32, "strikes": 50,
    "dram_bytes_read": val("dram__bytes_read.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[rows[1][H.index("dram__bytes_read.sum")]],
    "dram_bytes_write": val("dram__bytes_write.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[rows[1][H.index("dram__bytes_write.sum")]],
    "gpu_time_ms_under_ncu": val("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(rows[1][H.index("gpu__time_duration.sum")], 1),
    "fp64_pipe_active_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "thread_inst_dfma": dfma, "thread_inst_dmul": dmul, "thread_inst_dadd": dadd,
    "executed_flop_per_launch": 2 * dfma + dmul + dadd,
    "warp_inst_executed": val("smsp__inst_executed.sum"),
    "local_loads": val("sass__inst_executed_local_loads"), "local_stores": val("sass__inst_executed_local_stores"),
    "smem_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "smem_wavefronts_ideal_ld": val("smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ld_ideal.sum") if False else None,
    "smem_bank_conflict_counter": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "smem_ld_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
    "smem_st_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"),
    "smem_ld_requests": val("smsp__inst_executed_op_shared_ld.sum"),
    "smem_st_requests": val("smsp__inst_executed_op_shared_st.sum"),
    "source": cmd,
}
out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
