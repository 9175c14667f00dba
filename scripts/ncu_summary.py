"""Summarise an .ncu-rep (one row per captured launch) into the handful of numbers DESIGN.md / bench.py cite."""
import csv, subprocess, sys, io, json
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cross_attention_renderer_b200 import _lib
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum", "l1tex__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_op_hmma.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
res = []
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = {}
    for k in keys:
        if k in d:
            out[k] = f"{d[k][0]} {d[k][1]}".strip()
    res.append(out)
# build_id ties the capture to the kernel sources it was taken from (bench.py only quotes traffic of the build it times)
print(json.dumps({"build_id": _lib.build_id(), "report": os.path.basename(rep), "launches": res}, indent=1))
