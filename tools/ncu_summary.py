#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md / bench.py cite.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx_ncu_summary.txt
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --roofline R S N profiles/rNN_xxx_ncu_summary.txt
        additionally writes profiles/roofline_kernel.json (kernel name + DRAM bytes per launch of the first kernel in the
        report, captured at the shape R rays x S samples x N instances), which bench.py reports as roofline.traffic
"""
import json
import os
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {r[idx['Kernel Name']][:110]}  grid={r[idx.get('Grid Size', 0)]} block={r[idx.get('Block Size', 0)]}")
        for w in WANT:
            if w in idx:
                print(f"   {w:88s} {r[idx[w]]:>16s} {units[idx[w]]}")


def roofline(path, rays, samples, instances, source):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, first = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}

    def in_bytes(name):
        value, unit = float(first[idx[name]].replace(",", "")), units[idx[name]].lower()
        return value * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]

    out = dict(kernel=re.sub(r"\(.*", "", first[idx["Kernel Name"]]).replace("void ", ""), shape=[rays, samples, instances],
               traffic_bytes=in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum"),
               dram_bytes_read=in_bytes("dram__bytes_read.sum"), dram_bytes_write=in_bytes("dram__bytes_write.sum"),
               ncu_duration_us=float(first[idx["gpu__time_duration.sum"]].replace(",", "")), source=source)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "roofline_kernel.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out), file=sys.stderr)


if __name__ == "__main__":
    import re
    main(sys.argv[1])
    if len(sys.argv) > 2 and sys.argv[2] == "--roofline":
        roofline(sys.argv[1], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
