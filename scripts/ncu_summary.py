#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into the JSON summaries kept under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.json"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    res = []
    for r in rows[2:]:
        d = {k: (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip() for k in WANT if k in hdr}
        top = sorted(stall, key=lambda h: -float(r[hdr.index(h)] or 0))[:6]
        d["top_stalls_per_issue"] = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): r[hdr.index(h)] for h in top}
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["Kernel Name"][:60], d.get("gpu__time_duration.sum"), "fp64", d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
              "tensor", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
