#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the few numbers the design discussion uses.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.txt> "<title>" """
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
with open(out, "w") as f:
    f.write(title + "\n" + "source: ncu --set full --clock-control none, read with `ncu -i <rep> --page raw --csv`\n\n")
    for k, vals in enumerate(rows[2:]):
        f.write("launch %d\n" % k)
        for h, u, v in zip(hdr, units, vals):
            if h in want:
                f.write("  %-88s %-16s %s\n" % (h, u, v))
print(open(out).read()[:600])
