"""Summarise an ncu report (raw page) into the handful of metrics we track; used to fill profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "sm__throughput.avg.pct", "warp_issue_stalled.*per_warp_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_atom", "lts__t_sectors_op_red", "sm__pipe_fp64", "sm__inst_executed_pipe_fp64", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
import re
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if any(re.search(p, h) for p in pats):
            print("%-80s %-10s %s" % (h, units[i], r[i]))
