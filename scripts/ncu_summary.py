"""Condense `ncu --set full` reports into one CSV (profiles/r1_ncu_full_quarter_scale.csv):
   python scripts/ncu_summary.py out.csv a.ncu-rep b.ncu-rep ..."""
import csv, io, subprocess, sys
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_membar"]
out, reps = sys.argv[1], sys.argv[2:]
rows = []
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    hdr = r[0]
    idx = [hdr.index(k) if k in hdr else -1 for k in KEEP]
    units = [r[1][i] if i >= 0 else "" for i in idx]   # units differ between reports: keep them in the cells
    rows += [[(x[i] + (" " + u if u and k not in KEEP[:3] else "")) if i >= 0 else "" for i, u, k in zip(idx, units, KEEP)]
             for x in r[2:]]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(KEEP); w.writerows(rows)
print(len(rows), "kernels ->", out)
