"""Summarise an `ncu --set full` report (.ncu-rep) into the JSON kept under profiles/: one object per captured launch
with the metrics DESIGN.md / bench.py quote.  usage: ncu_summary.py <report.ncu-rep> > profiles/<tag>_ncu_full_summary.json"""
import csv, json, subprocess, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        # what bounds the rank-path count kernel: requests / sectors / wavefronts through L1TEX and the crossbar, and why warps wait
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
head, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = {"Kernel Name": r[head.index("Kernel Name")]}
    u = {}
    for k in KEEP:
        if k in head:
            d[k] = r[head.index(k)]
            u[k] = units[head.index(k)]
    d["units"] = u
    res.append(d)
print(json.dumps(res, indent=1))
