"""Turn .ncu-rep files into the small CSV summaries kept under profiles/ (one column per kernel launch).
usage: python tools/ncu_summary.py out.csv rep1.ncu-rep [rep2 ...]   (needs `ncu` on PATH to read the reports)"""
import csv, io, subprocess, sys
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"
cols = []
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        col = {"report": rep.split("/")[-1]}
        for k in KEYS:
            if k in d:
                col[k] = d[k] + (" " + u[k] if u.get(k) and k not in ("Kernel Name", "Grid Size", "Block Size") else "")
        for k in hdr:
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
                col["stall_" + k[len(STALL):-len("_per_issue_active.ratio")]] = d[k]
        cols.append(col)
keys = []
for c in cols:
    for k in c:
        if k not in keys:
            keys.append(k)
with open(sys.argv[1], "w", newline="") as f:
    w = csv.writer(f)
    for k in keys:
        w.writerow([k] + [c.get(k, "") for c in cols])
print("wrote", sys.argv[1], len(cols), "launches")
