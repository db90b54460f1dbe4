"""Per-kernel summary of an ncu --set full report as CSV (the columns DESIGN.md quotes).
Usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx_summary.csv"""
import csv, subprocess, sys, io
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
cols = [h.index(w) for w in WANT if w in h]
w = csv.writer(sys.stdout)
w.writerow([h[i] for i in cols])
w.writerow([rows[1][i] for i in cols])
for r in rows[2:]:
    w.writerow([r[i] for i in cols])
