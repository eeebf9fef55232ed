"""Markdown table of the headline metrics of every kernel in an .ncu-rep (--set full).
usage: python scripts/ncu_table.py report.ncu-rep > profiles/<name>.md"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("smsp__inst_executed.sum", "warp instructions")]
idx = [(h.index(k), n) for k, n in cols if k in h]
print("| " + " | ".join(f"{n} [{u[i]}]" if u[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    vals = []
    for i, n in idx:
        v = r[i]
        if n == "kernel": v = "`" + v.split("(")[0].replace("void ", "")[:40] + "`"
        else:
            try: v = f"{float(v.replace(',', '')):.1f}" if "." in v else v
            except ValueError: pass
        vals.append(v)
    print("| " + " | ".join(vals) + " |")
