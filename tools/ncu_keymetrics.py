"""Key metrics per kernel launch from an ncu report (raw page)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
seen = set()
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0][-40:]
    if name in seen and "--all" not in sys.argv:
        continue
    seen.add(name)
    print("---", name)
    for k in keys:
        if k in h:
            print(f"   {k:80s} {r[h.index(k)]:>16s} {u[h.index(k)]}")
