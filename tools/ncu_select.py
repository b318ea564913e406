#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> the columns the profiles/ summaries quote, one row per captured launch.
    python tools/ncu_select.py gpurun_out/X.ncu-rep > profiles/rNN_ncu_full_selected.csv"""
import csv, subprocess, sys
COLS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
idx = [(c, hdr.index(c)) for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + [f"{c} [{units[i]}]" for c, i in idx])
for r in data:
    name = r[ki].split("(")[0].replace("dod::<unnamed>::", "").replace("void ", "")
    w.writerow([name] + [r[i] for _, i in idx])
