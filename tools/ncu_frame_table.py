#!/usr/bin/env python
"""One line per captured launch of an .ncu-rep (`ncu --set full`): time, warp instructions, issue-active,
occupancy, registers, DRAM bytes, cache hit rates.  Usage: python tools/ncu_frame_table.py rep.ncu-rep"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "Minst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "rdMB"),
        ("dram__bytes_write.sum", "wrMB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"), ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%")]
print(f"{'kernel':34s}" + "".join(f"{c[1]:>9s}" for c in COLS))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0][:33]
    vals = []
    for m, _ in COLS:
        v = r[idx[m]].replace(",", "")
        try:
            x = float(v)
            u = units[idx[m]]
            if m.startswith("smsp__inst_executed"):
                x /= 1e6
            if "bytes" in m:
                x = x / 1e6 if u == "byte" else x * {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            vals.append(f"{x:9.1f}")
        except ValueError:
            vals.append(f"{v:>9s}")
    print(f"{name:34s}" + "".join(vals))
