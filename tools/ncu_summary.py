#!/usr/bin/env python
"""Summarises an .ncu-rep (read offline with `ncu -i`): raw metrics per captured launch and the hottest
source lines (instructions executed) per kernel, across inlined headers.  Usage:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel_regex] > profiles/<name>.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
kre = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(".")
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def run(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


rows = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}")
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if not kre.search(name):
        continue
    print(f"\n== launch: {name[:100]}")
    for m in RAW:
        if m in idx:
            print(f"   {m:72s} {r[idx[m]]} {units[idx[m]]}")

src = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
sections, cur = [], None
for r in src:
    if len(r) >= 2 and r[0] == "File Path":
        cur = {"file": r[1], "rows": [], "func": ""}
        sections.append(cur)
    elif len(r) >= 2 and r[0] == "Function Name" and cur is not None:
        cur["func"] = r[1]
    elif r and r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
per_kernel = collections.defaultdict(list)
for s in sections:
    per_kernel[s["func"]].append(s)
for func, secs in per_kernel.items():
    if not kre.search(func):
        continue
    lines, ops, total = [], collections.Counter(), 0
    for s in secs:
        h = s["hdr"]
        ii = h.index("Instructions Executed")
        for r in s["rows"]:
            if not r[ii].isdigit():
                continue
            n = int(r[ii])
            if r[2].strip().startswith("0x"):      # SASS row (has an address)
                t = r[3].split()
                op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "?")).split(".")[0]
                ops[op] += n
                total += n
            elif r[0].strip().isdigit():
                lines.append((n, s["file"].split("/")[-1], r[0], r[1].strip()))
    if not total:
        continue
    print(f"\n== source: {func[:100]}\n   warp instructions over the captured launches: {total}")
    print("   opcode mix: " + ", ".join(f"{o} {100 * n / total:.1f}%" for o, n in ops.most_common(16)))
    for n, f, ln, text in sorted(lines, reverse=True)[:28]:
        print(f"   {100 * n / total:5.1f}%  {f}:{ln:>4s}  {text[:110]}")
