"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean time and share of the step per kernel,
next to the in-run CUDA-event share from a bench.py JSON line (kernel shares must agree; absolute times under ncu are
cold-cache and serialised).  Usage: python tools/ncu_launch_shares.py launches.csv bench.json > profiles/<name>.txt"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
t = collections.defaultdict(list)
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r[ik]).split("(")[0].replace("st::", "")
    name = re.sub(r"<\(bool\)(\d), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>", r"<fast=\1,S=\2,J=\3,\4x\5>", name)
    name = re.sub(r"<\(bool\)(\d), \(int\)(\d+), \(int\)(\d+)>", r"<fast=\1,\2x\3>", name)
    name = re.sub(r"<\(bool\)(\d)>", r"<fast=\1>", name)
    t[name].append(float(r[iv].replace(",", "")) / 1000.0)   # ns -> us
bench = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1]) if len(sys.argv) > 2 else None
total = sum(sum(v) for v in t.values())
KERNEL_PASS = {"k_prim_gbuffer": "prim_gbuffer", "k_di_sampling": "di_sampling", "k_di_temporal": "di_temporal_resampling", "k_di_spatial_pick": "di_spatial_resampling_pick",
               "k_di_spatial_sample": "di_spatial_resampling_sample", "k_di_resolving": "di_resolving", "k_gi_reprojection": "gi_reprojection", "k_gi_sampling_a": "gi_sampling_a",
               "k_gi_sampling_b": "gi_sampling_b", "k_gi_temporal": "gi_temporal_resampling", "k_gi_spatial_pick": "gi_spatial_resampling_pick",
               "k_gi_spatial_sample": "gi_spatial_resampling_sample", "k_gi_preview": "gi_preview_resampling", "k_gi_resolving": "gi_resolving",
               "k_frame_reprojection": "frame_reprojection", "k_denoise_reproject_pair": "frame_denoising_reproject", "k_composition": "frame_composition"}
print(f"# {len(sum(t.values(), []))} launches, {total / 1000.0:.2f} ms under ncu")
print(f"{'kernel':52s} {'launches':>8s} {'avg us':>8s} {'ncu share':>10s} {'in-run share':>13s}")
inrun_total = sum(bench["pass_ms_per_frame"].values()) if bench else 0.0
for name, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    share = ""
    base = name.split("<")[0]
    if bench and base in KERNEL_PASS:
        share = f"{100.0 * bench['pass_ms_per_frame'].get(KERNEL_PASS[base], 0.0) / inrun_total:12.1f}%"
    print(f"{name:52s} {len(v):8d} {sum(v) / len(v):8.1f} {100.0 * sum(v) / total:9.1f}% {share:>13s}")
if bench:
    grp = {"wavelet (all variants)": ("k_denoise_wavelet", "frame_denoising_wavelet"), "variance (all variants)": ("k_denoise_variance", "frame_denoising_estimate_variance"),
           "spatial trace (DI + GI)": ("k_spatial_trace", None)}
    for label, (prefix, p) in grp.items():
        s = sum(sum(v) for n, v in t.items() if n.startswith(prefix))
        inrun = bench["pass_ms_per_frame"].get(p, 0.0) if p else bench["pass_ms_per_frame"].get("di_spatial_resampling_trace", 0.0) + bench["pass_ms_per_frame"].get("gi_spatial_resampling_trace", 0.0)
        print(f"{label:52s} {'':8s} {'':8s} {100.0 * s / total:9.1f}% {100.0 * inrun / inrun_total:12.1f}%")
