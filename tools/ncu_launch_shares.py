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
    # Product-default frames only (the bench also renders strict-tier frames, whose kernels carry other names): per-frame time of
    # every pass under ncu = mean launch time x launches per frame, next to the in-run CUDA-event time of the same pass.
    frames = len(t.get("stf::k_di_sample_temporal", [])) or 1
    PASS_OF = [("prim_gbuffer", ["k_prim_gbuffer"], 1), ("di_temporal_resampling", ["stf::k_di_sample_temporal"], None), ("di_spatial_resampling_pick", ["stf::k_di_spatial_fused"], None),
               ("di_resolving", ["stf::k_di_resolving"], None), ("gi_reprojection", ["stf::k_gi_reprojection"], None), ("gi_sampling_b", ["stf::k_gi_sampling_fused"], None),
               ("gi_temporal_resampling", ["stf::k_gi_temporal"], None), ("gi_spatial_resampling_pick", ["stf::k_gi_spatial_fused"], None),
               ("gi_preview_resampling", ["stf::k_gi_preview", "stf::k_gi_preview_resolve"], None), ("frame_denoising_reproject", ["k_denoise_reproject_pair"], 1),
               ("frame_denoising_estimate_variance", [n for n in t if n.startswith("k_denoise_variance") and "<1" in n], None),
               ("frame_denoising_wavelet", [n for n in t if n.startswith("k_denoise_wavelet") and "<1" in n], None), ("frame_composition", ["k_composition"], 1)]
    per = {}
    for pname, kernels, fixed in PASS_OF:
        us = 0.0
        for k in kernels:
            v = t.get(k, [])
            if v:
                us += (sum(v) / len(v)) * (fixed if fixed else len(v) / frames)
        per[pname] = us
    tot_ncu = sum(per.values()); tot_run = sum(bench["pass_ms_per_frame"].get(pn, 0.0) for pn in per) * 1000.0
    print(f"\n# product-default frames ({frames} captured): per-frame time of each pass, ncu (cold caches, serialised) vs in-run CUDA events")
    print(f"{'pass':40s} {'ncu us':>9s} {'ncu share':>10s} {'in-run us':>10s} {'in-run share':>13s}")
    for pname, us in sorted(per.items(), key=lambda kv: -kv[1]):
        run = bench["pass_ms_per_frame"].get(pname, 0.0) * 1000.0
        print(f"{pname:40s} {us:9.1f} {100.0 * us / tot_ncu:9.1f}% {run:10.1f} {100.0 * run / tot_run:12.1f}%")
    print(f"{'total':40s} {tot_ncu:9.1f} {'':10s} {tot_run:10.1f}")
