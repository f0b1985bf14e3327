"""Development aid: per-pass device time of tuning builds that cap registers per kernel (-DST_MINB_ALL=N).

    python tools/occupancy_tune.py [W H] [--json out.json]

Builds found as strolle_b200/_lib/libstrolle_b200_minb<N>.so (python -c "from strolle_b200 import build;
build.build(defines=['ST_MINB_ALL=8'], tag='minb8')") are each run in a child process (STROLLE_B200_LIB) on the
same Cornell workload; the table shows us/launch per pass and build, so that the best N per kernel can be
written into kernels.cu as ST_MINB_<KERNEL>.
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if "--child" in sys.argv:
    import strolle_b200
    from strolle_b200 import scenes
    w, h = int(sys.argv[2]), int(sys.argv[3])
    e = strolle_b200.Engine()
    cam = scenes.apply(e, scenes.dungeon(w, h) if "--dungeon" in sys.argv else scenes.cornell(w, h))
    for _ in range(12):
        e.tick(); e.render_camera(cam)
    e.synchronize(); e.enable_timing(True); e.pass_times(reset=True)
    N = 24
    for _ in range(N):
        e.tick(); e.render_camera(cam)
    e.synchronize()
    ms, launches = e.pass_times(reset=True)
    out = {n: float(ms[i]) / launches[i] * 1000.0 for i, n in enumerate(strolle_b200.PASS_NAMES) if launches[i]}
    out["_frame_us"] = float(sum(ms)) / N * 1000.0
    print("RESULT " + json.dumps(out))
    sys.exit(0)

args = [a for a in sys.argv[1:] if not a.startswith("--")]
w, h = (args[0], args[1]) if len(args) >= 2 else ("1920", "1080")
libs = {"base": os.path.join(ROOT, "strolle_b200", "_lib", "libstrolle_b200.so")}
for p in sorted(glob.glob(os.path.join(ROOT, "strolle_b200", "_lib", "libstrolle_b200_*.so"))):
    libs[os.path.basename(p)[len("libstrolle_b200_"):-3]] = p
table = {}
for tag, lib in libs.items():
    env = dict(os.environ, STROLLE_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", w, h] + (["--dungeon"] if "--dungeon" in sys.argv else []), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        print(tag, "FAILED", r.stdout[-400:]); continue
    table[tag] = json.loads(line[0][7:])
tags = list(table)
print(f"{'pass':36s} " + " ".join(f"{t:>9s}" for t in tags) + "   best")
for name in table[tags[0]]:
    vals = [table[t].get(name, float('nan')) for t in tags]
    best = tags[min(range(len(vals)), key=lambda i: vals[i])]
    print(f"{name:36s} " + " ".join(f"{v:9.1f}" for v in vals) + f"   {best}")
if "--json" in sys.argv:
    json.dump(table, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
