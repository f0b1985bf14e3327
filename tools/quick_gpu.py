"""Quick GPU check: per-pass device times at a given size (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import strolle_b200
from strolle_b200 import scenes

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
name = sys.argv[3] if len(sys.argv) > 3 else "cornell"
e = strolle_b200.Engine(exact=bool(int(os.environ.get('ST_EXACT', '0'))))
for env, opt in (('ST_FUSED', 11), ('ST_FAST', 9), ('ST_PAIRED', 13), ('ST_TILED', 4)):   # development switches: ST_OPT_FUSED_PASSES / _SHADING_FAST_MATH / _WAVELET_PAIRED / _WAVELET_TILED
    if env in os.environ:
        e.set_option(opt, int(os.environ[env]))
sc = {"cornell": scenes.cornell, "dungeon": scenes.dungeon, "demo": scenes.demo_level}[name](w, h)
cam = scenes.apply(e, sc)
for f in range(12):
    e.tick(); e.render_camera(cam)
e.synchronize()
e.enable_timing(True)
e.pass_times(reset=True)
t = time.time()
N = 24
for f in range(N):
    e.tick(); e.render_camera(cam)
e.synchronize()
wall = time.time() - t
ms, launches = e.pass_times(reset=True)
tot = 0
for i, nme in enumerate(strolle_b200.PASS_NAMES):
    if launches[i]:
        print(f"{nme:36s} launches {launches[i]:4d}  total {ms[i]:9.3f} ms  avg {ms[i]/launches[i]*1000:9.1f} us")
        tot += ms[i]
print(f"scene {name} {w}x{h}: device total {tot/N:.3f} ms/frame, wall {wall/N*1000:.3f} ms/frame (timing mode)")
wms, wl = e.wavelet_times(reset=True)
print("K22 per iteration (us):", [round(float(m) / max(int(l), 1) * 1000, 1) for m, l in zip(wms, wl)])
out = e.read_buffer(cam, "output").reshape(h, w, 4)
print("mean", out[..., :3].mean(axis=(0, 1)), "nan", int(np.isnan(out).sum()))
