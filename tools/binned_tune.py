"""Development aid: the fused launches with / without their direction-sorted CTA tracer (ST_OPT_BINNED_TRACE bit mask), per-pass us."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strolle_b200
from strolle_b200 import scenes
from strolle_b200.engine import OPT_BINNED_TRACE

for name in sys.argv[1:] or ["cornell", "demo"]:
    e = strolle_b200.Engine()
    cam = scenes.apply(e, {"cornell": scenes.cornell, "demo": scenes.demo_level}[name](1920, 1080))
    for _ in range(12):
        e.tick(); e.render_camera(cam)
    names = list(strolle_b200.PASS_NAMES)
    cols = ["gi_sampling_b", "gi_spatial_resampling_pick", "di_spatial_resampling_pick"]
    print(f"[{name}] us per launch: " + ", ".join(cols) + ", frame us")
    for mask in (0, 1, 2, 4, 7):
        e.set_option(OPT_BINNED_TRACE, mask)
        for _ in range(6):
            e.tick(); e.render_camera(cam)
        e.synchronize(); e.enable_timing(True); e.pass_times(reset=True)
        for _ in range(24):
            e.tick(); e.render_camera(cam)
        e.synchronize()
        ms, n = e.pass_times(reset=True)
        e.enable_timing(False)
        row = [float(ms[names.index(c)]) / max(int(n[names.index(c)]), 1) * 1000.0 for c in cols]
        print(f"  mask {mask}: " + "  ".join(f"{v:7.1f}" for v in row) + f"   {float(ms.sum()) / 24 * 1000.0:8.1f}")
    e.close()
