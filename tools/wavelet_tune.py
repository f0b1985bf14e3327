"""Development aid: K22 (à-trous wavelet) gather kernel vs the tile-staged (TMA) kernel, per iteration and tile shape.

    python tools/wavelet_tune.py [W H] [--json out.json]

For every iteration i (stride 2^i) and tile shape c, only that iteration is switched to the tile-staged kernel and
the per-pass CUDA-event time of the five wavelet launches is compared with the all-gather baseline; the difference
is that iteration's gain.  Also times K20 split vs fused.  Prints one table; optionally writes JSON.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strolle_b200
from strolle_b200 import scenes
from strolle_b200.engine import OPT_WAVELET_TILED, OPT_WAVELET_TILE_CFG, OPT_FUSE_REPROJECT, OPT_VARIANCE_TILED, STAT_WAVELET_TILED_ERRORS

args = [a for a in sys.argv[1:] if not a.startswith("--")]
w, h = (int(args[0]), int(args[1])) if len(args) >= 2 else (1920, 1080)
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
FRAMES = 12
SHAPES = ["32x8", "32x16", "64x4", "64x8"]

e = strolle_b200.Engine()
cam = scenes.apply(e, scenes.cornell(w, h))
names = list(strolle_b200.PASS_NAMES)
WAVELET, REPROJECT = names.index("frame_denoising_wavelet"), names.index("frame_denoising_reproject")
VARIANCE = names.index("frame_denoising_estimate_variance")


LAST = {}


def measure(mask, cfg, fuse=0, var_tiled=0):
    e.set_option(OPT_WAVELET_TILED, mask); e.set_option(OPT_WAVELET_TILE_CFG, cfg); e.set_option(OPT_FUSE_REPROJECT, fuse); e.set_option(OPT_VARIANCE_TILED, var_tiled)
    for _ in range(6):
        e.tick(); e.render_camera(cam)
    e.synchronize(); e.enable_timing(True); e.pass_times(reset=True)
    for _ in range(FRAMES):
        e.tick(); e.render_camera(cam)
    e.synchronize()
    ms, launches = e.pass_times(reset=True)
    e.enable_timing(False)
    LAST["variance_us"] = float(ms[VARIANCE]) / FRAMES * 1000.0
    return float(ms[WAVELET]) / FRAMES * 1000.0, float(ms[REPROJECT]) / FRAMES * 1000.0, float(sum(ms)) / FRAMES * 1000.0


for _ in range(12):
    e.tick(); e.render_camera(cam)
base_w, base_r, base_f = measure(0, 0)
base_w2, _, _ = measure(0, 0)
print(f"{w}x{h}: gather baseline, 5 wavelet launches: {base_w:.1f} us/frame (repeat {base_w2:.1f}); K20 x2: {base_r:.1f} us; frame (timed mode) {base_f:.1f} us")
results = {"size": [w, h], "baseline_wavelet_us": base_w, "baseline_reproject_us": base_r, "gain_us": {}}
best = {}
for i in range(5):
    row = []
    for c in range(len(SHAPES)):
        t, _, _ = measure(1 << i, c << (4 * i))
        gain = base_w - t
        row.append(gain)
        results["gain_us"][f"stride{1 << i}:{SHAPES[c]}"] = gain
    best[i] = max(range(len(row)), key=lambda c: row[c])
    print(f"  stride {1 << i:2d}: gain vs gather (us/launch) " + "  ".join(f"{SHAPES[c]} {row[c]:+6.1f}" for c in range(len(row))))
mask = 0; cfg = 0
for i in range(5):
    if results["gain_us"][f"stride{1 << i}:{SHAPES[best[i]]}"] > 1.0:
        mask |= 1 << i; cfg |= best[i] << (4 * i)
t, _, f_all = measure(mask, cfg)
_, r_fused, f_fused = measure(mask, cfg, fuse=1)
print(f"best mask {mask} cfg 0x{cfg:05x}: wavelet {t:.1f} us/frame (gather {base_w:.1f}); K20 fused {r_fused:.1f} us (split {base_r:.1f}); frame {f_fused:.1f} us (was {base_f:.1f})")
v_gather = LAST["variance_us"]
_, _, f_var = measure(mask, cfg, fuse=1, var_tiled=1)
print(f"K21 variance: gather {v_gather:.1f} us, tile-staged {LAST['variance_us']:.1f} us; frame with it {f_var:.1f} us")
results.update({"variance_gather_us": v_gather, "variance_tiled_us": LAST["variance_us"]})
print("tile errors:", e.get_stat(STAT_WAVELET_TILED_ERRORS))
results.update({"best_mask": mask, "best_cfg": cfg, "best_wavelet_us": t, "fused_reproject_us": r_fused, "frame_us": f_fused, "baseline_frame_us": base_f,
                "tile_errors": e.get_stat(STAT_WAVELET_TILED_ERRORS)})
if out_json:
    with open(out_json, "w") as fjs:
        json.dump(results, fjs, indent=1)
