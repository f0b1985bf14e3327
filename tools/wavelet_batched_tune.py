"""Development aid: K22 per à-trous iteration, default kernels vs the batched-gather kernel (ST_OPT_WAVELET_BATCHED), for the library
selected by STROLLE_B200_LIB (register-cap variants built with ST_WAVELET_BATCHED_MINB).  python tools/wavelet_batched_tune.py [scene]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strolle_b200
from strolle_b200 import scenes
from strolle_b200.engine import OPT_WAVELET_BATCHED, OPT_WAVELET_TILED

name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
e = strolle_b200.Engine()
cam = scenes.apply(e, {"cornell": scenes.cornell, "demo": scenes.demo_level}[name](1920, 1080))
for _ in range(12):
    e.tick(); e.render_camera(cam)


def per_iteration(batched, tiled=15):
    e.set_option(OPT_WAVELET_BATCHED, batched); e.set_option(OPT_WAVELET_TILED, tiled)
    for _ in range(4):
        e.tick(); e.render_camera(cam)
    e.synchronize(); e.enable_timing(True); e.wavelet_times(reset=True)
    for _ in range(18):
        e.tick(); e.render_camera(cam)
    e.synchronize()
    ms, n = e.wavelet_times(reset=True)
    e.enable_timing(False)
    return [float(ms[i]) / max(int(n[i]), 1) * 1000.0 for i in range(5)]


base = per_iteration(0)
gather = per_iteration(0, tiled=0)
bat = per_iteration(31)
print(f"{os.environ.get('STROLLE_B200_LIB', 'default lib')} [{name}] us per launch, strides 1 2 4 8 16")
print("  default (tiled 1-8, gather 16): " + " ".join(f"{v:6.1f}" for v in base))
print("  plain gather:                   " + " ".join(f"{v:6.1f}" for v in gather))
print("  batched gather:                 " + " ".join(f"{v:6.1f}" for v in bat))
