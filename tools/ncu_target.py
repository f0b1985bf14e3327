"""Development aid: a short fixed workload for Nsight Compute captures (Cornell, N frames, chosen K22/K20 variants).

    ncu --set full --clock-control none --import-source on -k regex:wavelet -s 50 -c 5 -o gpurun_out/x python tools/ncu_target.py 31 0x00000 1
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strolle_b200
from strolle_b200 import scenes
from strolle_b200.engine import OPT_WAVELET_TILED, OPT_WAVELET_TILE_CFG, OPT_FUSE_REPROJECT

mask = int(sys.argv[1], 0) if len(sys.argv) > 1 else 31
cfg = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
fuse = int(sys.argv[3], 0) if len(sys.argv) > 3 else 1
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 14
w, h = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (1920, 1080)
e = strolle_b200.Engine()
e.set_option(OPT_WAVELET_TILED, mask); e.set_option(OPT_WAVELET_TILE_CFG, cfg); e.set_option(OPT_FUSE_REPROJECT, fuse)
cam = scenes.apply(e, {"dungeon": scenes.dungeon, "demo": scenes.demo_level}.get(os.environ.get("ST_SCENE", ""), scenes.cornell)(w, h))
for _ in range(frames):
    e.tick(); e.render_camera(cam)
e.synchronize()
print("ok", e.read_buffer(cam, "output").reshape(h, w, 4)[..., :3].mean())
