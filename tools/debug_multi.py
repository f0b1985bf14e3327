"""Development aid: where a strip group first departs from the single-GPU frame (buffer, rank, rows)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import strolle_b200
from strolle_b200 import scenes
from strolle_b200.engine import OPT_STRIP_DMA, OPT_FUSED_PASSES
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import CAMERA_BUFFERS

n, w, h, exact = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
dma = int(sys.argv[5]) if len(sys.argv) > 5 else 1
frames = int(sys.argv[6]) if len(sys.argv) > 6 else 2
scene = scenes.cornell(w, h)
one = strolle_b200.Engine(exact=exact)
grp = strolle_b200.MultiEngine([0] * n, exact=exact)
grp.set_option(OPT_STRIP_DMA, dma)
c1, cn = scenes.apply(one, scene), scenes.apply(grp, scene)
from strolle_b200.multigpu import strip_bounds
bounds = strip_bounds(h, n)
for f in range(frames):
    one.tick(); grp.tick(); one.render_camera(c1); grp.render_camera(cn)
    print(f"frame {f + 1}: peer errors {grp.peer_errors(cn)}; first wait that gave up per rank (slot << 16 | awaited rank << 8 | seq): "
          + ", ".join(hex(grp.member(r).get_stat(7)) for r in range(n)))
    for name in CAMERA_BUFFERS:
        want = one.read_buffer(c1, name).reshape(h, -1)
        for r in range(n):
            m = grp.member(r)
            got = m.read_buffer(grp.member_camera(cn, r), name).reshape(h, -1)
            bad = (got.view(np.uint32) != want.view(np.uint32)) & ~(np.isnan(got) & np.isnan(want))
            rows = np.flatnonzero(bad.any(axis=1))
            own = [y for y in rows if bounds[r][0] <= y < bounds[r][1]]
            halo = [y for y in rows if not (bounds[r][0] <= y < bounds[r][1]) and bounds[r][0] - 128 <= y < bounds[r][1] + 128]
            if own or (halo and name in ("di_reservoirs_1", "gi_reservoirs_1", "gi_reservoirs_2", "gi_reservoirs_3", "surface_nd", "prim_gbuffer_d0_a", "prim_gbuffer_d0_b")):
                print(f"  {name:22s} rank {r}: own rows differing {own[:3]}..{own[-3:] if own else []} ({len(own)}), halo rows differing {halo[:3]}..{halo[-3:] if halo else []} ({len(halo)})")
