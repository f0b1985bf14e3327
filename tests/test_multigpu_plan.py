"""Host-side logic of the strip partition, on CPU: exchange plan, halo geometry, and the
torch.distributed transport with the gloo backend at world_size 2."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from strolle_b200 import multigpu as mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_strip_bounds_cover_frame():
    from strolle_b200.engine import strip_bounds_native
    for h, n in [(1080, 1), (1080, 2), (2160, 8), (4320, 8), (67, 3), (1528, 2), (2160, 4), (3056, 8), (720, 3), (400, 3), (5000, 16), (161, 16)]:
        b = mg.strip_bounds(h, n)
        assert b[0][0] == 0 and b[-1][1] == h
        assert all(b[i][1] == b[i + 1][0] for i in range(n - 1))
        assert all(y1 > y0 for y0, y1 in b)
        assert b == strip_bounds_native(h, n), "multigpu.strip_bounds and the engine's partition must be one function"


def test_strip_bounds_weigh_the_neighbours():
    """Outer strips (one neighbour) are STRIP_SIDE_ROWS taller than inner ones (two), when the strips are tall enough; equal otherwise."""
    rows = [y1 - y0 for y0, y1 in mg.strip_bounds(3056, 8)]
    assert rows[0] - rows[1] in (mg.STRIP_SIDE_ROWS - 1, mg.STRIP_SIDE_ROWS, mg.STRIP_SIDE_ROWS + 1) and rows[-1] - rows[-2] in (mg.STRIP_SIDE_ROWS - 1, mg.STRIP_SIDE_ROWS, mg.STRIP_SIDE_ROWS + 1)
    assert max(rows[1:-1]) - min(rows[1:-1]) <= 1 and min(rows) >= 160
    for h, n in [(1528, 2), (1080, 1), (400, 3), (1080, 8)]:
        rows = [y1 - y0 for y0, y1 in mg.strip_bounds(h, n)]
        assert max(rows) - min(rows) <= 1


def test_halo_transfers_cover_reach():
    bounds = mg.strip_bounds(1080, 8)   # 135-row strips: a 128-row halo stays inside the adjacent strips
    tr = mg.halo_transfers(bounds, 1080, 128)
    for dst, (d0, d1) in enumerate(bounds):
        need = set(range(max(0, d0 - 128), min(1080, d1 + 128))) - set(range(d0, d1))
        got = set()
        for s, d, a, b in tr:
            if d == dst:
                assert bounds[s][0] <= a < b <= bounds[s][1]
                got |= set(range(a, b))
        assert got == need
    # strips shorter than the reach pull from more than one neighbour
    tr = mg.halo_transfers(mg.strip_bounds(400, 8), 400, 128)
    assert len({s for s, d, a, b in tr if d == 3}) >= 4


def test_plan_frame_matches_schedule():
    P = mg
    image_odd = [P.P_PRIM_GBUFFER, P.P_FRAME_REPROJECTION, P.P_DI_SAMPLING, P.P_DI_TEMPORAL, P.P_DI_SPATIAL_PICK, P.P_DI_SPATIAL_TRACE,
                 P.P_DI_SPATIAL_SAMPLE, P.P_DI_RESOLVING, P.P_GI_REPROJECTION, P.P_GI_TEMPORAL, P.P_GI_SPATIAL_PICK, P.P_GI_SPATIAL_TRACE,
                 P.P_GI_SPATIAL_SAMPLE, P.P_GI_PREVIEW, P.P_GI_PREVIEW, P.P_GI_RESOLVING, P.P_DENOISE_REPROJECT, P.P_DENOISE_REPROJECT,
                 P.P_DENOISE_VARIANCE] + [P.P_DENOISE_WAVELET] * 5 + [P.P_COMPOSITION]
    plan = mg.plan_frame(image_odd, frame=1)
    by_step = {e.before_step: dict(e.buffers) for e in plan}
    assert by_step[0]["di_reservoirs_0"] == 16 and "prim_surface_map_a" in by_step[0]
    pick = by_step[image_odd.index(P.P_DI_SPATIAL_PICK)]
    assert pick == {"prim_gbuffer_d0_b": 128, "prim_gbuffer_d1_b": 128, "surface_nd": 128, "di_reservoirs_1": 128}
    assert by_step[image_odd.index(P.P_GI_SPATIAL_PICK)] == {"gi_reservoirs_1": 128}
    prev0 = image_odd.index(P.P_GI_PREVIEW)
    assert by_step[prev0] == {"prim_surface_map_b": 128, "gi_reservoirs_2": 128}
    assert by_step[prev0 + 1] == {"gi_reservoirs_3": 64}
    w0 = image_odd.index(P.P_DENOISE_WAVELET)
    assert [list(by_step[w0 + i].values())[0] for i in range(5)] == [1, 2, 4, 9, 19]
    assert list(by_step[w0].keys()) == ["di_diff_stash", "gi_diff_stash"] and list(by_step[w0 + 1].keys()) == ["di_diff_prev_colors", "gi_diff_prev_colors"]
    # even (sampling) frame: preview reads gi[1]
    even = [p for p in image_odd if p not in (P.P_GI_SPATIAL_PICK, P.P_GI_SPATIAL_TRACE, P.P_GI_SPATIAL_SAMPLE)]
    plan = mg.plan_frame(even, frame=2)
    by_step = {e.before_step: dict(e.buffers) for e in plan}
    assert by_step[even.index(P.P_GI_PREVIEW)] == {"prim_surface_map_a": 128, "gi_reservoirs_1": 128}


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from strolle_b200 import multigpu as mg
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
H, W = 64, 10
bounds = mg.strip_bounds(H, world)
full = torch.arange(H * W, dtype=torch.float32).view(H, W)
mine = torch.full((H, W), -1.0)
y0, y1 = bounds[rank]
mine[y0:y1] = full[y0:y1]            # each rank only holds its own rows
tr = mg.TorchDistTransport(rank)
for reach in (3, 20):
    ops = [(s, d, mine[a:b]) for s, d, a, b in mg.halo_transfers(bounds, H, reach) if s == rank or d == rank]
    tr.run(ops)
    lo, hi = max(0, y0 - reach), min(H, y1 + reach)
    assert torch.equal(mine[lo:hi], full[lo:hi]), (rank, reach)
    outside = torch.cat([mine[:lo], mine[hi:]])
    assert (outside == -1).all() or reach == 20
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_halo_exchange_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_native_plan_matches_python_plan():
    """The engine's C++ exchange plan (st_plan_frame, used by st_render_strips) is the same list as plan_frame."""
    from strolle_b200.engine import plan_frame_native
    full = ([mg.P_PRIM_GBUFFER, mg.P_FRAME_REPROJECTION, mg.P_DI_SAMPLING, mg.P_DI_TEMPORAL, mg.P_DI_SPATIAL_PICK, mg.P_DI_SPATIAL_TRACE,
             mg.P_DI_SPATIAL_SAMPLE, mg.P_DI_RESOLVING, mg.P_GI_REPROJECTION, mg.P_GI_SAMPLING_A, mg.P_GI_SAMPLING_B, mg.P_GI_TEMPORAL,
             mg.P_GI_SPATIAL_PICK, mg.P_GI_SPATIAL_TRACE, mg.P_GI_SPATIAL_SAMPLE, mg.P_GI_PREVIEW, mg.P_GI_PREVIEW, mg.P_GI_RESOLVING,
             mg.P_DENOISE_REPROJECT, mg.P_DENOISE_REPROJECT, mg.P_DENOISE_VARIANCE] + [mg.P_DENOISE_WAVELET] * 5 + [mg.P_COMPOSITION])
    gi_spatial = (mg.P_GI_SPATIAL_PICK, mg.P_GI_SPATIAL_TRACE, mg.P_GI_SPATIAL_SAMPLE)
    fused = list(full); fused.remove(mg.P_DENOISE_REPROJECT)   # ST_OPT_FUSE_REPROJECT (default): K20 for DI and GI is one step
    schedules = [full, fused, [p for p in fused if p not in gi_spatial], [p for p in full if p not in gi_spatial], [p for p in full if p not in gi_spatial + (mg.P_GI_PREVIEW,)],
                 [mg.P_PRIM_GBUFFER, mg.P_COMPOSITION]]
    for sched in schedules:
        for frame in (1, 2, 6, 7):
            for reach in (0, 16):
                native = plan_frame_native(sched, frame, reach)
                python = [(ex.before_step, name, r) for ex in mg.plan_frame(sched, frame, reach) for name, r in ex.buffers]
                assert native == python
    # the fused schedule shifts every later exchange point by one step and changes nothing else
    a = [(ex.before_step, tuple(ex.buffers)) for ex in mg.plan_frame(full, 1, 16)]
    b = [(ex.before_step, tuple(ex.buffers)) for ex in mg.plan_frame(fused, 1, 16)]
    cut = full.index(mg.P_DENOISE_REPROJECT)
    assert [(s - 1 if s > cut else s, bufs) for s, bufs in a] == b


# ---- fused strip transport: the order of one frame (engine.cu plan_strip_order, exported through st_plan_strip_order) --------------------

def _schedules():
    """Pass-id schedules as build_schedule (engine.cu) emits them: reference order and fused-pass order, the three GI cadences, DI-only,
    GI-only, denoise off, no instances."""
    P = mg
    svgf = [P.P_DENOISE_REPROJECT, P.P_DENOISE_VARIANCE] + [P.P_DENOISE_WAVELET] * 5
    di_ref = [P.P_DI_SAMPLING, P.P_DI_TEMPORAL, P.P_DI_SPATIAL_PICK, P.P_DI_SPATIAL_TRACE, P.P_DI_SPATIAL_SAMPLE, P.P_DI_RESOLVING]
    di_fused = [P.P_DI_TEMPORAL, P.P_DI_SPATIAL_PICK, P.P_DI_RESOLVING]
    def gi(kind, fused):
        samp = [P.P_GI_SAMPLING_B] if fused else [P.P_GI_SAMPLING_A, P.P_GI_SAMPLING_B]
        spat = [P.P_GI_SPATIAL_PICK] if fused else [P.P_GI_SPATIAL_PICK, P.P_GI_SPATIAL_TRACE, P.P_GI_SPATIAL_SAMPLE]
        tail = [P.P_GI_PREVIEW, P.P_GI_PREVIEW] + ([] if fused else [P.P_GI_RESOLVING])
        head = [] if (fused and kind != "validate") else [P.P_GI_REPROJECTION]
        body = {"even": samp + [P.P_GI_TEMPORAL], "odd": [P.P_GI_TEMPORAL] + spat, "validate": samp + [P.P_GI_TEMPORAL]}[kind]
        return head + body + tail
    out = {}
    for fused in (False, True):
        di = di_fused if fused else di_ref
        for kind in ("even", "odd", "validate"):
            out[f"image-{kind}-{'fused' if fused else 'ref'}"] = [P.P_PRIM_GBUFFER, P.P_FRAME_REPROJECTION] + di + gi(kind, fused) + svgf + [P.P_COMPOSITION]
        out[f"di-only-{'fused' if fused else 'ref'}"] = [P.P_PRIM_GBUFFER, P.P_FRAME_REPROJECTION] + di + svgf + [P.P_COMPOSITION]
        out[f"gi-only-nodenoise-{'fused' if fused else 'ref'}"] = [P.P_PRIM_GBUFFER, P.P_FRAME_REPROJECTION] + gi("odd", fused) + [P.P_COMPOSITION]
    out["no-instances"] = [P.P_PRIM_GBUFFER] + svgf + [P.P_COMPOSITION]
    return out


@pytest.mark.parametrize("still", [False, True])
@pytest.mark.parametrize("dma", [3, 2, 1, 0])
def test_fused_strip_order_invariants(dma, still):
    """Every pass runs exactly once and chains keep their internal order; every pass that gathers from a neighbouring strip is preceded
    by a wait on the flag that its producer's signal (or copy-engine push) raises; nothing a neighbour may still pull is overwritten before
    every rank signalled PULL_DONE; the frame opens with the previous frame's FRAME_DONE and closes with this frame's.  On a frame where
    nothing moved (`still`) there is no pull and no wait for PULL_DONE (it is still raised, for ranks that might pull)."""
    from strolle_b200.engine import plan_strip_order
    P = mg
    producer_slot = {P.P_DI_TEMPORAL: "DI1", P.P_GI_TEMPORAL: "GI1", P.P_DENOISE_REPROJECT: "SVGF"}
    for name, sched in _schedules().items():
        ops = plan_strip_order(sched, dma, still)
        steps = [int(o.split(":")[1]) for o in ops if o.startswith("step:")]
        assert sorted(steps) == list(range(len(sched))), name
        for chain in ([P.P_DI_SAMPLING, P.P_DI_TEMPORAL, P.P_DI_SPATIAL_PICK, P.P_DI_SPATIAL_TRACE, P.P_DI_SPATIAL_SAMPLE, P.P_DI_RESOLVING],
                      [P.P_GI_REPROJECTION, P.P_GI_SAMPLING_A, P.P_GI_SAMPLING_B, P.P_GI_TEMPORAL, P.P_GI_SPATIAL_PICK, P.P_GI_SPATIAL_TRACE, P.P_GI_SPATIAL_SAMPLE, P.P_GI_PREVIEW, P.P_GI_RESOLVING],
                      [P.P_DENOISE_REPROJECT, P.P_DENOISE_VARIANCE, P.P_DENOISE_WAVELET, P.P_COMPOSITION]):
            inside = [i for i in steps if sched[i] in chain]
            assert inside == sorted(inside), f"{name}: chain order"
        assert ops[0] == "step:0" and ops[1] == "wait:FRAME_DONE:all:prev" and ops[-1] == "signal:FRAME_DONE:all", name
        if still:
            assert "pull" not in ops and ops[2] == "signal:PULL_DONE:all" and not any(o.startswith("wait:PULL_DONE") for o in ops), name
        else:
            assert ops[2] == "pull" and ops[3] == "signal:PULL_DONE:all", name
        raised, waited = set(), set()
        previews = 0
        for o in ops:
            f = o.split(":")
            if f[0] == "signal":
                raised.add(f[1])
            elif f[0] == "push":
                assert dma and f[1] in ("gi_reservoirs_1", "gi_reservoirs_2", "@gbuffer", "di_reservoirs_1", "gi_reservoirs_3") and (f[1] in ("gi_reservoirs_1", "gi_reservoirs_2") or dma >= 3 or (dma == 2 and f[1] == "@gbuffer"))
                raised.add(f[2])
            elif f[0] == "wait" and len(f) == 3:
                assert f[1] in raised, f"{name}: waits for {f[1]} before this rank raised it itself (ranks run the same order: nobody would)"
                waited.add(f[1])
            elif f[0] == "step":
                p = sched[int(f[1])]
                need = None
                if p == P.P_DI_SPATIAL_PICK: need = "DI1"
                elif p == P.P_GI_SPATIAL_PICK: need = "GI1"
                elif p == P.P_GI_PREVIEW:
                    previews += 1
                    need = ("GI2" if any(q == P.P_GI_SPATIAL_PICK for q in sched) else "GI1") if previews == 1 else "GI3"
                elif p == P.P_DENOISE_VARIANCE: need = "SVGF"
                if need:
                    assert need in waited, f"{name}: pass {p} gathers before wait:{need}"
                    assert dma < 2 or "GBUF" in waited, f"{name}: pass {p} reads the neighbours' G-buffer rows before wait:GBUF"
                if p in (P.P_DI_RESOLVING, P.P_GI_RESOLVING, P.P_DENOISE_WAVELET) or (p == P.P_GI_PREVIEW and previews == 2):
                    assert still or "PULL_DONE" in waited, f"{name}: pass {p} overwrites pulled buffers before every rank pulled"
                if p in producer_slot and producer_slot[p] not in raised:
                    pass   # raised right after the producer (checked through the waits above)
        if any(q == P.P_GI_PREVIEW for q in sched):
            assert "GI3" in raised and "GI3" in waited, name
        assert ("GBUF" in raised) == (dma >= 2) and ("GBUF" in waited) == (dma >= 2), f"{name}: the G-buffer flag is raised and consumed every frame, or never"
