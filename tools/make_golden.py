#!/usr/bin/env python
"""Generates tests/golden/*.json: SHA-256 digests + a few statistics of oracle outputs for fixed scenes, seeds
and frame counts.  The reference ships no golden vectors for this path and cannot run here, so these
fixtures pin the *oracle* (regression guard for the checker itself); tests/test_golden.py replays them on CPU
and tests/test_gpu_parity.py::test_golden_on_gpu on the CUDA path (exact mode).

    python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from strolle_b200 import scenes  # noqa: E402

BUFFERS = ["prim_triangle_ids", "prim_gbuffer_d0_a", "prim_gbuffer_d0_b", "di_reservoirs_0", "gi_reservoirs_0", "di_diff_curr_colors", "gi_diff_curr_colors", "output"]
CASES = {
    "cornell_128x128_f1": dict(scene=("cornell", dict(width=128, height=128)), frames=1),            # BASELINE config C1
    "cornell_96x54_f13": dict(scene=("cornell", dict(width=96, height=54)), frames=13),
    "cornell_ref_64x48_f4": dict(scene=("cornell", dict(width=64, height=48, mode=scenes.MODE_REFERENCE, ref_depth=1)), frames=4),   # config C5 shape
    "dungeon_96x54_f7": dict(scene=("dungeon", dict(width=96, height=54, cells=6)), frames=7),        # config C3 stand-in
    "cornell_spots_80x56_f5": dict(scene=("cornell_spots", dict(width=80, height=56)), frames=5),     # Light::Spot cone (glam acos_approx)
    "demo_level_96x54_f5": dict(scene=("demo_level", dict(width=96, height=54)), frames=5),          # config C3: the reference's dungeon asset, textures, sun + atmosphere
    "textured_room_80x44_f4": dict(scene=("textured_room", dict(width=80, height=44)), frames=4),     # texture atlas, alpha cut-outs, normal maps
}


def digest(a):
    """SHA-256 of the little-endian words with every NaN canonicalised (payloads differ between CPU and GPU)."""
    w = np.ascontiguousarray(a, dtype=np.float32).copy()
    w[np.isnan(w)] = np.float32(np.nan)
    u = w.view(np.uint32).copy()
    u[np.isnan(w)] = 0x7FC00000
    return hashlib.sha256(u.tobytes()).hexdigest()


def run_case(engine, case):
    kind, kw = case["scene"]
    sc = {"cornell": scenes.cornell, "dungeon": scenes.dungeon, "cornell_spots": scenes.cornell_spots, "demo_level": scenes.demo_level, "textured_room": scenes.textured_room}[kind](**kw)
    cam = scenes.apply(engine, sc)
    for _ in range(case["frames"]):
        engine.tick(); engine.render_camera(cam)
    out = {}
    for name in BUFFERS + ["ref_colors"]:
        try:
            b = engine.read_buffer(cam, name)
        except Exception:
            continue
        out[name] = {"sha256": digest(b), "mean": float(np.nanmean(b)), "nan": int(np.isnan(b).sum())}
    for name in ["triangles", "bvh", "lights", "sky_lut"]:
        out["scene:" + name] = {"sha256": digest(engine.read_scene(name))}
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    bn = scenes.blue_noise()
    only = sys.argv[1:]   # optional: names of the cases to (re)generate
    for name, case in CASES.items():
        if only and name not in only:
            continue
        e = pyoracle.OracleEngine(blue_noise=bn)
        res = run_case(e, case)
        with open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w") as f:
            json.dump({"case": {"scene": case["scene"][0], "args": case["scene"][1], "frames": case["frames"], "seed_base": "0xC0FFEE"}, "buffers": res}, f, indent=1, sort_keys=True)
        print(name, res["output"]["mean"] if "output" in res else "")
