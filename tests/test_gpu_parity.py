"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Integer/index results (triangle ids, material ids, reservoir light ids, validity bits, RNG state)
and every f32 buffer are compared bit-for-bit; NaN matches NaN (empty GI reservoirs store NaN
oct-normals in the reference too).  north_star's tolerance for colours is 1e-3 relative L2 — also
asserted, trivially, and against the libm flavour of the oracle.
"""
import numpy as np
import pytest

from strolle_b200 import scenes
from tests.util import CAMERA_BUFFERS, SCENE_BUFFERS, assert_bits_equal, bits_equal, random_rays, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import strolle_b200
    return strolle_b200


def make_pair(gpu, oracle, blue_noise, scene, libm=False, exact=True):
    eg = gpu.Engine(blue_noise=blue_noise, exact=exact)
    eo = oracle.OracleEngine(libm=libm, blue_noise=blue_noise)
    cg = scenes.apply(eg, scene)
    co = scenes.apply(eo, scene)
    return eg, cg, eo, co


def test_device_math_bit_exact(gpu, oracle, blue_noise):
    e = gpu.Engine(blue_noise=blue_noise)
    rng = np.random.RandomState(0)
    cases = {
        "sin": (np.concatenate([rng.uniform(-20, 20, 200000), [0.0, -0.0, np.pi, 2 * np.pi]]), None),
        "cos": (rng.uniform(-20, 20, 200000), None),
        "acos": (np.concatenate([rng.uniform(-1, 1, 200000), [1.0, -1.0, 0.5, -0.5, 0.0, 1.5]]), None),
        "atan2": (rng.normal(size=200000), rng.normal(size=200000)),
        "acos_approx": (np.concatenate([rng.uniform(-1.2, 1.2, 200000), [1.0, -1.0, 0.0, -0.0, 2.0, -2.0]]), None),   # glam's polynomial (spot-light cone)
        "exp": (np.concatenate([rng.uniform(-110, 90, 200000), [0.0, -1e9, 1e9]]), None),
        "pow": (np.concatenate([rng.uniform(0, 2, 200000), [0.0, 1.0]]), np.concatenate([rng.choice([2.2, 1 / 2.2, 8.0, 5.0, 64.0, 3.0, 1.5, 2.0], 200000), [2.2, 5.0]])),
    }
    for op, (a, b) in cases.items():
        assert_bits_equal(e.device_math(op, a, b), oracle.math(op, a, b), f"device {op}")


@pytest.mark.parametrize("scene_name", ["cornell", "dungeon"])
def test_scene_upload_bit_exact(gpu, oracle, blue_noise, scene_name):
    scene = scenes.cornell(64, 64) if scene_name == "cornell" else scenes.dungeon(64, 64, cells=6)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    eg.tick(); eo.tick()
    for name in SCENE_BUFFERS:
        assert_bits_equal(eg.read_scene(name), eo.read_scene(name), f"{scene_name}:{name}")
    assert eg.bvh_depth() == eo.bvh_depth()
    assert_bits_equal(eg.read_buffer(cg, "curr_camera"), eo.read_buffer(co, "curr_camera"), "camera uniform")


@pytest.mark.parametrize("scene_name", ["cornell", "dungeon"])
def test_trace_streams_bit_exact(gpu, oracle, blue_noise, scene_name):
    if scene_name == "cornell":
        scene, lo, hi = scenes.cornell(64, 64), (-1.0, 0.0, -1.0), (1.0, 2.0, 3.2)
    else:
        scene, lo, hi = scenes.dungeon(64, 64, cells=8), (-20.0, 0.1, -40.0), (10.0, 2.9, -5.0)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    eg.tick(); eo.tick()
    rays = random_rays(200000, 3, lo, hi)
    hg, ho = eg.trace_closest(rays), eo.trace_closest(rays)
    assert (hg[:, 9].view(np.uint32) == ho[:, 9].view(np.uint32)).all(), "hit triangle ids are bit-exact"
    assert (hg[:, 10].view(np.uint32) == ho[:, 10].view(np.uint32)).all(), "hit material ids are bit-exact"
    assert_bits_equal(hg, ho, "closest-hit records (point, oct normal, uv, distance, used_memory)")
    assert (ho[:, 8] < 3e38).mean() > 0.3
    rays_any = random_rays(200000, 4, lo, hi, max_len=4.0)
    og, oo = eg.trace_any(rays_any), eo.trace_any(rays_any)
    assert (og == oo).all() and 0.05 < og.mean() < 0.95


def run_and_compare(eg, cg, eo, co, frames, buffers=CAMERA_BUFFERS, what=""):
    for f in range(frames):
        eg.tick(); eo.tick()
        eg.render_camera(cg); eo.render_camera(co)
        for name in buffers:
            ok, msg = bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name))
            assert ok, f"{what} frame {f + 1} buffer {name}: {msg}"
    img_g = eg.read_buffer(cg, "output").reshape(-1, 4)[:, :3]
    img_o = eo.read_buffer(co, "output").reshape(-1, 4)[:, :3]
    assert rel_l2(img_g, img_o) <= 1e-3   # north_star tolerance (bit-exact is asserted above)
    return img_g


def test_cornell_c1_128_single_frame(gpu, oracle, blue_noise):
    """BASELINE config C1: Cornell 128x128, single frame (frame id 1), Image mode."""
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(128, 128))
    img = run_and_compare(eg, cg, eo, co, 1, what="C1")
    assert np.isfinite(img).all() and img.mean() > 0.05


def test_cornell_full_pipeline_13_frames(gpu, oracle, blue_noise):
    """Two full 6-frame GI cycles + 1 (tracing-even, tracing-odd/spatial and validation frames),
    DI + GI + SVGF, every per-camera buffer compared after every frame."""
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(160, 90))
    run_and_compare(eg, cg, eo, co, 13, what="cornell 160x90")


def test_cornell_odd_sizes_and_small_screens(gpu, oracle, blue_noise):
    """Ragged sizes: odd width (unpaired checkerboard column), non-multiple-of-8, and a screen
    smaller than the 128 px tap radius (taps land outside after Camera::contain's single mirror)."""
    for (w, h) in [(121, 67), (50, 40), (8, 8)]:
        eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(w, h))
        run_and_compare(eg, cg, eo, co, 4, what=f"cornell {w}x{h}")


def test_moving_camera_reprojection(gpu, oracle, blue_noise):
    """Camera motion exercises velocity, the bilinear history fetch and validity bits (K4/K20)."""
    scene = scenes.cornell(128, 72)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    c = scene["camera"]
    for f in range(6):
        eye = (0.02 * f, 1.0 + 0.01 * f, 3.2 - 0.03 * f)
        t = scenes.look_at_transform(eye, (0.0, 1.0, 0.0))
        for e, cam in ((eg, cg), (eo, co)):
            e.update_camera(cam, c["mode"], c["denoise"], c["ref_depth"], c["w"], c["h"], t, c["projection"])
            e.tick(); e.render_camera(cam)
        for name in CAMERA_BUFFERS:
            ok, msg = bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name))
            assert ok, f"moving camera frame {f + 1} {name}: {msg}"
    rp = eo.read_buffer(co, "reprojection_map").reshape(-1, 4)
    assert ((rp[:, 0] % 1.0) != 0).any(), "non-exact reprojection path was exercised"


def test_dungeon_with_atmosphere(gpu, oracle, blue_noise):
    """Config C3 stand-in (synthetic dungeon, 6 point lights + sun above the horizon -> sky LUT live)."""
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.dungeon(160, 90, cells=8))
    run_and_compare(eg, cg, eo, co, 7, what="dungeon 160x90")


def test_reference_mode_and_heatmap(gpu, oracle, blue_noise):
    """Config C5 shape: Reference{depth:2} accumulation; plus the BVH heatmap (K3)."""
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(128, 72, mode=scenes.MODE_REFERENCE, ref_depth=2))
    run_and_compare(eg, cg, eo, co, 5, buffers=["ref_hits", "ref_rays", "ref_colors", "output"], what="reference mode")
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(128, 72, mode=scenes.MODE_BVH_HEATMAP))
    run_and_compare(eg, cg, eo, co, 1, buffers=["ref_colors", "output"], what="heatmap")


def test_light_updates_and_removal(gpu, oracle, blue_noise):
    """Light slot kill/remap protocol (strolle/src/lights.rs:101-162) seen through K6."""
    scene = scenes.cornell(96, 54)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    extra = scenes.point_light((0.5, 1.0, 0.2), 0.1, (2.0, 1.0, 0.5), 10.0)
    for e in (eg, eo):
        e.insert_light(401, scenes.LIGHT_POINT, extra)
        e.insert_light(402, scenes.LIGHT_POINT, scenes.point_light((-0.5, 0.7, 0.1), 0.1, (0.5, 1.0, 2.0), 10.0))
    for f in range(6):
        if f == 2:
            for e in (eg, eo):
                e.remove_light(401)
        if f == 4:
            for e in (eg, eo):
                e.insert_light(402, scenes.LIGHT_POINT, scenes.point_light((-0.4, 0.8, 0.1), 0.1, (0.5, 1.0, 2.0), 10.0))
        eg.tick(); eo.tick()
        assert_bits_equal(eg.read_scene("lights"), eo.read_scene("lights"), f"lights frame {f + 1}")
        eg.render_camera(cg); eo.render_camera(co)
        for name in ["di_reservoirs_0", "di_reservoirs_1", "di_reservoirs_2", "di_diff_samples", "output"]:
            ok, msg = bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name))
            assert ok, f"lights frame {f + 1} {name}: {msg}"


def test_libm_oracle_within_tolerance(gpu, oracle, blue_noise):
    """Against the libm flavour of the oracle (host libm instead of the shared polynomial kernels) the
    CUDA image stays inside north_star's 1e-3 relative-L2 tolerance after a single frame."""
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(160, 90), libm=True)
    eg.tick(); eo.tick(); eg.render_camera(cg); eo.render_camera(co)
    a = eg.read_buffer(cg, "output").reshape(-1, 4)[:, :3]
    b = eo.read_buffer(co, "output").reshape(-1, 4)[:, :3]
    assert rel_l2(a, b) <= 1e-3
    tid_g = eg.read_buffer(cg, "prim_triangle_ids").view(np.uint32)
    tid_o = eo.read_buffer(co, "prim_triangle_ids").view(np.uint32)
    assert (tid_g == tid_o).all(), "primary-hit triangle indices are bit-exact regardless of the libm flavour"


class _LocalTransport:
    """Several strip engines inside one process on one GPU: the same exchange plan, executed as row copies
    between the engines' buffers (stands in for NCCL so that strip correctness is testable on a 1-GPU box)."""

    def __init__(self):
        self.runners = []

    def run_all(self, per_rank_ops):
        import torch
        torch.cuda.synchronize()
        for ops in per_rank_ops:
            for src, dst, name, a, b in ops:
                self.runners[dst]._view(name)[a:b].copy_(self.runners[src]._view(name)[a:b])
        torch.cuda.synchronize()


@pytest.mark.parametrize("world,size,tiled", [(2, (160, 96), 0), (4, (128, 160), 0), (2, (160, 96), 31), (3, (200, 150), 31),
                                              (4, (1920, 1080), 31)])   # the last one: BASELINE.json's frame size, 270-row strips (config C4's strip height)
def test_row_strips_reproduce_single_gpu_frame(gpu, blue_noise, world, size, tiled):
    """Strip-partitioned rendering (SURVEY §8e, config C4's shape) is bit-identical to the single-GPU frame:
    `world` engines each compute one row strip and exchange halo rows according to multigpu.plan_frame."""
    import torch
    from strolle_b200 import multigpu as mg
    w, h = size
    scene = scenes.cornell(w, h)
    from strolle_b200.engine import OPT_WAVELET_TILED, OPT_FUSE_REPROJECT, OPT_VARIANCE_TILED
    ref = gpu.Engine(blue_noise=blue_noise)   # default (fast SVGF weights): strips must match it bit for bit too
    ref.set_option(OPT_WAVELET_TILED, 0); ref.set_option(OPT_FUSE_REPROJECT, 0); ref.set_option(OPT_VARIANCE_TILED, 0)   # the full frame runs the plain kernels ...
    cref = scenes.apply(ref, scene)
    engines, cams, runners = [], [], []
    lt = _LocalTransport()
    for r in range(world):
        e = gpu.Engine(blue_noise=blue_noise)
        e.set_option(OPT_WAVELET_TILED, tiled); e.set_option(OPT_FUSE_REPROJECT, 1 if tiled else 0); e.set_option(OPT_VARIANCE_TILED, 1 if tiled else 0)   # ... the strips also the tile-staged K22 / fused K20
        c = scenes.apply(e, scene)
        rn = mg.StripRunner(e, c, w, h, rank=0, world=1)   # built as single, then configured as a strip by hand
        rn.rank, rn.world = r, world
        rn.bounds = mg.strip_bounds(h, world)
        rn.y0, rn.y1 = rn.bounds[r]
        e.set_strip(c, rn.y0, rn.y1)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        engines.append(e); cams.append(c); runners.append(rn)
    lt.runners = runners
    for f in range(7 if w * h < 500000 else 3):
        ref.tick(); ref.render_camera(cref)
        for e in engines:
            e.tick()
        schedule = engines[0].frame_schedule(cams[0])
        plan = mg.plan_frame(schedule, engines[0].frame() - 1, temporal_reach=16)
        first = 0
        for ex in plan:
            if ex.before_step > first:
                for e, c in zip(engines, cams):
                    e.render_range(c, first, ex.before_step - 1)
                first = ex.before_step
            per_rank = []
            for name, reach in ex.buffers:
                per_rank.append([(s, d, name, a, b) for s, d, a, b in mg.halo_transfers(runners[0].bounds, h, reach)])
            lt.run_all(per_rank)
        for e, c in zip(engines, cams):
            e.render_range(c, first, len(schedule) - 1)
        full = ref.read_buffer(cref, "output").reshape(h, w, 4)
        for name in ["output", "di_reservoirs_0", "gi_reservoirs_0", "di_diff_curr_colors", "gi_diff_curr_colors", "di_diff_prev_colors"]:
            want = ref.read_buffer(cref, name).reshape(h, -1)
            for rn, e, c in zip(runners, engines, cams):
                got = e.read_buffer(c, name).reshape(h, -1)
                ok, msg = bits_equal(got[rn.y0:rn.y1], want[rn.y0:rn.y1])
                assert ok, f"strip {rn.rank}/{world} frame {f + 1} {name}: {msg}"
        assert np.isfinite(full).all()


SVGF_BUFFERS = {"di_diff_prev_colors", "di_diff_curr_colors", "di_diff_stash", "gi_diff_prev_colors", "gi_diff_curr_colors", "gi_diff_stash", "output"}


def test_default_fast_svgf_mode_within_tolerance(gpu, oracle, blue_noise):
    """The product default evaluates the SVGF edge-stopping weights with SFU approximations (ST_OPT_SVGF_FAST_MATH).
    Everything that is not a denoiser colour buffer stays bit-exact; the denoised colours and the composed image stay
    inside north_star's tolerance: 1e-3 relative per-channel L2, after 13 frames of temporal feedback."""
    from strolle_b200.engine import OPT_SHADING_FAST_MATH, OPT_FUSED_PASSES
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(192, 108), exact=False)
    eg.set_option(OPT_SHADING_FAST_MATH, 0); eg.set_option(OPT_FUSED_PASSES, 0)   # only the denoiser's weights are approximate here; test_fast_shading_* covers the product default
    for f in range(13):
        eg.tick(); eo.tick(); eg.render_camera(cg); eo.render_camera(co)
        for name in CAMERA_BUFFERS:
            if name.endswith("_stash"):
                continue   # scratch: with ST_OPT_WAVELET_PAIRED the stride-8 iteration's output goes to interleaved records instead (test_paired_wavelet_records_match_planar)
            a, b = eg.read_buffer(cg, name), eo.read_buffer(co, name)
            if name in SVGF_BUFFERS:
                a3, b3 = a.reshape(-1, 4)[:, :3], b.reshape(-1, 4)[:, :3]
                for ch in range(3):
                    assert rel_l2(a3[:, ch], b3[:, ch]) <= 1e-3, f"frame {f + 1} {name} channel {ch}: {rel_l2(a3[:, ch], b3[:, ch])}"
            else:
                ok, msg = bits_equal(a, b)
                assert ok, f"fast-SVGF mode must leave {name} bit-exact (frame {f + 1}): {msg}"


def test_sample_parallel_reference_mode(gpu, blue_noise):
    """Config C5's shape: Reference{depth:1} accumulations split across ranks (here 4 engines on one GPU, the
    NCCL reduce replaced by a tensor sum) equal the single-engine accumulation up to f32 addition order."""
    import torch
    from strolle_b200 import multigpu as mg
    w, h, total, world = 128, 72, 16, 4
    scene = scenes.cornell(w, h, mode=scenes.MODE_REFERENCE, ref_depth=1)
    ref = gpu.Engine(blue_noise=blue_noise)
    cref = scenes.apply(ref, scene)
    for _ in range(total):
        ref.tick(); ref.render_camera(cref)
    want = ref.read_buffer(cref, "ref_colors").reshape(-1, 4)
    assert (want[:, 3] == total).all()
    parts = []
    for r in range(world):
        e = gpu.Engine(blue_noise=blue_noise)
        c = scenes.apply(e, scene)
        acc = mg.ReferenceAccumulator(e, c, rank=r, world=1)
        acc.rank, acc.world = r, world
        acc.accumulate(total)
        parts.append(e.read_buffer(c, "ref_colors").reshape(-1, 4))
        assert (parts[-1][:, 3] == total // world).all()
    got = np.sum(parts, axis=0, dtype=np.float32)
    assert (got[:, 3] == total).all()
    for ch in range(3):
        assert rel_l2(got[:, ch], want[:, ch]) <= 1e-6


def test_texture_atlas_and_alpha_cutout(gpu, oracle, blue_noise):
    """sRGB atlas textures (repeat-wrapped, negative uvs), emissive / metallic-roughness textures and an
    AlphaMode::Blend cutout that primary, shadow and bounce rays must pass through — bit-exact over 7 frames, plus a
    ray stream through the fence."""
    scene = scenes.textured_room(192, 108)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    eg.tick(); eo.tick()
    assert_bits_equal(eg.read_scene("materials"), eo.read_scene("materials"), "materials with atlas rects")
    bvh = eo.read_scene("bvh").reshape(-1, 4).view(np.uint32)
    assert ((bvh[:, 3] == 1) & ((bvh[:, 0] & 2) == 2)).any(), "the fence's leaf entries carry the alpha-blend flag"
    rays = random_rays(100000, 9, (-2.0, 0.1, 1.0), (2.0, 2.0, 4.0))
    rays[:, 4:7] = np.array([0.0, 0.0, -1.0], np.float32) + 0.2 * rays[:, 4:7]
    rays[:, 4:7] /= np.linalg.norm(rays[:, 4:7], axis=1, keepdims=True)
    hg, ho = eg.trace_closest(rays), eo.trace_closest(rays)
    assert_bits_equal(hg, ho, "closest hits through the alpha cutout")
    fence = (ho[:, 10].view(np.uint32) == 2) & (ho[:, 8] < 3e38)
    behind = (ho[:, 10].view(np.uint32) != 2) & (ho[:, 8] < 3e38) & (ho[:, 2] < 0.4)
    assert fence.mean() > 0.1 and behind.mean() > 0.1, "some rays stop at opaque fence texels, others pass through the holes"
    eg2, cg2, eo2, co2 = make_pair(gpu, oracle, blue_noise, scene)
    run_and_compare(eg2, cg2, eo2, co2, 7, what="textured room")


def test_moving_instance_rebuilds_bvh_and_velocity(gpu, oracle, blue_noise):
    """Dynamic scene: re-inserting an instance with a new transform re-bakes its triangles, rebuilds the BVH
    (strolle/src/instances.rs:69-139, bvh.rs:48-70) and feeds prev_transform into the velocity map
    (passes/prim_raster.rs:198-223); removing an instance frees its triangle slots."""
    scene = scenes.cornell(128, 72)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    h, mesh, mat, xf = scene["instances"][6]   # the short box
    for f in range(7):
        if 1 <= f <= 4:
            moved = np.array(xf, np.float32).copy(); moved[9] += 0.05 * f; moved[10] += 0.02 * f
            for e in (eg, eo):
                e.insert_instance(h, mesh, mat, moved)
        if f == 5:
            for e in (eg, eo):
                e.remove_instance(scene["instances"][7][0])   # drop the tall box
        eg.tick(); eo.tick()
        for name in ["triangles", "bvh"]:
            assert_bits_equal(eg.read_scene(name), eo.read_scene(name), f"frame {f + 1} scene:{name}")
        eg.render_camera(cg); eo.render_camera(co)
        for name in CAMERA_BUFFERS:
            ok, msg = bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name))
            assert ok, f"moving instance frame {f + 1} {name}: {msg}"
    vel = eo.read_buffer(co, "velocity_map").reshape(-1, 4)
    assert (vel[:, :2] != 0).any(), "the moved box still reports a velocity (stale prev_transform, as in the reference)"


@pytest.mark.parametrize("reuse", [True, False])
def test_bvh_subtree_reuse_and_stale_material_quirk(gpu, oracle, blue_noise, reuse):
    """BVH refreshes graft the unchanged subtrees of the previous tree (strolle/src/bvh/builder.rs:245-359).
    Moving a box rebuilds only its side of the tree; giving a box another material without moving it changes no
    primitive centre, so the reference's centre-only hash (primitive.rs:27-37) reuses the whole tree and the old material id
    stays in the leaves (quirk C-20).  reuse=False (ST_OPT_BVH_REUSE 0) builds from scratch instead.  Either way the CUDA
    path and the oracle agree bit for bit."""
    from strolle_b200.engine import OPT_BVH_REUSE, STAT_BVH_GRAFTED_SUBTREES
    scene = scenes.cornell(96, 64)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    eg.set_option(OPT_BVH_REUSE, int(reuse)); eo.set_bvh_reuse(reuse)
    h, mesh, mat, xf = scene["instances"][6]
    other_mat = scene["instances"][1][2]
    assert other_mat != mat
    bvh_before = None
    for f in range(5):
        if f == 1:
            moved = np.array(xf, np.float32).copy(); moved[9] += 0.1
            for e in (eg, eo):
                e.insert_instance(h, mesh, mat, moved)
        if f == 3:
            for e in (eg, eo):
                e.insert_instance(h, mesh, other_mat, moved)   # same place, other material
        eg.tick(); eo.tick()
        for name in ["triangles", "bvh"]:
            assert_bits_equal(eg.read_scene(name), eo.read_scene(name), f"frame {f + 1} scene:{name}")
        if f == 1:
            assert (eg.get_stat(STAT_BVH_GRAFTED_SUBTREES) > 0) == reuse and (eo.bvh_reused() > 0) == reuse
        if f == 2:
            bvh_before = eg.read_scene("bvh").copy()
        if f == 3:
            same = bits_equal(eg.read_scene("bvh"), bvh_before)[0]
            assert same == reuse, "reuse keeps the old material id in the leaves; a fresh build does not"
        eg.render_camera(cg); eo.render_camera(co)
        for name in ["prim_gbuffer_d1_a", "prim_gbuffer_d1_b", "di_reservoirs_0", "gi_reservoirs_0", "output"]:
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"frame {f + 1} {name}")


DENOISER_BUFFERS = ["di_diff_prev_colors", "di_diff_curr_colors", "di_diff_stash", "di_diff_moments_a", "di_diff_moments_b",
                    "gi_diff_prev_colors", "gi_diff_curr_colors", "gi_diff_stash", "gi_diff_moments_a", "gi_diff_moments_b", "output"]


@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
@pytest.mark.parametrize("exact", [True, False])
def test_tiled_wavelet_matches_gather(gpu, blue_noise, cfg, exact):
    """K22 staged through shared memory by TMA tensor copies (all five strides, every tile shape) produces the same
    bits as the per-tap gather kernel, in the strict-IEEE and in the fast-SVGF flavour, on frame sizes that are
    not multiples of the tile (zero-filled borders) and smaller than the largest stride's reach."""
    from strolle_b200.engine import OPT_WAVELET_TILED, OPT_WAVELET_TILE_CFG, OPT_WAVELET_PAIRED, STAT_WAVELET_TILED_LAUNCHES, STAT_WAVELET_TILED_ERRORS
    for size in [(200, 120), (67, 45)]:
        scene = scenes.cornell(*size)
        ea, eb = gpu.Engine(blue_noise=blue_noise, exact=exact), gpu.Engine(blue_noise=blue_noise, exact=exact)
        ea.set_option(OPT_WAVELET_PAIRED, 0); eb.set_option(OPT_WAVELET_PAIRED, 0)   # planar buffers: every stride can run tile-staged
        ea.set_option(OPT_WAVELET_TILED, 0)
        eb.set_option(OPT_WAVELET_TILED, 31); eb.set_option(OPT_WAVELET_TILE_CFG, cfg * 0x11111)
        ca, cb = scenes.apply(ea, scene), scenes.apply(eb, scene)
        for f in range(4):
            ea.tick(); eb.tick(); ea.render_camera(ca); eb.render_camera(cb)
            for name in DENOISER_BUFFERS:
                assert_bits_equal(eb.read_buffer(cb, name), ea.read_buffer(ca, name), f"{size} cfg {cfg} frame {f + 1} {name}")
        assert ea.get_stat(STAT_WAVELET_TILED_LAUNCHES) == 0
        assert eb.get_stat(STAT_WAVELET_TILED_LAUNCHES) == 20, "the tile-staged kernel must be the one that ran"
        assert eb.get_stat(STAT_WAVELET_TILED_ERRORS) == 0


@pytest.mark.parametrize("tiled", [0, 31])
def test_paired_wavelet_records_match_planar(gpu, blue_noise, tiled):
    """ST_OPT_WAVELET_PAIRED: the wide-stride à-trous iterations reading the two signals as interleaved 32-byte records (written by the
    iteration before them, gather or tile-staged) give the same bits as the planar buffers — denoised colours, history, output — on
    frame sizes that are not multiples of the tile and with sky pixels (whose GI half of a record is never written)."""
    from strolle_b200.engine import OPT_WAVELET_TILED, OPT_WAVELET_PAIRED, STAT_WAVELET_TILED_ERRORS
    keep = [n for n in DENOISER_BUFFERS if not n.endswith("_stash")]   # the stash keeps the last planar iteration's output
    for scene in [scenes.cornell(200, 120), scenes.cornell(67, 45), scenes.demo_level(160, 96)]:
        engines = []
        for paired in (0, 1, 2):
            e = gpu.Engine(blue_noise=blue_noise)
            e.set_option(OPT_WAVELET_PAIRED, paired); e.set_option(OPT_WAVELET_TILED, tiled)
            engines.append((e, scenes.apply(e, scene)))
        for f in range(5):
            for e, c in engines:
                e.tick(); e.render_camera(c)
            for name in keep:
                want = engines[0][0].read_buffer(engines[0][1], name)
                for k in (1, 2):
                    assert_bits_equal(engines[k][0].read_buffer(engines[k][1], name), want, f"paired {k} tiled {tiled} frame {f + 1} {name}")
        for e, _ in engines:
            assert e.get_stat(STAT_WAVELET_TILED_ERRORS) == 0


@pytest.mark.parametrize("exact", [True, False])
def test_tiled_variance_matches_gather(gpu, oracle, blue_noise, exact):
    """K21 with its 6x5 window staged in shared memory by TMA == the gather kernel, bit for bit, in both arithmetic flavours;
    frames 1-3 walk the window for every pixel (history < 4), later frames mix both paths; strict mode also == oracle."""
    from strolle_b200.engine import OPT_VARIANCE_TILED, STAT_VARIANCE_TILED_LAUNCHES, STAT_WAVELET_TILED_ERRORS
    for size in [(200, 120), (67, 45)]:
        scene = scenes.cornell(*size)
        ea, eb = gpu.Engine(blue_noise=blue_noise, exact=exact), gpu.Engine(blue_noise=blue_noise, exact=exact)
        ea.set_option(OPT_VARIANCE_TILED, 0); eb.set_option(OPT_VARIANCE_TILED, 1)
        ca, cb = scenes.apply(ea, scene), scenes.apply(eb, scene)
        eo = oracle.OracleEngine(blue_noise=blue_noise) if exact else None
        co = scenes.apply(eo, scene) if exact else None
        for f in range(6):
            ea.tick(); eb.tick(); ea.render_camera(ca); eb.render_camera(cb)
            if exact:
                eo.tick(); eo.render_camera(co)
            for name in DENOISER_BUFFERS:
                assert_bits_equal(eb.read_buffer(cb, name), ea.read_buffer(ca, name), f"{size} frame {f + 1} {name}")
                if exact:
                    assert_bits_equal(eb.read_buffer(cb, name), eo.read_buffer(co, name), f"{size} frame {f + 1} {name} vs oracle")
        assert ea.get_stat(STAT_VARIANCE_TILED_LAUNCHES) == 0 and eb.get_stat(STAT_VARIANCE_TILED_LAUNCHES) == 6
        assert eb.get_stat(STAT_WAVELET_TILED_ERRORS) == 0


def test_tiled_wavelet_against_oracle(gpu, oracle, blue_noise):
    """The tile-staged K22 (mixed tile shapes) against the CPU oracle: bit-exact in strict mode."""
    from strolle_b200.engine import OPT_WAVELET_TILED, OPT_WAVELET_TILE_CFG
    for scene in (scenes.cornell(160, 96), scenes.dungeon(131, 77, cells=6)):
        eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
        eg.set_option(OPT_WAVELET_TILED, 31); eg.set_option(OPT_WAVELET_TILE_CFG, 0x10310)
        run_and_compare(eg, cg, eo, co, 3, buffers=DENOISER_BUFFERS, what="tiled wavelet")


def test_fused_reproject_matches_two_launches(gpu, oracle, blue_noise):
    """K20 for both signals in one launch == the reference's two dispatches, bit for bit (moving camera, so the
    bilinear history path is exercised), and still equal to the oracle."""
    from strolle_b200.engine import OPT_FUSE_REPROJECT
    scene = scenes.cornell(144, 90)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    eg.set_option(OPT_FUSE_REPROJECT, 1)
    e2 = gpu.Engine(blue_noise=blue_noise, exact=True); e2.set_option(OPT_FUSE_REPROJECT, 0)
    c2 = scenes.apply(e2, scene)
    c = scene["camera"]
    for f in range(5):
        t = scenes.look_at_transform((0.02 * f, 1.0 + 0.01 * f, 3.2 - 0.03 * f), (0.0, 1.0, 0.0))
        for e, cam in ((eg, cg), (eo, co), (e2, c2)):
            e.update_camera(cam, c["mode"], c["denoise"], c["ref_depth"], c["w"], c["h"], t, c["projection"])
            e.tick(); e.render_camera(cam)
        assert len(eg.frame_schedule(cg)) == len(e2.frame_schedule(c2)) - 1
        for name in DENOISER_BUFFERS:
            assert_bits_equal(eg.read_buffer(cg, name), e2.read_buffer(c2, name), f"fused vs split frame {f + 1} {name}")
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"fused vs oracle frame {f + 1} {name}")


def test_async_rgba8_output_pipeline(gpu, blue_noise):
    """ST_OPT_ASYNC_OUTPUT: the Rgba8UnormSrgb frame is converted on the engine stream and copied to the caller's pinned buffer on
    a copy stream while the next frame renders; after st_synchronize every buffer holds exactly the frame a blocking read-back
    returns (two host buffers in flight, three frames deep, then a format the async path does not cover)."""
    import torch
    from strolle_b200.engine import OPT_ASYNC_OUTPUT, FORMAT_RGBA8_SRGB, FORMAT_RGBA32F
    w, h = 160, 90
    scene = scenes.cornell(w, h)
    ea, eb = gpu.Engine(blue_noise=blue_noise), gpu.Engine(blue_noise=blue_noise)
    ca, cb = scenes.apply(ea, scene), scenes.apply(eb, scene)
    ea.set_option(OPT_ASYNC_OUTPUT, 1)
    pinned = [torch.zeros((h, w, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    views = [p.numpy() for p in pinned]
    blocking = np.zeros((h, w, 4), dtype=np.uint8)
    for f in range(7):
        ea.tick(); eb.tick()
        ea.render_camera(ca, out=views[f & 1], fmt=FORMAT_RGBA8_SRGB)
        eb.render_camera(cb, out=blocking, fmt=FORMAT_RGBA8_SRGB)
        if f >= 1 and f % 2 == 0:      # buffers are only looked at after a synchronize
            ea.synchronize()
            assert (views[f & 1] == blocking).all(), f"async frame {f + 1} differs from the blocking read-back"
    ea.synchronize()
    assert (views[0] == blocking).all() and blocking[..., :3].max() > 0 and (blocking[..., 3] == 255).all()
    f32a, f32b = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    ea.copy_output(ca, f32a, FORMAT_RGBA32F); eb.copy_output(cb, f32b, FORMAT_RGBA32F)
    ea.synchronize()
    assert_bits_equal(f32a, f32b, "float read-back in async mode")


def test_full_size_properties(gpu, oracle, blue_noise):
    """BASELINE.json's full size (Cornell 1920x1080, configs[1]), checked through properties that do not need the oracle to
    render 2 M pixels per pass: (1) every kernel variant behind an option gives the bits of the plain kernels on all 37
    buffers; (2) the same seeds give the same bits again (determinism); (3) one closest-hit and one any-hit ray per pixel
    (2,073,600 each, both scenes) are bit-identical to the oracle's traversal — hit triangle / material ids included — and the
    Cornell ones to brute force over all triangles; (4) the frame is finite and lit."""
    from strolle_b200.engine import OPT_WAVELET_TILED, OPT_FUSE_REPROJECT, OPT_VARIANCE_TILED, OPT_BVH_REUSE
    w, h = 1920, 1080
    scene = scenes.cornell(w, h)
    plain = gpu.Engine(blue_noise=blue_noise, exact=True)
    for opt in (OPT_WAVELET_TILED, OPT_FUSE_REPROJECT, OPT_VARIANCE_TILED, OPT_BVH_REUSE):
        plain.set_option(opt, 0)
    tuned, again = gpu.Engine(blue_noise=blue_noise, exact=True), gpu.Engine(blue_noise=blue_noise, exact=True)
    cams = [scenes.apply(e, scene) for e in (plain, tuned, again)]
    for f in range(3):
        for e, c in zip((plain, tuned, again), cams):
            e.tick(); e.render_camera(c)
    for name in CAMERA_BUFFERS:
        a = plain.read_buffer(cams[0], name)
        assert_bits_equal(tuned.read_buffer(cams[1], name), a, f"1080p tuned vs plain kernels: {name}")
        assert_bits_equal(again.read_buffer(cams[2], name), tuned.read_buffer(cams[1], name), f"1080p determinism: {name}")
    img = tuned.read_buffer(cams[1], "output").reshape(h, w, 4)
    assert np.isfinite(img).all() and img[..., :3].mean() > 0.01
    n = w * h
    for scene_name in ("cornell", "dungeon"):
        if scene_name == "cornell":
            sc, lo, hi = scenes.cornell(64, 64), (-1.0, 0.0, -1.0), (1.0, 2.0, 3.2)
        else:
            sc, lo, hi = scenes.dungeon(64, 64), (-20.0, 0.1, -40.0), (10.0, 2.9, -5.0)
        eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, sc)
        eg.tick(); eo.tick()
        rays = random_rays(n, 11, lo, hi)
        hg, ho = eg.trace_closest(rays), eo.trace_closest(rays)
        assert (hg[:, 9].view(np.uint32) == ho[:, 9].view(np.uint32)).all() and (hg[:, 10].view(np.uint32) == ho[:, 10].view(np.uint32)).all(), f"{scene_name}: hit ids"
        assert_bits_equal(hg, ho, f"{scene_name}: {n} closest-hit records")
        rays_any = random_rays(n, 12, lo, hi, max_len=4.0)
        assert (eg.trace_any(rays_any) == eo.trace_any(rays_any)).all(), f"{scene_name}: {n} any-hit answers"
        if scene_name == "cornell":
            dist_brute, tri_brute = eo.trace_brute(rays)
            hit = ho[:, 8] < 3e38
            assert (dist_brute[hit].view(np.uint32) == hg[hit, 8].view(np.uint32)).all(), "closest distances == brute force over all triangles"


def test_spot_lights_bit_exact(gpu, oracle, blue_noise):
    """Light::Spot (strolle-gpu/src/light.rs:144-152): the cone factor goes through glam's `angle_between` = acos_approx(...) and
    powf(3); a narrow and a wide spot next to the point light, one of them moved mid-run (light slot update protocol)."""
    scene = scenes.cornell(112, 80)
    h, _, p = scene["lights"][0]
    scene["lights"].append((h + 1, scenes.LIGHT_SPOT, scenes.spot_light((0.3, 1.8, 0.2), 0.1, (8.0, 6.0, 4.0), 20.0, (0.0, -1.0, 0.0), 0.35)))
    scene["lights"].append((h + 2, scenes.LIGHT_SPOT, scenes.spot_light((-0.6, 1.2, 1.5), 0.05, (2.0, 3.0, 5.0), 20.0, (0.4, -0.5, -0.77), 1.1)))
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    for f in range(5):
        if f == 3:
            for e in (eg, eo):
                e.insert_light(h + 1, scenes.LIGHT_SPOT, scenes.spot_light((0.1, 1.7, 0.4), 0.1, (8.0, 6.0, 4.0), 20.0, (0.1, -1.0, 0.1), 0.5))
        eg.tick(); eo.tick()
        assert_bits_equal(eg.read_scene("lights"), eo.read_scene("lights"), f"frame {f + 1} lights")
        eg.render_camera(cg); eo.render_camera(co)
        for name in CAMERA_BUFFERS:
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"spot lights frame {f + 1} {name}")
    di = eo.read_buffer(co, "di_reservoirs_0").reshape(-1, 8)
    ids = di[:, 7].view(np.uint32)
    assert len(set(ids.tolist()) & {2, 3}) > 0, "spot lights were sampled"


# ---- round 2: BASELINE's own configurations compared with the oracle directly -----------------------------------------------------

def test_cornell_1080p_direct_oracle_three_frames(gpu, oracle, blue_noise):
    """BASELINE.json configs[1] at its full size (Cornell 1920x1080, Image{denoise}), strict mode: all 37 per-camera buffers of the
    CUDA path against the oracle's, bit for bit, after each of three frames (frame 2 is a GI sampling frame, frame 3 runs the GI
    spatial passes) — the frame the benchmark times, not a smaller stand-in."""
    oracle.set_threads()
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.cornell(1920, 1080))
    for f in range(3):
        eg.tick(); eo.tick(); eg.render_camera(cg); eo.render_camera(co)
        names = CAMERA_BUFFERS if f == 2 else ["prim_triangle_ids", "di_reservoirs_0", "gi_reservoirs_0", "gi_reservoirs_1", "di_diff_curr_colors", "gi_diff_curr_colors", "output"]
        for name in names:
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"1080p frame {f + 1} {name}")
    img = eg.read_buffer(cg, "output").reshape(1080, 1920, 4)
    assert np.isfinite(img).all() and img[..., :3].mean() > 0.01


def test_demo_level_c3_bit_exact(gpu, oracle, blue_noise):
    """BASELINE config C3 on the reference's own dungeon (bevy-strolle/assets/demo.zip via tools/make_assets.py: 8,393 textured
    triangles + 3 emissive tori, 6 point lights, sun above the horizon): scene upload (materials with atlas rects, triangles, BVH)
    and 7 frames of every per-camera buffer, bit for bit; then the full 1920x1080 frame for two frames."""
    oracle.set_threads()
    scene = scenes.demo_level(256, 144)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    eg.tick(); eo.tick()
    for name in SCENE_BUFFERS:
        assert_bits_equal(eg.read_scene(name), eo.read_scene(name), f"demo level scene:{name}")
    assert eg.bvh_depth() == eo.bvh_depth() and eg.read_scene("triangles").size == 13001 * 36
    eg.render_camera(cg); eo.render_camera(co)
    run_and_compare(eg, cg, eo, co, 6, what="demo level 256x144")
    tid = eo.read_buffer(co, "prim_triangle_ids").reshape(-1, 4)[:, 0].view(np.uint32)
    assert len(np.unique(tid)) > 50 and (tid != 0xffffffff).mean() > 0.9, "the level's geometry is what the camera sees"
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scenes.demo_level(1920, 1080))
    for f in range(2):
        eg.tick(); eo.tick(); eg.render_camera(cg); eo.render_camera(co)
        for name in ["prim_gbuffer_d0_a", "prim_gbuffer_d1_a", "prim_gbuffer_d1_b", "prim_triangle_ids", "di_reservoirs_0", "gi_reservoirs_0", "di_diff_curr_colors", "gi_diff_curr_colors", "output"]:
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"demo level 1080p frame {f + 1} {name}")


@pytest.mark.parametrize("mode,denoise", [(scenes.MODE_DI_DIFFUSE, True), (scenes.MODE_DI_DIFFUSE, False), (scenes.MODE_DI_SPECULAR, True),
                                          (scenes.MODE_GI_DIFFUSE, True), (scenes.MODE_GI_DIFFUSE, False), (scenes.MODE_GI_SPECULAR, True),
                                          (scenes.MODE_IMAGE, False)])
def test_camera_modes_bit_exact(gpu, oracle, blue_noise, mode, denoise):
    """The CameraMode variants besides Image{denoise:true} (strolle/src/camera.rs:83-105): which passes run (needs_di / needs_gi,
    camera_controller.rs:124-160), which buffers frame_composition reads (frame_composition.rs:19-82) and whether the denoiser runs.
    The textured room has a metallic box, so the specular signals are not identically zero."""
    scene = scenes.textured_room(160, 90, mode=mode, denoise=denoise)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    img = run_and_compare(eg, cg, eo, co, 7, what=f"mode {mode} denoise {denoise}")
    assert np.isfinite(img).all() and img.max() > 0.0
    # switching the mode of a live camera re-creates its buffers (CameraController::update -> invalidate, camera_controller.rs:45-63)
    c = scene["camera"]
    for e, cam in ((eg, cg), (eo, co)):
        e.update_camera(cam, scenes.MODE_IMAGE, True, c["ref_depth"], c["w"], c["h"], c["transform"], c["projection"])
    run_and_compare(eg, cg, eo, co, 3, what=f"mode {mode} -> Image")


def test_libm_oracle_within_tolerance_13_frames(gpu, oracle, blue_noise):
    """Against the libm flavour of the oracle (host libm's sin/cos/acos/atan2/exp/pow in place of the shared polynomial kernels —
    the freedom a GLSL.std.450 driver has) the CUDA frame stays inside north_star's tolerance through 13 frames of temporal
    feedback (two GI cycles + 1): per-channel relative L2 <= 1e-3 every frame, primary-hit triangle ids bit-exact, on both scenes."""
    for scene in (scenes.cornell(160, 90), scenes.demo_level(160, 90)):
        eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene, libm=True)
        for f in range(13):
            eg.tick(); eo.tick(); eg.render_camera(cg); eo.render_camera(co)
            a = eg.read_buffer(cg, "output").reshape(-1, 4)[:, :3]
            b = eo.read_buffer(co, "output").reshape(-1, 4)[:, :3]
            for ch in range(3):
                assert rel_l2(a[:, ch], b[:, ch]) <= 1e-3, f"{scene['name']} frame {f + 1} channel {ch}: {rel_l2(a[:, ch], b[:, ch])}"
            assert (eg.read_buffer(cg, "prim_triangle_ids").view(np.uint32) == eo.read_buffer(co, "prim_triangle_ids").view(np.uint32)).all()


def test_empty_scene_and_remove_all_instances(gpu, oracle, blue_noise):
    """A camera rendered before any instance exists (the first frames of an asynchronously loading host) and after the LAST
    instance was removed: the BVH stream is empty (strolle/src/bvh/serializer.rs:20-110 emits nothing for a leaf without
    primitives), every ray misses, the G-buffer is clear and the old geometry is gone — same bits as the oracle throughout,
    including the ray-stream entry points and the modes that trace without a G-buffer."""
    scene = scenes.cornell(96, 64)
    empty = dict(scene, instances=[])
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, empty)
    run_and_compare_allow_dark(eg, cg, eo, co, 2, "no instances yet")
    assert eg.read_scene("bvh").size == 0 and eo.read_scene("bvh").size == 0
    rays = random_rays(4096, 21, (-1.0, 0.0, -1.0), (1.0, 2.0, 3.2))
    assert_bits_equal(eg.trace_closest(rays), eo.trace_closest(rays), "closest hits in an empty scene")
    assert not eg.trace_any(rays).any() and not eo.trace_any(rays).any()
    for e in (eg, eo):                      # the scene arrives ...
        for h, mesh, mat, xf in scene["instances"]:
            e.insert_instance(h, mesh, mat, xf)
    run_and_compare(eg, cg, eo, co, 3, what="instances arrived")
    for e in (eg, eo):                      # ... and leaves again, instance by instance
        for h, _, _, _ in scene["instances"]:
            e.remove_instance(h)
    run_and_compare_allow_dark(eg, cg, eo, co, 3, "all instances removed")
    assert eg.read_scene("bvh").size == 0
    tid = eg.read_buffer(cg, "prim_triangle_ids").reshape(-1, 4)[:, 0].view(np.uint32)
    assert (tid == 0xffffffff).all(), "no pixel still sees the removed geometry"
    for mode in (scenes.MODE_REFERENCE, scenes.MODE_BVH_HEATMAP):
        e2, c2, o2, d2 = make_pair(gpu, oracle, blue_noise, dict(scenes.cornell(64, 48, mode=mode), instances=[]))
        run_and_compare_allow_dark(e2, c2, o2, d2, 2, f"empty scene, mode {mode}", buffers=["ref_hits", "ref_rays", "ref_colors", "output"])


def run_and_compare_allow_dark(eg, cg, eo, co, frames, what, buffers=CAMERA_BUFFERS):
    for f in range(frames):
        eg.tick(); eo.tick(); eg.render_camera(cg); eo.render_camera(co)
        for name in buffers:
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"{what} frame {f + 1} {name}")


def test_multi_gpu_transports_reproduce_single_gpu_frame():
    """On a box with >= 2 GPUs: tools/verify_multigpu.py under torchrun (one process per GPU) — every halo transport and the
    sample-parallel reference mode against the single-GPU frame.  Skipped on single-GPU boxes."""
    import os, subprocess, sys, socket
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tools", "verify_multigpu.py")], capture_output=True, text=True, timeout=900, cwd=root)
    lines = [l for l in p.stdout.splitlines() if l.startswith(("OK", "FAIL"))]
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert lines and all(l.startswith("OK") for l in lines), "\n".join(lines)


EXACT_IN_FAST_MODE = ["prim_gbuffer_d0_a", "prim_gbuffer_d0_b", "prim_gbuffer_d1_a", "prim_gbuffer_d1_b", "prim_surface_map_a", "prim_surface_map_b",
                      "reprojection_map", "velocity_map", "prim_triangle_ids"]


@pytest.mark.parametrize("scene_name", ["cornell", "demo_level", "textured_room"])
def test_fast_shading_mode_within_tolerance(gpu, oracle, blue_noise, scene_name):
    """The product default (ST_OPT_SHADING_FAST_MATH + ST_OPT_SVGF_FAST_MATH): ReSTIR radiance / BRDF / pdf / MIS arithmetic with FMA
    contraction and SFU approximations.  Over 13 frames of temporal feedback (two GI cycles + 1) against the strict oracle:
    primary-hit triangle ids, G-buffer, surface, velocity and reprojection maps stay bit-exact (traversal and the primary pass are
    the same code in both builds); the composed frame and the denoised DI/GI signals stay inside north_star's 1e-3 relative
    per-channel L2; the frame's energy matches.  (The textured room is there for its metallic box: GGX's D term
    `(n.h * a2 - n.h) * n.h + 1` (brdf.rs:20-24) cancels catastrophically near the highlight, so ANY change in rounding — FMA
    contraction included, on the reference's own GPU path as well — moves those few pixels by percents; its bound is 2e-2.)"""
    tol = 2e-2 if scene_name == "textured_room" else 1e-3
    scene = {"cornell": scenes.cornell, "demo_level": scenes.demo_level, "textured_room": scenes.textured_room}[scene_name](224, 126)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene, exact=False)
    worst = 0.0
    for f in range(13):
        eg.tick(); eo.tick(); eg.render_camera(cg); eo.render_camera(co)
        for name in EXACT_IN_FAST_MODE:
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"fast shading must leave {name} bit-exact (frame {f + 1})")
        for name in ["output", "di_diff_curr_colors", "gi_diff_curr_colors"]:
            a = eg.read_buffer(cg, name).reshape(-1, 4)[:, :3]; b = eo.read_buffer(co, name).reshape(-1, 4)[:, :3]
            for ch in range(3):
                err = rel_l2(a[:, ch], b[:, ch]); worst = max(worst, err)
                assert err <= tol, f"{scene_name} frame {f + 1} {name} channel {ch}: rel L2 {err:.2e}"
    a = eg.read_buffer(cg, "output").reshape(-1, 4)[:, :3]; b = eo.read_buffer(co, "output").reshape(-1, 4)[:, :3]
    assert abs(float(a.mean()) / float(b.mean()) - 1.0) <= tol, "frame energy"
    print(f"fast shading {scene_name}: worst per-channel rel L2 over 13 frames = {worst:.2e}")


def test_fast_shading_traversal_is_the_exact_traversal(gpu, blue_noise):
    """Both builds of the ReSTIR kernels walk the BVH with the same flag-independent arithmetic: fed the SAME rays (the spatial
    visibility pass K8 reads its rays from buffers), the fast build returns the same visibility bits as the strict one."""
    from strolle_b200.engine import OPT_SHADING_FAST_MATH
    scene = scenes.demo_level(320, 180)
    ea, eb = gpu.Engine(blue_noise=blue_noise, exact=True), gpu.Engine(blue_noise=blue_noise, exact=True)
    ca, cb = scenes.apply(ea, scene), scenes.apply(eb, scene)
    from strolle_b200 import multigpu as mg
    for f in range(3):
        ea.tick(); eb.tick()
        sched = ea.frame_schedule(ca)
        k = sched.index(mg.P_DI_SPATIAL_TRACE)
        ea.render_range(ca, 0, k); eb.render_range(cb, 0, k - 1)
        eb.set_option(OPT_SHADING_FAST_MATH, 1); eb.render_range(cb, k, k); eb.set_option(OPT_SHADING_FAST_MATH, 0)   # only K8 from the fast build
        assert_bits_equal(ea.read_buffer(ca, "di_diff_stash"), eb.read_buffer(cb, "di_diff_stash"), f"frame {f + 1}: K8 visibility, fast build vs strict build on identical rays")
        ea.render_range(ca, k + 1, len(sched) - 1); eb.render_range(cb, k + 1, len(sched) - 1)
    vis = ea.read_buffer(ca, "di_reservoirs_0").reshape(-1, 8)[:, 3].view(np.uint32) & 0xff
    assert 0.02 < (vis > 0).mean() < 0.98


# ---- row strips across devices of one process (st_multi_*, fused transport) --------------------------------------------------------

def _devices(n):
    import torch
    have = max(torch.cuda.device_count(), 1)
    return [k % have for k in range(n)]   # a single-GPU box runs every strip on device 0: same protocol, same kernels


@pytest.mark.parametrize("n,size,exact,fused", [(2, (320, 288), True, True), (3, (256, 400), True, True), (2, (320, 288), False, True),
                                                (4, (200, 520), False, True), (2, (320, 288), True, False), (3, (200, 400), False, "mirror"),
                                                (3, (128, 720), False, True), (2, (320, 288), False, "dma3"), (3, (256, 400), True, "dma3")])   # 720 rows over 3: outer strips 252 rows, inner 216 (weighted partition)
def test_multi_device_group_matches_single_gpu(gpu, blue_noise, n, size, exact, fused):
    """st_multi_* (one process, n devices): the frame rendered as n row strips — producer kernels mirroring their
    boundary rows into the neighbours, neighbour-only sequence flags, G-buffer / SVGF halo rows recomputed, temporal rows pulled on
    demand — is the single-GPU frame, bit for bit, in every per-camera buffer, through two GI cycles with the camera first still,
    then drifting, then jumping by more rows than any fixed temporal halo would cover."""
    from strolle_b200.engine import OPT_STRIP_FUSED, OPT_STRIP_DMA, FORMAT_RGBA8_SRGB
    w, h = size
    scene = scenes.cornell(w, h)
    one = gpu.Engine(blue_noise=blue_noise, exact=exact)
    grp = gpu.MultiEngine(_devices(n), blue_noise=blue_noise, exact=exact)
    grp.set_option(OPT_STRIP_FUSED, int(bool(fused)))
    # ST_OPT_STRIP_DMA: "mirror" = every halo by in-kernel stores, G-buffer rows recomputed; "dma3" = every halo with slack by copy engine;
    # default = GI halos by copy engine, and from three strips on the G-buffer rows too
    if fused in ("mirror", "dma3"):
        grp.set_option(OPT_STRIP_DMA, 0 if fused == "mirror" else 3)
    c1, cn = scenes.apply(one, scene), scenes.apply(grp, scene)
    c = scene["camera"]
    for f in range(13):
        eye = (0.0, 1.0, 3.2)
        if 4 <= f < 9:
            eye = (0.01 * (f - 3), 1.0 + 0.03 * (f - 3), 3.2)           # drifting up: reprojection crosses strip edges
        if f >= 9:
            eye = (0.05, 1.0 + 0.15 + (0.35 if f % 2 else 0.0), 3.15)   # jumping up and down by tens of rows
        t = scenes.look_at_transform(eye, (0.0, 1.0 + (eye[1] - 1.0), 0.0))
        for e, cam in ((one, c1), (grp, cn)):
            e.update_camera(cam, c["mode"], c["denoise"], c["ref_depth"], w, h, t, c["projection"])
            e.tick(); e.render_camera(cam)
        for name in CAMERA_BUFFERS:
            assert_bits_equal(grp.read_buffer(cn, name), one.read_buffer(c1, name), f"{n} strips frame {f + 1} {name}")
    assert grp.peer_errors(cn) == 0
    vel = one.read_buffer(c1, "velocity_map").reshape(h, w, 4)
    assert np.abs(vel[..., 1]).max() > 16.0, "the jump moved pixels by more than 16 rows"
    if fused:
        from strolle_b200.engine import STAT_STRIP_PULLED_ROWS, STAT_LAST_FRAME_FUSED_STRIPS
        assert all(grp.member(r).get_stat(STAT_LAST_FRAME_FUSED_STRIPS) == 1 for r in range(n)), "the fused transport is what ran"
        assert sum(grp.member(r).get_stat(STAT_STRIP_PULLED_ROWS) for r in range(n)) > 0, "the moving frames pulled rows from their owners"
    a, b = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 4), np.uint8)
    one.tick(); grp.tick()
    one.render_camera(c1, a, FORMAT_RGBA8_SRGB); grp.render_camera(cn, b, FORMAT_RGBA8_SRGB)
    assert (a == b).all() and a[..., :3].max() > 0


def test_multi_device_group_demo_level_and_modes(gpu, blue_noise):
    """The same on the reference's dungeon (textures, atmosphere, 6 lights) and for a DI-only / GI-only camera mode (chains that are
    interleaved in Image mode run alone there)."""
    for scene in (scenes.demo_level(288, 300), scenes.cornell(256, 272, mode=scenes.MODE_DI_DIFFUSE), scenes.cornell(256, 272, mode=scenes.MODE_GI_DIFFUSE, denoise=False)):
        one = gpu.Engine(blue_noise=blue_noise)
        grp = gpu.MultiEngine(_devices(2), blue_noise=blue_noise)
        c1, cn = scenes.apply(one, scene), scenes.apply(grp, scene)
        for f in range(7):
            one.tick(); grp.tick(); one.render_camera(c1); grp.render_camera(cn)
            for name in ["output", "di_reservoirs_0", "gi_reservoirs_0", "gi_reservoirs_3", "di_diff_curr_colors", "gi_diff_curr_colors", "di_diff_prev_colors", "gi_diff_moments_a"]:
                assert_bits_equal(grp.read_buffer(cn, name), one.read_buffer(c1, name), f"{scene['name']} mode {scene['camera']['mode']} frame {f + 1} {name}")
        assert grp.peer_errors(cn) == 0


NOT_WRITTEN_WHEN_FUSED = {"gi_d0", "gi_d1", "gi_d2", "gi_reservoirs_2"}   # scratch between fused members / K11's entry when it runs inside K14


@pytest.mark.parametrize("scene_name,size", [("cornell", (200, 120)), ("cornell", (121, 67)), ("textured_room", (192, 108)), ("demo_level", (176, 99))])
def test_fused_passes_bit_exact(gpu, oracle, blue_noise, scene_name, size):
    """ST_OPT_FUSED_PASSES with strict arithmetic: K5+K6, K7+K8+K9, K12+K13, K11-in-K14, K15+K16+K17 and preview#2+K19 as single launches
    leave every reservoir, sample, colour, moment and the composed frame bit-identical to the oracle's one-dispatch-per-pass frame, over
    two GI cycles with a moving camera (widths 200 and 121 have columns the checkerboard passes do not cover; the textured room has
    alpha-tested and metallic surfaces)."""
    from strolle_b200.engine import OPT_FUSED_PASSES
    scene = {"cornell": scenes.cornell, "demo_level": scenes.demo_level, "textured_room": scenes.textured_room}[scene_name](*size)
    eg, cg, eo, co = make_pair(gpu, oracle, blue_noise, scene)
    eg.set_option(OPT_FUSED_PASSES, 1)
    c = scene["camera"]
    base = np.asarray(c["transform"], np.float32).copy()
    names = [n for n in CAMERA_BUFFERS if n not in NOT_WRITTEN_WHEN_FUSED]
    for f in range(13):
        t = base.copy()
        if f >= 5:
            t[12] += 0.01 * (f - 4); t[13] += 0.006 * (f - 4)   # translate the eye
        for e, cam in ((eg, cg), (eo, co)):
            e.update_camera(cam, c["mode"], c["denoise"], c["ref_depth"], c["w"], c["h"], t, c["projection"])
            e.tick(); e.render_camera(cam)
        assert len(eg.frame_schedule(cg)) <= 18, "fused schedule"
        for name in names:
            assert_bits_equal(eg.read_buffer(cg, name), eo.read_buffer(co, name), f"fused passes {scene_name} {size} frame {f + 1} {name}")
