"""Scene descriptions (input data) for the benchmark configurations.

A scene is plain data: meshes, materials, instances, lights, sun, camera — the
arguments a host application would pass through the Engine API
(strolle/src/lib.rs:161-245).  `apply(engine, scene)` drives any object exposing
that API (the CUDA engine in strolle_b200.engine, or the test oracle).
"""
import json
import math
import os

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

MODE_IMAGE, MODE_DI_DIFFUSE, MODE_DI_SPECULAR, MODE_GI_DIFFUSE, MODE_GI_SPECULAR, MODE_BVH_HEATMAP, MODE_REFERENCE = range(7)
LIGHT_POINT, LIGHT_SPOT = 1, 2


def blue_noise():
    """256x256 RGBA8 blue-noise tile (strolle/assets/blue-noise.png as raw bytes)."""
    return np.fromfile(os.path.join(_ASSETS, "blue_noise_256_rgba8.bin"), dtype=np.uint8).reshape(256, 256, 4)


def perspective_infinite_reverse_rh(fov_y, aspect, near):
    """glam Mat4::perspective_infinite_reverse_rh, column-major 16 floats (Bevy's default projection)."""
    f = np.float32(1.0) / np.float32(math.tan(0.5 * fov_y))
    m = np.zeros((4, 4), dtype=np.float32)  # m[col][row]
    m[0][0] = f / np.float32(aspect)
    m[1][1] = f
    m[2][3] = -1.0
    m[3][2] = near
    return m.reshape(-1)


def look_at_transform(eye, target, up=(0.0, 1.0, 0.0)):
    """Bevy Transform::from_translation(eye).looking_at(target, up) as a column-major Mat4."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    back = -fwd
    right = np.cross(np.asarray(up, dtype=np.float64), back)
    right /= np.linalg.norm(right)
    upv = np.cross(back, right)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0][:3] = right
    m[1][:3] = upv
    m[2][:3] = back
    m[3][:3] = eye
    m[3][3] = 1.0
    return m.reshape(-1)


def material(base_color, emissive=(0, 0, 0, 0), perceptual_roughness=1.0, metallic=0.0, reflectance=0.5, ior=1.0):
    return np.array(list(base_color) + list(emissive) + [perceptual_roughness, metallic, reflectance, ior], dtype=np.float32)


def point_light(position, radius, color, rng):
    return np.array(list(position) + [radius] + list(color) + [rng, 0, 0, 0, 0], dtype=np.float32)


def spot_light(position, radius, color, rng, direction, angle):
    """Light::Spot (strolle/src/light.rs:14-22): a point light restricted to a cone of half-angle `angle` around `direction`."""
    return np.array(list(position) + [radius] + list(color) + [rng] + list(direction) + [angle], dtype=np.float32)


def tri36(positions, normals, uvs=None, tangents=None):
    uvs = uvs if uvs is not None else [[0, 0]] * 3
    tangents = tangents if tangents is not None else [[0, 0, 0, 0]] * 3
    return np.concatenate([np.asarray(positions, np.float32).reshape(-1), np.asarray(normals, np.float32).reshape(-1),
                           np.asarray(uvs, np.float32).reshape(-1), np.asarray(tangents, np.float32).reshape(-1)])


def affine_from_colmajor4x4(m16):
    m = np.asarray(m16, dtype=np.float32).reshape(4, 4)  # m[col][row]
    return np.concatenate([m[0][:3], m[1][:3], m[2][:3], m[3][:3]]).astype(np.float32)


IDENTITY_AFFINE = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float32)


def cornell(width=1920, height=1080, mode=MODE_IMAGE, denoise=True, ref_depth=1):
    """Config C1/C2/C4/C5: Cornell box, 32 triangles, one point light.

    bevy-strolle/examples/cornell.rs:39-94: camera eye (0,1,3.2) -> (0,1,0), point light
    (0,1.5,0.5) r=0.15 range 20 intensity 50, sun altitude -1.
    """
    doc = json.load(open(os.path.join(_ASSETS, "cornell.json")))
    meshes, materials, instances = {}, {}, []
    for i, m in enumerate(doc["materials"]):
        materials[100 + i] = (material(m["base_color"], perceptual_roughness=m["perceptual_roughness"], metallic=m["metallic"]), False)
    for i, m in enumerate(doc["meshes"]):
        tris = np.stack([tri36(t["positions"], t["normals"]) for t in m["triangles"]])
        meshes[200 + i] = tris
        instances.append((300 + i, 200 + i, 100 + m["material"], affine_from_colmajor4x4(m["transform_colmajor"])))
    intensity = 50.0 / (4.0 * math.pi)
    lights = [(400, LIGHT_POINT, point_light((0.0, 1.5, 0.5), 0.15, (intensity,) * 3, 20.0))]
    cam = dict(mode=mode, denoise=denoise, ref_depth=ref_depth, w=width, h=height,
               transform=look_at_transform((0.0, 1.0, 3.2), (0.0, 1.0, 0.0)),
               projection=perspective_infinite_reverse_rh(math.pi / 4.0, width / height, 0.1))
    return dict(name="cornell", meshes=meshes, materials=materials, instances=instances, lights=lights, sun=(0.0, -1.0), camera=cam)


def cornell_spots(width=160, height=120, **kw):
    """Cornell box with two Light::Spot lights next to its point light (a narrow one pointing down, a wide oblique one)."""
    scene = cornell(width, height, **kw)
    h = scene["lights"][0][0]
    scene["lights"].append((h + 1, LIGHT_SPOT, spot_light((0.3, 1.8, 0.2), 0.1, (8.0, 6.0, 4.0), 20.0, (0.0, -1.0, 0.0), 0.35)))
    scene["lights"].append((h + 2, LIGHT_SPOT, spot_light((-0.6, 1.2, 1.5), 0.05, (2.0, 3.0, 5.0), 20.0, (0.4, -0.5, -0.77), 1.1)))
    return scene


def _box(lo, hi):
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    c = [np.array([x, y, z], np.float32) for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])]
    faces = [((0, 1, 3, 2), (-1, 0, 0)), ((4, 6, 7, 5), (1, 0, 0)), ((0, 4, 5, 1), (0, -1, 0)), ((2, 3, 7, 6), (0, 1, 0)),
             ((0, 2, 6, 4), (0, 0, -1)), ((1, 5, 7, 3), (0, 0, 1))]
    tris = []
    for (a, b, cc, d), n in faces:
        tris.append(tri36([c[a], c[b], c[cc]], [n] * 3, [[0, 0], [1, 0], [1, 1]]))
        tris.append(tri36([c[a], c[cc], c[d]], [n] * 3, [[0, 0], [1, 1], [0, 1]]))
    return tris


def _torus(major=0.5, minor=0.25, nu=24, nv=12):
    tris = []
    def pt(i, j):
        u, v = 2 * math.pi * i / nu, 2 * math.pi * j / nv
        cx, cz = math.cos(u), math.sin(u)
        p = np.array([(major + minor * math.cos(v)) * cx, minor * math.sin(v), (major + minor * math.cos(v)) * cz], np.float32)
        n = np.array([math.cos(v) * cx, math.sin(v), math.cos(v) * cz], np.float32)
        return p, n
    for i in range(nu):
        for j in range(nv):
            (p00, n00), (p10, n10), (p01, n01), (p11, n11) = pt(i, j), pt(i + 1, j), pt(i, j + 1), pt(i + 1, j + 1)
            tris.append(tri36([p00, p10, p11], [n00, n10, n11]))
            tris.append(tri36([p00, p11, p01], [n00, n11, n01]))
    return tris


def dungeon(width=1920, height=1080, mode=MODE_IMAGE, denoise=True, seed=7, cells=13):
    """A *synthetic* dungeon (small parity cases and golden fixtures; the benchmark's C3 is `demo_level` below).

    The reference's demo level (bevy-strolle/assets/demo.zip, 8,393 triangles + 3 emissive tori, 6 point
    lights, sun az 3.0 / alt 0.35, bevy-strolle/examples/demo.rs:150-237) needs the texture atlas (SURVEY
    §8f-3, a "next" row); this generator builds an untextured level of the same scale procedurally: a grid of
    4 m cells with floor/ceiling slabs, random partition walls, pillars and crates (boxes), an open corridor in
    front of the camera, 3 emissive tori, 6 point lights, the same sun, the same camera pose
    (eye (-5.75, 0.5, -16.8) looking down -Z).  Some ceiling slabs are missing so that the sky is visible.
    """
    rng = np.random.RandomState(seed)
    meshes, materials, instances, lights = {}, {}, [], []
    palette = [(0.55, 0.5, 0.45, 1), (0.4, 0.38, 0.36, 1), (0.5, 0.3, 0.2, 1), (0.3, 0.35, 0.4, 1), (0.6, 0.55, 0.4, 1)]
    for i, col in enumerate(palette):
        materials[100 + i] = (material(col, perceptual_roughness=1.0, reflectance=0.0), False)
    materials[150] = (material((0.9, 0.6, 0.3, 1), emissive=(9.0, 6.0, 3.0, 1.0), perceptual_roughness=1.0, reflectance=0.0), False)
    nid = [0]

    def add(tris, mat, xf=IDENTITY_AFFINE):
        h = nid[0]
        nid[0] += 1
        meshes[1000 + h] = np.stack(tris)
        instances.append((5000 + h, 1000 + h, mat, np.asarray(xf, np.float32)))

    S, wall_h = 4.0, 3.0
    eye = np.array([-5.75, 0.5, -16.8])
    half = cells // 2
    # cell (i, j) spans x in [cx(i), cx(i)+S), z in [cz(j), cz(j)+S); the camera sits in the middle of cell (half, half)
    cx = lambda i: eye[0] - S / 2 + (i - half) * S
    cz = lambda j: eye[2] - S / 2 + (j - half) * S
    corridor = {(half, j) for j in range(max(half - 5, 0), half + 1)}
    for i in range(cells):
        for j in range(cells):
            x0, z0 = cx(i), cz(j)
            add(_box((x0, -0.2, z0), (x0 + S, 0.0, z0 + S)), 100 + (i + j) % 2)
            if rng.rand() < 0.85:
                add(_box((x0, wall_h, z0), (x0 + S, wall_h + 0.2, z0 + S)), 101)
            # partition walls on the -z and -x borders, never across the corridor
            if (i, j) not in corridor or (i, j - 1) not in corridor:
                if rng.rand() < 0.4 and not ((i, j) in corridor and (i, j - 1) in corridor):
                    add(_box((x0, 0.0, z0), (x0 + S, wall_h, z0 + 0.3)), 102 + rng.randint(0, 3))
            if (i, j) not in corridor and (i - 1, j) not in corridor and rng.rand() < 0.4:
                add(_box((x0, 0.0, z0), (x0 + 0.3, wall_h, z0 + S)), 102 + rng.randint(0, 3))
            if (i, j) in corridor:
                for side in (-1, 1):   # crates along the corridor sides
                    if rng.rand() < 0.6:
                        px, pz = x0 + S / 2 + side * 1.5, z0 + rng.rand() * (S - 1) + 0.5
                        hgt = 0.3 + rng.rand() * 0.9
                        add(_box((px - 0.25, 0.0, pz - 0.25), (px + 0.25, hgt, pz + 0.25)), 102 + rng.randint(0, 3))
                continue
            for _ in range(rng.randint(0, 4)):
                px, pz = x0 + rng.rand() * (S - 1) + 0.5, z0 + rng.rand() * (S - 1) + 0.5
                hgt = 0.3 + rng.rand() * 1.2
                add(_box((px - 0.25, 0.0, pz - 0.25), (px + 0.25, hgt, pz + 0.25)), 102 + rng.randint(0, 3))
            if rng.rand() < 0.3:
                px, pz = x0 + S / 2, z0 + S / 2
                add(_box((px - 0.2, 0.0, pz - 0.2), (px + 0.2, wall_h, pz + 0.2)), 103)
    lo, hi = cx(0), cx(cells)
    zlo, zhi = cz(0), cz(cells)
    add(_box((lo - 0.3, 0.0, zlo), (lo, wall_h, zhi)), 104); add(_box((hi, 0.0, zlo), (hi + 0.3, wall_h, zhi)), 104)
    add(_box((lo, 0.0, zlo - 0.3), (hi, wall_h, zlo)), 104); add(_box((lo, 0.0, zhi), (hi, wall_h, zhi + 0.3)), 104)
    torus = _torus()
    for k, (dx, dz) in enumerate([(0.0, -6.0), (4.0, -9.0), (-4.0, -12.0)]):
        xf = np.array([0.5, 0, 0, 0, 0.5, 0, 0, 0, 0.5, eye[0] + dx, 1.2, eye[2] + dz], np.float32)
        add(torus, 150, xf)
    inten = 120.0 / (4.0 * math.pi)
    for k, (dx, dz) in enumerate([(0.0, -3.0), (4.5, -7.0), (-4.5, -7.0), (0.0, -13.0), (8.0, -2.0), (-8.0, -2.0)]):
        lights.append((9000 + k, LIGHT_POINT, point_light((eye[0] + dx, 2.4, eye[2] + dz), 0.15, (inten,) * 3, 35.0)))
    cam = dict(mode=mode, denoise=denoise, ref_depth=1, w=width, h=height,
               transform=look_at_transform(tuple(eye), (eye[0], eye[1], eye[2] - 0.2)),
               projection=perspective_infinite_reverse_rh(math.pi / 4.0, width / height, 0.1))
    return dict(name="dungeon_synthetic", meshes=meshes, materials=materials, instances=instances, lights=lights, sun=(3.0, 0.35), camera=cam)


def _bevy_torus(radius=1.0, ring_radius=0.5, segments=32, sides=24):
    """bevy 0.12.1 `shape::Torus::default()` (third-party crate pinned in Cargo.toml:16, not vendored under /root/reference; restated
    from its published mesh generator): (segments+1) x (sides+1) vertices, two triangles (lt, rt, lb), (rt, rb, lb) per face."""
    pos, nor, uvs = [], [], []
    seg_stride, side_stride = np.float32(2.0 * math.pi) / np.float32(segments), np.float32(2.0 * math.pi) / np.float32(sides)
    for seg in range(segments + 1):
        theta = float(seg_stride * np.float32(seg))
        for side in range(sides + 1):
            phi = float(side_stride * np.float32(side))
            p = np.array([math.cos(theta) * (radius + ring_radius * math.cos(phi)), ring_radius * math.sin(phi),
                          math.sin(theta) * (radius + ring_radius * math.cos(phi))], np.float64)
            c = np.array([radius * math.cos(theta), 0.0, radius * math.sin(theta)], np.float64)
            n = (p - c) / np.linalg.norm(p - c)
            pos.append(p.astype(np.float32)); nor.append(n.astype(np.float32)); uvs.append([seg / segments, side / sides])
    tris = []
    row = sides + 1
    for seg in range(segments):
        for side in range(sides):
            lt, rt, lb, rb = side + seg * row, side + 1 + seg * row, side + (seg + 1) * row, side + 1 + (seg + 1) * row
            for a, b, c in ((lt, rt, lb), (rt, rb, lb)):
                tris.append(tri36([pos[a], pos[b], pos[c]], [nor[a], nor[b], nor[c]], [uvs[a], uvs[b], uvs[c]]))
    return tris


def demo_level(width=1920, height=1080, mode=MODE_IMAGE, denoise=True, textures=True):
    """BASELINE config C3: the reference's dungeon demo (bevy-strolle/examples/demo.rs).

    Geometry, materials and textures come from the reference's own asset (assets/demo.zip -> demo/level.glb, extracted
    by tools/make_assets.py into assets/dungeon.npz: 45 meshes, 8,393 triangles, 45 64x64 base-colour textures); the
    rest follows demo.rs: every material re-lit with reflectance 0 / perceptual_roughness 1 (`adjust_materials`,
    demo.rs:246-261 — it walks ALL StandardMaterials, the tori's included), three emissive tori
    (shape::Torus::default(), rotation_z(1.0), scale 0.5, emissive 10 x (0.9, 0.6, 0.3), demo.rs:195-219), six point
    lights of intensity 5000 (x 1/4pi in bevy-strolle/src/stages/extract.rs:285), range 35, radius 0.15 (demo.rs:169-191), the
    flashlight dropped for its zero intensity (extract.rs:306-311), camera eye (-5.75, 0.5, -16.8) -> (-5.75, 0.5, -17.0)
    (demo.rs:150-152), sun azimuth 3.0 (_common.rs:166-173) and strolle::Sun's default altitude 0.35 (strolle/src/sun.rs:7-13).
    `textures=False` is demo.rs's T key (base_color_texture = None on every material)."""
    d = np.load(os.path.join(_ASSETS, "dungeon.npz"))
    meshes, materials, instances, lights, images, material_textures = {}, {}, [], [], {}, {}
    nmesh = len(d["mesh_material"])
    zeros_t = np.zeros((3, 4), np.float32)
    for k in range(nmesh):
        pos, nor, uv = d[f"pos{k}"], d[f"nor{k}"], d[f"uv{k}"]
        tris = np.concatenate([pos.reshape(-1, 9), nor.reshape(-1, 9), uv.reshape(-1, 6), np.tile(zeros_t.reshape(1, 12), (len(pos), 1))], axis=1).astype(np.float32)
        meshes[1000 + k] = tris
        instances.append((5000 + k, 1000 + k, 100 + int(d["mesh_material"][k]), affine_from_colmajor4x4(d["mesh_transform_colmajor"][k])))
    for i, bmr in enumerate(d["material_base_metallic_roughness"]):
        # bevy_gltf StandardMaterial (base colour factor, metallic factor) then demo.rs::adjust_materials: reflectance 0, roughness 1
        materials[100 + i] = (material(tuple(float(v) for v in bmr[:4]), perceptual_roughness=1.0, metallic=float(bmr[4]), reflectance=0.0), False)
        t = int(d["material_texture"][i])
        if textures and t >= 0:
            material_textures[100 + i] = dict(base_color=700 + t)
    if textures:
        for t, img in enumerate(d["images_rgba8"]):
            images[700 + t] = np.ascontiguousarray(img)
    torus = np.stack(_bevy_torus())
    meshes[2000] = torus
    c1, s1 = math.cos(1.0), math.sin(1.0)   # Quat::from_rotation_z(1.0) * scale 0.5
    for k, (tx, ty, tz) in enumerate([(-0.5, 0.33, -5.5), (-11.0, 0.33, 28.0), (-11.5, 0.33, 13.5)]):
        materials[160 + k] = (material((0.9, 0.6, 0.3, 1.0), emissive=(9.0, 6.0, 3.0, 1.0), perceptual_roughness=1.0, reflectance=0.0), False)
        xf = np.array([0.5 * c1, 0.5 * s1, 0.0, -0.5 * s1, 0.5 * c1, 0.0, 0.0, 0.0, 0.5, tx, ty, tz], np.float32)
        instances.append((6000 + k, 2000, 160 + k, xf))
    inten = 5000.0 / (4.0 * math.pi)
    for k, p in enumerate([(-3.0, 0.75, -23.0), (-23.5, 0.75, -31.0), (1.25, 0.75, -10.5), (-3.15, 0.75, 1.25), (-3.25, 0.75, 20.25), (13.25, 0.75, -28.25)]):
        lights.append((9000 + k, LIGHT_POINT, point_light(p, 0.15, (inten,) * 3, 35.0)))
    cam = dict(mode=mode, denoise=denoise, ref_depth=1, w=width, h=height,
               transform=look_at_transform((-5.75, 0.5, -16.8), (-5.75, 0.5, -17.0)),
               projection=perspective_infinite_reverse_rh(math.pi / 4.0, width / height, 0.1))
    return dict(name="dungeon_demo_level", meshes=meshes, materials=materials, instances=instances, lights=lights, sun=(3.0, 0.35), camera=cam,
                images=images, material_textures=material_textures)


def _quad(p0, p1, p2, p3, normal, uv_lo=(0.0, 0.0), uv_hi=(1.0, 1.0)):
    """Two triangles wound so that the geometric normal agrees with `normal` (Triangle::hit flips the shading normal by
    the sign of the determinant, strolle-gpu/src/triangle.rs:95-101, i.e. it trusts the winding)."""
    (u0, v0), (u1, v1) = uv_lo, uv_hi
    c = [np.asarray(p, np.float64) for p in (p0, p1, p2, p3)]
    uv = [[u0, v0], [u1, v0], [u1, v1], [u0, v1]]
    if np.dot(np.cross(c[1] - c[0], c[2] - c[0]), np.asarray(normal, np.float64)) < 0:
        c = [c[0], c[3], c[2], c[1]]
        uv = [uv[0], uv[3], uv[2], uv[1]]
    return [tri36([c[0], c[1], c[2]], [normal] * 3, [uv[0], uv[1], uv[2]]), tri36([c[0], c[2], c[3]], [normal] * 3, [uv[0], uv[2], uv[3]])]


def textured_room(width=320, height=180, mode=MODE_IMAGE, denoise=True):
    """Exercises the texture atlas: sRGB base-colour textures with repeat-wrapped (also negative) uvs,
    an emissive texture, a metallic-roughness texture, and an alpha-cutout AlphaMode::Blend fence whose holes must let
    primary, shadow and bounce rays through (strolle-gpu/src/ray.rs:212-229, material.rs:76-104)."""
    rng = np.random.RandomState(3)
    def checker(n, cell, a, b):
        img = np.zeros((n, n, 4), np.uint8)
        yy, xx = np.mgrid[0:n, 0:n]
        m = ((xx // cell) + (yy // cell)) % 2 == 0
        img[m] = a; img[~m] = b
        return img
    images = {
        700: checker(64, 8, (200, 180, 150, 255), (90, 60, 40, 255)),                      # floor base colour
        701: checker(32, 4, (255, 255, 255, 255), (0, 0, 0, 0)),                             # fence: alpha cutout
        702: (rng.randint(0, 256, size=(16, 48, 4))).astype(np.uint8),                       # emissive noise (non-square)
        703: checker(16, 2, (255, 64, 255, 255), (255, 255, 32, 255)),                       # metallic-roughness (g = roughness, b = metallic)
    }
    images[702][..., 3] = 255
    materials = {
        100: (material((1.0, 1.0, 1.0, 1.0)), False), 101: (material((0.7, 0.7, 0.75, 1.0)), False),
        102: (material((1.0, 1.0, 1.0, 1.0)), True),                                           # fence, AlphaMode::Blend
        103: (material((0.2, 0.2, 0.2, 1.0), emissive=(3.0, 2.0, 1.0, 1.0)), False),
        104: (material((0.9, 0.8, 0.6, 1.0), perceptual_roughness=0.8, metallic=1.0), False),
    }
    material_textures = {100: dict(base_color=700), 102: dict(base_color=701), 103: dict(emissive=702), 104: dict(metallic_roughness=703, base_color=700)}
    meshes = {
        200: np.stack(_quad((-3, 0, -3), (3, 0, -3), (3, 0, 3), (-3, 0, 3), (0, 1, 0), (-1.5, -1.5), (2.5, 2.5))),     # floor, uvs span [-1.5, 2.5]
        201: np.stack(_quad((-3, 0, -3), (-3, 3, -3), (3, 3, -3), (3, 0, -3), (0, 0, 1))),                              # back wall
        202: np.stack(_quad((-1.5, 0, 0.5), (1.5, 0, 0.5), (1.5, 2.0, 0.5), (-1.5, 2.0, 0.5), (0, 0, 1), (0, 0), (3, 2))),  # fence
        203: np.stack(_quad((-3, 0.5, -2.9), (-3, 2.5, -2.9), (-1, 2.5, -2.9), (-1, 0.5, -2.9), (0, 0, 1))),            # emissive panel
        204: np.stack(_box((0.8, 0.0, -1.6), (1.8, 1.0, -0.6))),                                                         # metal box
    }
    instances = [(300, 200, 100, IDENTITY_AFFINE), (301, 201, 101, IDENTITY_AFFINE), (302, 202, 102, IDENTITY_AFFINE),
                 (303, 203, 103, IDENTITY_AFFINE), (304, 204, 104, IDENTITY_AFFINE)]
    lights = [(400, LIGHT_POINT, point_light((0.0, 2.5, 2.0), 0.1, (6.0, 6.0, 6.0), 20.0)), (401, LIGHT_POINT, point_light((-1.0, 1.5, -1.5), 0.1, (2.0, 2.0, 3.0), 20.0))]
    cam = dict(mode=mode, denoise=denoise, ref_depth=1, w=width, h=height, transform=look_at_transform((0.3, 1.2, 4.0), (0.0, 0.9, 0.0)),
               projection=perspective_infinite_reverse_rh(math.pi / 4.0, width / height, 0.1))
    return dict(name="textured_room", meshes=meshes, materials=materials, instances=instances, lights=lights, sun=(1.0, 0.6), camera=cam,
                images=images, material_textures=material_textures)


def apply(engine, scene):
    """Feed a scene through the Engine API in a fixed order; returns the camera handle."""
    for h, rgba in scene.get("images", {}).items():
        engine.insert_image(h, rgba)
    for h, tris in scene["meshes"].items():
        engine.insert_mesh(h, tris)
    for h, (params, alpha) in scene["materials"].items():
        engine.insert_material(h, params, alpha)
    for h, tex in scene.get("material_textures", {}).items():
        engine.set_material_textures(h, **tex)
    for h, mesh, mat, xf in scene["instances"]:
        engine.insert_instance(h, mesh, mat, xf)
    for h, kind, params in scene["lights"]:
        engine.insert_light(h, kind, params)
    engine.update_sun(*scene["sun"])
    c = scene["camera"]
    return engine.create_camera(c["mode"], c["denoise"], c["ref_depth"], c["w"], c["h"], c["transform"], c["projection"])
