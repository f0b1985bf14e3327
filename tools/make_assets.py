#!/usr/bin/env python
"""Generates the committed scene/noise fixtures from the reference's *asset data*.

Run in the build container only (reads /root/reference, which does not exist on
the GPU box).  Outputs (committed):

  strolle_b200/assets/cornell.json     Cornell box: 8 meshes (object-space triangles with
      normals), 8 materials, the glTF root transform — from
      bevy-strolle/assets/cornell.zip ("Cornell Box - Original" by t-ly, CC-BY-4.0,
      https://sketchfab.com/3d-models/cornell-box-original-0d18de8d108c4c9cab1a4405698cc6b6)
  strolle_b200/assets/blue_noise_256_rgba8.bin   256x256 RGBA8 blue noise, raw bytes of
      strolle/assets/blue-noise.png (Christoph Peters, momentsingraphics.de/BlueNoise.html, CC0)

  strolle_b200/assets/dungeon.npz      the reference's demo level (BASELINE config C3): 45 meshes / 8,393 triangles with
      normals and uvs, the world transform of every mesh node, 45 materials and their 64x64 base-colour textures decoded
      to RGBA8 — from bevy-strolle/assets/demo.zip -> demo/level.glb ("Low Poly Game Level" by MaxDeaconVR, CC-BY-4.0,
      https://sketchfab.com/3d-models/low-poly-game-level-82b7a937ae504cfa9f277d9bf6874ad2)

No reference *source code* is copied; these are data assets with their licences noted.
"""
import io, json, struct, zipfile, sys, os
import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "strolle_b200", "assets")


def cornell():
    z = zipfile.ZipFile(f"{REF}/bevy-strolle/assets/cornell.zip")
    g = json.loads(z.read("cornell/scene.gltf"))
    blob = z.read("cornell/scene.bin")

    def accessor(i):
        a = g["accessors"][i]
        bv = g["bufferViews"][a["bufferView"]]
        off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        n = a["count"]
        if a["componentType"] == 5126:
            comps = {"VEC3": 3, "VEC2": 2, "SCALAR": 1}[a["type"]]
            stride = bv.get("byteStride", 4 * comps)
            return np.array([struct.unpack_from("<%df" % comps, blob, off + i * stride) for i in range(n)], dtype=np.float32)
        elif a["componentType"] == 5125:
            return np.frombuffer(blob, dtype="<u4", count=n, offset=off)
        raise ValueError(a)

    # node hierarchy -> world matrix per mesh node (column-major 4x4 lists as in glTF)
    def node_matrix(n):
        if "matrix" in n:
            return np.array(n["matrix"], dtype=np.float64).reshape(4, 4).T
        m = np.eye(4)
        if "translation" in n:
            m[:3, 3] = n["translation"]
        if "scale" in n:
            m = m @ np.diag(list(n["scale"]) + [1.0])
        assert "rotation" not in n
        return m

    out_meshes = []
    def walk(idx, parent):
        n = g["nodes"][idx]
        m = parent @ node_matrix(n)
        if "mesh" in n:
            mesh = g["meshes"][n["mesh"]]
            for prim in mesh["primitives"]:
                pos = accessor(prim["attributes"]["POSITION"])
                nor = accessor(prim["attributes"]["NORMAL"])
                idxs = accessor(prim["indices"])
                tris = []
                for t in range(0, len(idxs), 3):
                    i0, i1, i2 = int(idxs[t]), int(idxs[t + 1]), int(idxs[t + 2])
                    tris.append({"positions": [pos[i].tolist() for i in (i0, i1, i2)],
                                 "normals": [nor[i].tolist() for i in (i0, i1, i2)]})
                out_meshes.append({"name": mesh["name"], "node": idx, "material": prim["material"],
                                   "transform_colmajor": np.asarray(m, dtype=np.float32).T.reshape(-1).tolist(),
                                   "triangles": tris})
        for c in n.get("children", []):
            walk(c, m)
    for r in g["scenes"][g.get("scene", 0)]["nodes"]:
        walk(r, np.eye(4))
    mats = []
    for m in g["materials"]:
        pbr = m["pbrMetallicRoughness"]
        mats.append({"name": m["name"], "base_color": pbr["baseColorFactor"], "metallic": pbr.get("metallicFactor", 1.0),
                     "perceptual_roughness": pbr.get("roughnessFactor", 1.0)})
    doc = {"source": "bevy-strolle/assets/cornell.zip (Cornell Box - Original, t-ly, CC-BY-4.0)",
           "meshes": sorted(out_meshes, key=lambda m: m["node"]), "materials": mats}
    ntri = sum(len(m["triangles"]) for m in out_meshes)
    print("cornell: meshes", len(out_meshes), "triangles", ntri)
    with open(os.path.join(OUT, "cornell.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))


def blue_noise():
    from PIL import Image
    img = Image.open(f"{REF}/strolle/assets/blue-noise.png")
    print("blue noise", img.size, img.mode)
    arr = np.asarray(img.convert("RGBA"), dtype=np.uint8)
    assert arr.shape == (256, 256, 4)
    arr.tofile(os.path.join(OUT, "blue_noise_256_rgba8.bin"))


def dungeon():
    """demo/level.glb -> dungeon.npz.  Per mesh node: object-space triangle soup (positions, normals, TEXCOORD_0 in index
    order, as bevy-strolle/src/stages/prepare.rs:22-122 feeds them), its material index and its world matrix (product of the
    node chain in f64, rounded to f32; Bevy propagates GlobalTransform the same way up to f32 rounding)."""
    from PIL import Image
    z = zipfile.ZipFile(f"{REF}/bevy-strolle/assets/demo.zip")
    b = z.read("demo/level.glb")
    magic, ver, length = struct.unpack_from("<III", b, 0)
    assert magic == 0x46546C67 and ver == 2
    off = 12
    clen, ctype = struct.unpack_from("<II", b, off); off += 8
    g = json.loads(b[off:off + clen]); off += clen
    blen, btype = struct.unpack_from("<II", b, off); off += 8
    blob = b[off:off + blen]

    def accessor(i):
        a = g["accessors"][i]
        bv = g["bufferViews"][a["bufferView"]]
        o = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        n = a["count"]
        comps = {"VEC4": 4, "VEC3": 3, "VEC2": 2, "SCALAR": 1}[a["type"]]
        dt = {5126: "<f4", 5125: "<u4", 5123: "<u2", 5121: "u1"}[a["componentType"]]
        size = np.dtype(dt).itemsize * comps
        stride = bv.get("byteStride", size)
        if stride == size:
            return np.frombuffer(blob, dtype=dt, count=n * comps, offset=o).reshape(n, comps).copy()
        return np.stack([np.frombuffer(blob, dtype=dt, count=comps, offset=o + k * stride) for k in range(n)])

    def quat_matrix(q):
        x, y, zz, w = q
        return np.array([[1 - 2 * (y * y + zz * zz), 2 * (x * y - zz * w), 2 * (x * zz + y * w)],
                         [2 * (x * y + zz * w), 1 - 2 * (x * x + zz * zz), 2 * (y * zz - x * w)],
                         [2 * (x * zz - y * w), 2 * (y * zz + x * w), 1 - 2 * (x * x + y * y)]])

    def node_matrix(n):
        if "matrix" in n:
            return np.array(n["matrix"], dtype=np.float64).reshape(4, 4).T
        m = np.eye(4)
        r = quat_matrix(n["rotation"]) if "rotation" in n else np.eye(3)
        sc = np.array(n.get("scale", [1.0, 1.0, 1.0]))
        m[:3, :3] = r @ np.diag(sc)
        m[:3, 3] = n.get("translation", [0.0, 0.0, 0.0])
        return m

    out = {}
    mesh_nodes = []
    def walk(idx, parent):
        n = g["nodes"][idx]
        m = parent @ node_matrix(n)
        if "mesh" in n:
            mesh_nodes.append((idx, n["mesh"], m))
        for c in n.get("children", []):
            walk(c, m)
    for r in g["scenes"][g.get("scene", 0)]["nodes"]:
        walk(r, np.eye(4))
    ntri = 0
    mats, xforms = [], []
    for k, (idx, mi, m) in enumerate(mesh_nodes):
        prims = g["meshes"][mi]["primitives"]
        assert len(prims) == 1 and prims[0].get("mode", 4) == 4
        pr = prims[0]
        pos, nor, uv = accessor(pr["attributes"]["POSITION"]), accessor(pr["attributes"]["NORMAL"]), accessor(pr["attributes"]["TEXCOORD_0"])
        ix = accessor(pr["indices"]).reshape(-1).astype(np.int64)
        out[f"pos{k}"] = pos[ix].astype(np.float32).reshape(-1, 3, 3)
        out[f"nor{k}"] = nor[ix].astype(np.float32).reshape(-1, 3, 3)
        out[f"uv{k}"] = uv[ix].astype(np.float32).reshape(-1, 3, 2)
        ntri += len(ix) // 3
        mats.append(pr["material"])
        xforms.append(np.asarray(m, dtype=np.float32).T.reshape(-1))   # column-major 4x4
    out["mesh_material"] = np.array(mats, dtype=np.int32)
    out["mesh_transform_colmajor"] = np.stack(xforms)
    base, tex = [], []
    for m in g["materials"]:
        pbr = m.get("pbrMetallicRoughness", {})
        assert m.get("alphaMode", "OPAQUE") == "OPAQUE" and "emissiveFactor" not in m
        base.append(pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0]) + [pbr.get("metallicFactor", 1.0), pbr.get("roughnessFactor", 1.0)])
        tex.append(g["textures"][pbr["baseColorTexture"]["index"]]["source"] if "baseColorTexture" in pbr else -1)
    out["material_base_metallic_roughness"] = np.array(base, dtype=np.float32)
    out["material_texture"] = np.array(tex, dtype=np.int32)
    imgs = []
    for im in g["images"]:
        bv = g["bufferViews"][im["bufferView"]]
        o = bv.get("byteOffset", 0)
        imgs.append(np.asarray(Image.open(io.BytesIO(blob[o:o + bv["byteLength"]])).convert("RGBA"), dtype=np.uint8))
    assert all(i.shape == imgs[0].shape for i in imgs)
    out["images_rgba8"] = np.stack(imgs)
    out["source"] = np.array("bevy-strolle/assets/demo.zip demo/level.glb: Low Poly Game Level, MaxDeaconVR, CC-BY-4.0, "
                             "https://sketchfab.com/3d-models/low-poly-game-level-82b7a937ae504cfa9f277d9bf6874ad2")
    print("dungeon: meshes", len(mesh_nodes), "triangles", ntri, "materials", len(base), "images", out["images_rgba8"].shape)
    np.savez_compressed(os.path.join(OUT, "dungeon.npz"), **out)


if __name__ == "__main__":
    cornell()
    blue_noise()
    dungeon()
