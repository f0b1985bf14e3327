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


if __name__ == "__main__":
    cornell()
    blue_noise()
