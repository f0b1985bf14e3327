"""The reference's own six unit tests, restated against the oracle.

These are the only golden/known-answer values the reference holds for this path; they pin the
wire layouts (G-buffer packing, DI reservoir, reprojection record, u32 bytes), Camera::contain and
the triangle-slot Allocator.  Everything else in the oracle is "parity unpinned" (see DESIGN.md).
"""
import numpy as np


def test_gbuffer_serialization(oracle):
    # strolle-gpu/src/gbuffer.rs:123-165, EPSILON = 0.005 (alpha 0.1)
    lib = oracle.lib()
    g = np.array([0.1, 0.2, 0.3, 0.4, 0.26, 0.53, 0.80, 0.33, 2.0, 3.0, 4.0, 0.05, 0.25, 123.456], dtype=np.float32)
    packed = np.zeros(8, dtype=np.float32)
    out = np.zeros(14, dtype=np.float32)
    lib.orc_gbuffer_pack(g, packed)
    lib.orc_gbuffer_unpack(packed, out)
    eps = 0.005

    def rel_eq(a, b, e):   # approx::assert_relative_eq!(a, b, epsilon = e): |a-b| <= e or relative
        return abs(a - b) <= e or abs(a - b) <= max(abs(a), abs(b)) * 1.1920929e-7
    for i, want in enumerate([0.1, 0.2, 0.3]):
        assert rel_eq(out[i], want, eps)
    assert rel_eq(out[3], 0.4, 0.1)
    for i, want in zip(range(4, 14), [0.26, 0.53, 0.80, 0.33, 2.0, 3.0, 4.0, 0.05, 0.25, 123.456]):
        assert rel_eq(out[i], want, eps), (i, out[i], want)


def test_camera_contain(oracle):
    # strolle-gpu/src/camera.rs:146-176
    lib = oracle.lib()
    cases = [((0, 0), (0, 0)), ((123, 456), (123, 456)), ((1023, 767), (1023, 767)), ((1024, 768), (1023, 767)),
             ((1025, 768), (1022, 767)), ((1030, 768), (1017, 767)), ((1030, 783), (1017, 752))]
    for (x, y), want in cases:
        out = np.zeros(2, dtype=np.uint32)
        lib.orc_camera_contain(1024.0, 768.0, x, y, out)
        assert tuple(out.tolist()) == want


def test_di_reservoir_serialization(oracle):
    # strolle-gpu/src/reservoir/di.rs:126-163: write 10 slots, read back, exact ==
    lib = oracle.lib()
    buf = np.zeros(2 * 10 * 4, dtype=np.float32)
    wants = []
    for idx in range(10):
        light_id = np.array([3 * idx], dtype=np.uint32).view(np.float32)[0]
        w = np.array([11.0, 12.0 + idx, 123.0, float(idx % 2 == 0), light_id, 1.0, 2.0, 3.0 + idx, float(idx % 2 == 0)], dtype=np.float32)
        wants.append(w)
        out = np.zeros(9, dtype=np.float32)
        lib.orc_di_reservoir_roundtrip(buf, idx, w, out)
    for idx in range(10):   # re-read after all writes (the reference reads in a second loop)
        out = np.zeros(9, dtype=np.float32)
        scratch = buf.copy()
        lib.orc_di_reservoir_roundtrip(scratch, idx, wants[idx], out)
        assert out.view(np.uint32).tolist() == wants[idx].view(np.uint32).tolist()
        assert scratch.view(np.uint32).tolist() == buf.view(np.uint32).tolist()


def test_reprojection_serialization(oracle):
    # strolle-gpu/src/reprojection.rs:77-97
    lib = oracle.lib()
    src = np.array([123.45, 234.56, 1.23], dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    val = np.zeros(1, dtype=np.uint32)
    lib.orc_reprojection_roundtrip(src, 0xCAFEBABE, out, val)
    assert out.tolist() == src.tolist() and int(val[0]) == 0xCAFEBABE


def test_u32_from_to_bytes(oracle):
    # strolle-gpu/src/utils/u32_ext.rs:27-35
    assert oracle.lib().orc_u32_bytes_roundtrip(0xCAFEBABE) == 0xCAFEBABE


def test_allocator(oracle):
    # strolle/src/utils/allocator.rs:60-126 (give = op 0, take = op 1)
    G, T = 0, 1
    script = [(T, 16, 0, None),
              (G, 0, 32, None), (T, 8, 0, (0, 8)), (T, 8, 0, (8, 16)), (T, 8, 0, (16, 24)), (T, 8, 0, (24, 32)), (T, 8, 0, None),
              (G, 0, 8, None), (G, 10, 15, None), (T, 4, 0, (0, 4)), (T, 4, 0, (4, 8)), (T, 4, 0, (10, 14)), (T, 4, 0, None), (T, 1, 0, (14, 15)), (T, 1, 0, None),
              (G, 0, 8, None), (G, 8, 16, None), (G, 16, 24, None), (G, 24, 32, None), (G, 32, 40, None), (G, 64, 256, None),
              (T, 64, 0, (64, 128)), (T, 20, 0, (0, 20)), (T, 20, 0, (20, 40)), (T, 20, 0, (128, 148))]
    def run(script):
        ops = np.array([[op, a, b] for op, a, b, _ in script], dtype=np.int64).reshape(-1)
        out = np.zeros_like(ops)
        oracle.lib().orc_allocator_script(ops, len(script), out)
        out = out.reshape(-1, 3)
        for i, (op, a, b, want) in enumerate(script):
            if op == T:
                got = (int(out[i][1]), int(out[i][2])) if out[i][0] else None
                assert got == want, (i, got, want)
    run(script)
    # case 3b: fresh allocator, reverse order gives
    run([(G, 64, 256, None), (G, 32, 40, None), (G, 24, 32, None), (G, 16, 24, None), (G, 8, 16, None), (G, 0, 8, None),
         (T, 64, 0, (64, 128)), (T, 20, 0, (0, 20)), (T, 20, 0, (20, 40)), (T, 20, 0, (128, 148))])
