"""Host-side BVH builder: product (C ABI, no GPU needed) against the oracle, and the
subtree-reuse path against from-scratch builds.

The reference holds no test for its builder, so the pins are: product == oracle bit for bit on the
serialised stream, structural invariants of that stream, reuse == fresh whenever primitives only move, and the
reference's centre-only hash (strolle/src/bvh/primitive.rs:27-37) behaving as restated when they do not.
"""
import numpy as np
import pytest

F32_MAX = np.float32(3.4028234663852886e38)


def make_prims(rng, n_objects=12, tris_per_object=40):
    """Clustered triangle soup: objects = clouds of small triangles; returns (n, 11) float32 primitive records."""
    rows = []
    for o in range(n_objects):
        c = rng.uniform(-10, 10, size=3)
        for _ in range(tris_per_object):
            p = c + rng.normal(scale=0.8, size=(3, 3))
            p = p.astype(np.float32)
            center = ((p[0] + p[1]) + p[2]) / np.float32(3.0)
            rows.append([0, 0, *center, *p.min(axis=0), *p.max(axis=0)])
    a = np.array(rows, dtype=np.float32)
    a[:, 0] = np.arange(len(a), dtype=np.uint32).view(np.float32)
    a[:, 1] = (np.arange(len(a), dtype=np.uint32) // tris_per_object).view(np.float32)
    return a


def move_object(prims, obj, tris_per_object, delta):
    s = slice(obj * tris_per_object, (obj + 1) * tris_per_object)
    d = np.asarray(delta, dtype=np.float32)
    prims[s, 2:5] += d; prims[s, 5:8] += d; prims[s, 8:11] += d


def same_bits(a, b):
    return a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all()


def check_stream(stream, n_alive):
    """serializer.rs:20-110 invariants: every live primitive appears in exactly one leaf entry, internal nodes point forward."""
    bits = stream.view(np.uint32)
    leaf_tris = []
    i = 0
    n = len(stream)
    stack = [0] if n else []
    seen = set()
    while stack:
        ptr = stack.pop()
        assert ptr not in seen
        seen.add(ptr)
        if bits[ptr, 3] == 0:           # internal: left child follows, right child pointer in d1.w
            right = int(bits[ptr + 1, 3])
            assert ptr + 4 < n and ptr + 4 < right < n
            lo_l, hi_l, lo_r, hi_r = stream[ptr, :3], stream[ptr + 1, :3], stream[ptr + 2, :3], stream[ptr + 3, :3]
            assert (lo_l <= hi_l).all() and (lo_r <= hi_r).all()
            stack.append(right); stack.append(ptr + 4)
        else:                            # run of leaf entries, flag bit 0 = another one follows
            while True:
                assert bits[ptr, 3] == 1
                leaf_tris.append(int(bits[ptr, 1]))
                if not (bits[ptr, 0] & 1):
                    break
                ptr += 1
    assert len(leaf_tris) == n_alive and len(set(leaf_tris)) == n_alive


@pytest.fixture(scope="module")
def builders(oracle):
    from strolle_b200.engine import BvhBuilder
    return BvhBuilder, oracle.OracleBvhBuilder


def test_fresh_build_product_equals_oracle(builders):
    Product, Oracle = builders
    rng = np.random.RandomState(1)
    for n_obj, per in [(1, 1), (1, 2), (3, 7), (12, 40), (30, 100)]:
        prims = make_prims(rng, n_obj, per)
        p, o = Product(), Oracle()
        sp, so = p.build(prims, reuse=False), o.build(prims, reuse=False)
        assert same_bits(sp, so), f"{n_obj}x{per}: product and oracle streams differ"
        assert p.depth == o.depth
        check_stream(sp, len(prims))
    assert Product().build(np.zeros((0, 11), np.float32)).shape[0] == Oracle().build(np.zeros((0, 11), np.float32)).shape[0]


def test_reuse_equals_fresh_when_objects_move(builders):
    """Moving / killing / reviving objects changes centres, so every reused subtree is one a fresh build would produce too."""
    Product, Oracle = builders
    rng = np.random.RandomState(2)
    per = 40
    prims = make_prims(rng, 12, per)
    p, o = Product(), Oracle()
    p.build(prims); o.build(prims)
    total_grafted = 0
    for step in range(12):
        if step % 4 == 3:      # kill an object (primitive.rs:18-20), later bring it back somewhere else
            obj = rng.randint(12)
            prims[obj * per:(obj + 1) * per, 2:5] = F32_MAX
        elif step % 4 == 0 and step:
            dead = np.flatnonzero(prims[:, 2] == F32_MAX)
            if len(dead):
                obj = dead[0] // per
                fresh = make_prims(rng, 1, per)
                prims[obj * per:(obj + 1) * per, 2:] = fresh[:, 2:]
        else:
            move_object(prims, rng.randint(12), per, rng.normal(scale=0.05, size=3))
        sp, so = p.build(prims, reuse=True), o.build(prims, reuse=True)
        fresh = Product().build(prims, reuse=False)
        assert same_bits(sp, so), f"step {step}: product and oracle differ with reuse"
        assert p.grafted == o.grafted
        assert same_bits(sp, fresh), f"step {step}: reuse changed the tree"
        check_stream(sp, int((prims[:, 2] != F32_MAX).sum()))
        total_grafted += p.grafted
    assert total_grafted > 0, "nothing was ever reused"


def test_unchanged_scene_is_reused_wholesale(builders):
    Product, Oracle = builders
    prims = make_prims(np.random.RandomState(3), 6, 30)
    for B in (Product, Oracle):
        b = B()
        first = b.build(prims)
        again = b.build(prims)
        assert same_bits(first, again) and b.grafted == 2, "both root children are taken over"


def test_centre_only_hash_keeps_stale_primitives(builders):
    """Quirk C-20: the reuse hash covers the primitive centres only (primitive.rs:27-37), and a reused subtree brings its
    old primitives along (builder.rs:321-337): changing only the material id (or the bounds) of an instance leaves the
    old values in the tree.  ST_OPT_BVH_REUSE = 0 (reuse=False) gives the tree of the new primitives."""
    Product, Oracle = builders
    prims = make_prims(np.random.RandomState(4), 5, 25)
    changed = prims.copy()
    changed[:25, 1] = np.array([77], dtype=np.uint32).view(np.float32)[0]        # new material for object 0
    changed[25:50, 8:11] += np.float32(0.25)                                      # object 1: bounds grow, centres stay
    for B in (Product, Oracle):
        b = B()
        first = b.build(prims)
        stale = b.build(changed, reuse=True)
        assert same_bits(stale, first), "centres unchanged -> the previous tree is reused with its previous primitives"
        fresh = B().build(changed, reuse=False)
        assert not same_bits(fresh, first)
        mats = fresh.view(np.uint32)[fresh.view(np.uint32)[:, 3] == 1][:, 2]
        assert (mats == 77).sum() == 25
    assert same_bits(Product().build(changed, reuse=False), Oracle().build(changed, reuse=False))


def test_builder_rejects_bad_arguments():
    import ctypes as C
    import strolle_b200
    lib = strolle_b200.load_library()
    assert lib.st_bvh_builder_create(None) != 0
    n = C.c_size_t(0)
    assert lib.st_bvh_builder_build(None, None, 0, 1, None, 0, C.byref(n), None, None) != 0


@pytest.mark.parametrize("scene_name", ["cornell", "dungeon"])
def test_product_builder_reproduces_the_oracle_engine_bvh(oracle, blue_noise, scene_name):
    """The benchmark scenes themselves: primitives rebuilt from the oracle engine's baked triangle buffer (centre = (p0 + p1 + p2) / 3
    in f32, strolle/src/triangle.rs:16-22) go through the product's host builder and must give the oracle engine's BVH stream."""
    from strolle_b200 import scenes
    from strolle_b200.engine import BvhBuilder
    scene = scenes.cornell(32, 32) if scene_name == "cornell" else scenes.dungeon(32, 32)
    eo = oracle.OracleEngine(blue_noise=blue_noise)
    scenes.apply(eo, scene)
    eo.tick()
    tris = eo.read_scene("triangles").reshape(-1, 9, 4)
    want = eo.read_scene("bvh").reshape(-1, 4)
    bits = want.view(np.uint32)
    leaf = bits[:, 3] == 1
    mat_of = {int(t): int(m) for t, m in zip(bits[leaf, 1], bits[leaf, 2])}
    p0, p1, p2 = tris[:, 0, :3], tris[:, 3, :3], tris[:, 6, :3]
    prims = np.zeros((len(tris), 11), dtype=np.float32)
    prims[:, 0] = np.arange(len(tris), dtype=np.uint32).view(np.float32)
    prims[:, 1] = np.array([mat_of.get(i, 0) for i in range(len(tris))], dtype=np.uint32).view(np.float32)
    prims[:, 2:5] = ((p0 + p1) + p2) / np.float32(3.0)
    prims[:, 5:8] = np.minimum(np.minimum(p0, p1), p2)
    prims[:, 8:11] = np.maximum(np.maximum(p0, p1), p2)
    alive = np.array([i in mat_of for i in range(len(tris))])
    prims[~alive, 2:5] = F32_MAX
    got = BvhBuilder().build(prims, reuse=False)
    assert same_bits(got, want), f"{scene_name}: host builder stream differs from the oracle engine's BVH"
    assert alive.sum() > 30
