"""GPU parity tests for the query path (replaces Scene::intersect): every call goes through the C ABI
(b2pt_trace_closest / b2pt_trace_any) and is compared BIT-EXACTLY with the CPU oracle — triangle ids,
hit/miss, the fp32 bits of t and of the barycentrics."""
import os

import numpy as np
import pytest

import oracle
import path_tracer_ai_b200 as pt
from oracle import PortOracle, RefOracle
from path_tracer_ai_b200 import scenes

from conftest import bits

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rays(n, seed, extent=1.2, centre=(0, 0, 0)):
    rng = np.random.default_rng(seed)
    o = (rng.random((n, 3)) * 2 * extent - extent + np.asarray(centre)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    return o, d


def check_closest(eng, P, o, d, tmax=None, ref=None):
    """GPU (through the C ABI) vs the CPU restatement, and — when `ref` is given — vs THE REFERENCE ITSELF
    (oracle/_ref: the unmodified headers; its BVH is built from the same pre-build triangle list): ids and t bits."""
    tri, t, uv = eng.trace_closest(o, d, tmax)
    rt, rtt, ruv = P.trace_closest(o, d, tmax)
    assert np.array_equal(tri, rt), f"{int((tri != rt).sum())} of {len(tri)} ids differ"
    assert np.array_equal(bits(t), bits(rtt))
    assert np.array_equal(bits(uv), bits(ruv))
    if ref is not None:
        ft, ftt = ref.trace_closest(o, d, tmax)
        assert np.array_equal(tri, ft), f"{int((tri != ft).sum())} of {len(tri)} ids differ from the reference's own intersector"
        assert np.array_equal(bits(t), bits(ftt))
    return tri


def reference_of(pos_prebuild, P):
    """The reference's own BVH over the same triangles (None if oracle/_ref was not built); its order must be ours."""
    if not oracle.ref_available() or len(pos_prebuild) == 0:
        return None
    R = RefOracle(pos_prebuild)
    assert np.array_equal(R.order(), P.order())
    return R


def upload(eng, P):
    pos, nrm, mat = P.triangles()
    eng.upload_scene(pos, nrm, mat)


@pytest.mark.parametrize("ntri", [0, 1, 7, 8, 9, 16, 17, 65, 1000])
def test_small_and_ragged_triangle_counts(engine, ntri):
    pos = scenes.random_soup(ntri, 10 + ntri, size=0.8)
    P = PortOracle(pos)
    assert np.array_equal(pt.reference_order(pos), P.order())
    upload(engine, P)
    o, d = rays(30011, ntri)   # ragged: not a multiple of the block size
    check_closest(engine, P, o, d, ref=reference_of(pos, P))
    occ = engine.trace_any(o, d)
    assert np.array_equal(occ, P.trace_any(o, d))


def test_zero_rays(engine):
    P = PortOracle(scenes.random_soup(100, 1))
    upload(engine, P)
    tri, t, uv = engine.trace_closest(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert tri.shape == (0,) and t.shape == (0,)


def test_soup_closest_and_any(engine):
    pos = scenes.random_soup(50000, 3)
    P = PortOracle(pos)
    upload(engine, P)
    o, d = rays(500000, 4)
    tri = check_closest(engine, P, o, d)
    assert (tri >= 0).mean() > 0.3
    st = engine.stats()
    assert st["extend_rays"] == len(o) and st["kernel_launches"] >= 1
    tmax = np.random.default_rng(9).random(len(o)).astype(np.float32) * 1.5
    check_closest(engine, P, o, d, tmax)
    assert np.array_equal(engine.trace_any(o, d, tmax), P.trace_any(o, d, tmax))


def test_mesh_scene_with_vertex_aimed_rays(engine):
    """Tessellated mesh with shared vertices; half of the rays are aimed exactly at mesh vertices so that
    edge/vertex hits, bit-equal ties and leaf-box-face hits (the cases the fast kernel must hand to the exact
    kernel) actually occur."""
    ms = scenes.mesh_scene(120000, seed=2)
    P = PortOracle(ms["pos"], ms["nrm"], ms["mat"], ms["materials8"])
    pos, nrm, mat = P.triangles()
    engine.upload_scene(pos, nrm, mat, ms["materials8"])
    o, d = scenes.random_rays(400000, ms["lo"], ms["hi"], 5)
    V = ms["pos"].reshape(-1, 3)
    rng = np.random.default_rng(6)
    half = len(o) // 2
    d[:half] = V[rng.integers(0, len(V), half)] - o[:half]
    check_closest(engine, P, o, d, ref=reference_of(ms["pos"], P))


def test_axis_aligned_geometry_reference_quirks(engine):
    """An axis-aligned plane (all its reference leaves are flat => invisible to the reference, aabb.hpp:21) plus
    an axis-aligned box sitting in a soup: the engine must reproduce the reference's answer, quirks included."""
    g = 24
    xs, zs = np.meshgrid(np.arange(g, dtype=np.float32), np.arange(g, dtype=np.float32), indexing="ij")
    a = np.stack([xs, np.zeros_like(xs), zs], -1).reshape(-1, 3)
    plane = np.concatenate([np.concatenate([a, a + [1, 0, 0], a + [1, 0, 1]], 1),
                            np.concatenate([a, a + [1, 0, 1], a + [0, 0, 1]], 1)]).astype(np.float32) * np.float32(0.1) - np.float32(1.2)
    V, f = scenes._box(0.0, 0.0, 0.4, 0.8, 0.4, -0.4, 0.0)
    box = np.array([np.concatenate([V[i], V[j], V[k]]) for (i, j, k) in f], np.float32)
    pos = np.concatenate([plane, box, scenes.random_soup(3000, 5)])
    P = PortOracle(pos)
    upload(engine, P)
    o, d = rays(400000, 12, extent=1.5)
    R = reference_of(pos, P)
    check_closest(engine, P, o, d, ref=R)
    # rays straight down / along axes (zero direction components -> inf / NaN slabs)
    o2 = o.copy()
    d2 = np.zeros_like(d)
    d2[:, 1] = -1.0
    d2[::3] = [1.0, 0.0, 0.0]
    d2[1::3] = [0.0, 0.0, -1.0]
    check_closest(engine, P, o2, d2, ref=R)
    assert np.array_equal(engine.trace_any(o2, d2), P.trace_any(o2, d2))
    assert engine.stats()["shadow_rays"] == len(o2)


def big_triangles(n, seed, extent, size):
    rng = np.random.default_rng(seed)
    c = (rng.random((n, 1, 3)) * 2 - 1) * extent
    return (c + (rng.random((n, 3, 3)) - 0.5) * size).reshape(n, 9).astype(np.float32)


@pytest.mark.parametrize("flags", [0, pt.FLAG_LANE_KERNELS, pt.FLAG_POOL_EXTEND])
def test_hoisted_leaves_and_scenes_that_are_all_hoisted(built, flags):
    """Leaves whose box covers most of the scene are kept out of the traversal tree and tested up front (ctx.cuh):
    a soup inside a few room-sized triangles (the loader's situation), and a scene of nothing but huge triangles
    (every leaf hoisted, no tree at all)."""
    eng = pt.Engine(flags=flags)
    for pos in (np.concatenate([big_triangles(8, 1, 0.3, 6.0), scenes.random_soup(6000, 2)]),
                big_triangles(10, 3, 0.2, 5.0),
                big_triangles(40, 4, 0.5, 4.0)):
        P = PortOracle(pos)
        upload(eng, P)
        info = eng.accel_info()
        assert info["hoisted_leaves"] >= 1, info
        o, d = rays(200000, 17, extent=2.0)
        check_closest(eng, P, o, d, ref=reference_of(pos, P))
        tm = np.random.default_rng(5).random(len(o)).astype(np.float32) * 3
        assert np.array_equal(eng.trace_any(o, d, tm), P.trace_any(o, d, tm))
    eng.close()


def test_large_coordinates_large_triangles_near_surface_grazing_rays(engine):
    """The distance cull of the ordered traversal skips subtrees entered later than best*(1+2^-10) + 2^-12*(R+|o|)
    (traverse.cuh): Möller–Trumbore's t has an ABSOLUTE error that grows with the triangle's size and distance, so a
    purely relative slack is not safe for big triangles far from the origin hit from close by at grazing angles."""
    rng = np.random.default_rng(77)
    centre = np.float32([900.0, -400.0, 650.0])
    pos = big_triangles(4000, 5, 40.0, 25.0) + np.tile(centre, 3)
    P = PortOracle(pos)
    upload(engine, P)
    V = P.triangles()[0].reshape(-1, 3, 3)
    n = 300000
    k = rng.integers(0, len(V), n)
    w = rng.random((n, 3)).astype(np.float32); w /= w.sum(1, keepdims=True)
    on = (V[k] * w[:, :, None]).sum(1)                                   # a point on triangle k
    nrm = np.cross(V[k, 1] - V[k, 0], V[k, 2] - V[k, 0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True) + 1e-30
    tang = V[k, 1] - V[k, 0]; tang /= np.linalg.norm(tang, axis=1, keepdims=True) + 1e-30
    lift = (10.0 ** rng.uniform(-3.0, -0.5, (n, 1))).astype(np.float32)   # 0.001 .. 0.3 above the surface
    o = (on + nrm * lift).astype(np.float32)
    graze = (10.0 ** rng.uniform(-4.0, 0.0, (n, 1))).astype(np.float32)
    d = (tang - nrm * graze).astype(np.float32)                          # towards the surface at 0.006 .. 45 degrees
    check_closest(engine, P, o, d, ref=reference_of(pos, P))
    o2, d2 = rays(200000, 3, extent=60.0, centre=centre)
    check_closest(engine, P, o2, d2)


@pytest.mark.parametrize("flags", [0, pt.FLAG_LANE_KERNELS, pt.FLAG_POOL_EXTEND, pt.FLAG_EXACT_ONLY])
def test_nonfinite_rays_are_misses(built, flags):
    """NaN / infinite rays (zero or denormal directions normalise to NaN / inf, ray.hpp:12) are misses in the
    reference after a full-tree walk; the kernels answer at once and must not follow EMPTY child slots.  A scene
    of 2500 triangles has wide nodes with empty slots (10-triangle subtrees)."""
    eng = pt.Engine(flags=flags)
    pos = scenes.random_soup(2500, 21)
    P = PortOracle(pos)
    upload(eng, P)
    o, d = rays(64000, 9)
    nan, inf = np.float32(np.nan), np.float32(np.inf)
    d[0::8] = 0.0                       # normalize(0) = NaN
    d[1::8] = [1e-30, 2e-30, -1e-30]    # |d|^2 underflows: direction becomes +-inf
    d[2::8, 1] = nan
    o[3::8, 0] = nan
    o[4::8, 2] = inf
    o[5::8] = [-inf, inf, nan]
    d[6::8] = [1e-30, 0.0, 0.0]         # (inf, NaN, NaN)
    tri, t, uv = eng.trace_closest(o, d)
    rt, rtt, _ = P.trace_closest(o, d)
    assert np.array_equal(tri, rt) and np.array_equal(bits(t), bits(rtt))
    bad = np.ones(len(o), bool); bad[7::8] = False
    assert (tri[bad] == -1).all() and (tri[~bad] >= 0).any()
    tm = np.full(len(o), 0.9, np.float32)
    occ = eng.trace_any(o, d, tm)
    assert np.array_equal(occ, P.trace_any(o, d, tm)) and not occ[bad].any()
    eng.close()


def test_exact_only_kernel_agrees(built):
    """B2PT_FLAG_EXACT_ONLY routes everything through the flattened reference recursion."""
    eng = pt.Engine(flags=pt.FLAG_EXACT_ONLY)
    pos = scenes.random_soup(20000, 8)
    P = PortOracle(pos)
    upload(eng, P)
    o, d = rays(200000, 3)
    check_closest(eng, P, o, d)
    eng.close()


def test_fetch_counters(built):
    eng = pt.Engine(flags=pt.FLAG_COUNT_FETCHES)
    pos = scenes.random_soup(20000, 8)
    P = PortOracle(pos)
    upload(eng, P)
    o, d = rays(100000, 3)
    check_closest(eng, P, o, d)
    st = eng.stats()
    assert st["node_fetches"] >= len(o) and st["tri_fetches"] > 0
    eng.close()


def test_golden_vectors_from_the_reference(engine):
    """tests/golden/*.npz were produced by the reference itself (oracle/_ref)."""
    g = np.load(os.path.join(GOLD, "trace_soup.npz"))
    order = pt.reference_order(g["pos"])
    assert np.array_equal(order, g["order"])
    engine.upload_scene(g["pos"][order])
    tri, t, _ = engine.trace_closest(g["o"], g["d"])
    assert np.array_equal(tri, g["tri"]) and np.array_equal(bits(t), g["t_bits"])
    tri, t, _ = engine.trace_closest(g["o"], g["d"], g["tmax"])
    assert np.array_equal(tri, g["tri_tmax"]) and np.array_equal(bits(t), g["t_tmax_bits"])
    g = np.load(os.path.join(GOLD, "trace_cornell.npz"))
    order = pt.reference_order(g["pos"])
    assert np.array_equal(order, g["order"])
    engine.upload_scene(g["pos"][order], g["nrm"][order], g["mat"][order])
    tri, t, _ = engine.trace_closest(g["o"], g["d"])
    assert np.array_equal(tri, g["tri"]) and np.array_equal(bits(t), g["t_bits"])


def test_device_pointer_entry_points(engine):
    import torch
    pos = scenes.random_soup(5000, 8)
    P = PortOracle(pos)
    upload(engine, P)
    o, d = rays(100000, 3)
    dev = torch.device("cuda:0")
    to, td = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    tri = torch.empty(len(o), dtype=torch.int32, device=dev)
    t = torch.empty(len(o), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    engine.trace_closest_device(to.data_ptr(), td.data_ptr(), None, len(o), tri.data_ptr(), t.data_ptr(), None)
    rt, rtt, _ = P.trace_closest(o, d)
    assert np.array_equal(tri.cpu().numpy(), rt) and np.array_equal(bits(t.cpu().numpy()), bits(rtt))


def test_call_order_errors(built):
    eng = pt.Engine()
    with pytest.raises(pt.B2ptError, match="no scene uploaded"):
        eng.trace_closest(np.zeros((1, 3), np.float32), np.ones((1, 3), np.float32))
    with pytest.raises(pt.B2ptError, match="no scene uploaded"):
        eng.render(pt.Camera().c, 16, 9, 1, 1)
    eng.close()


def test_full_size_properties_1m_triangles(engine):
    """BASELINE config 4 scale (1M triangles, 8M rays here): oracle-checked on a stratified 200k subset, plus
    size-independent properties on the whole batch — determinism (64-bit hash of ids+t), t consistent with the
    winning triangle's plane, any-hit(tmax = t*(1+eps)) true exactly where closest hits, false with tmax < t."""
    ms = scenes.mesh_scene(1_000_000, seed=1234)
    order = pt.reference_order(ms["pos"])
    pos = ms["pos"][order]
    engine.upload_scene(pos, ms["nrm"][order], ms["mat"][order], ms["materials8"])
    n = 8_000_000
    o, d = scenes.random_rays(n, ms["lo"], ms["hi"], 99)
    tri, t, uv = engine.trace_closest(o, d)
    st = engine.stats()
    assert st["fallback_rays"] < n // 1000
    tri2, t2, _ = engine.trace_closest(o, d)
    assert np.array_equal(tri, tri2) and np.array_equal(bits(t), bits(t2))
    hit = tri >= 0
    assert 0.2 < hit.mean() < 0.98
    # geometric consistency: o + d_n * t lies on the winning triangle (barycentric reconstruction)
    dn = d / np.linalg.norm(d.astype(np.float64), axis=1, keepdims=True)
    Ph = o[hit] + dn[hit] * t[hit, None]
    T = pos[tri[hit]].reshape(-1, 3, 3).astype(np.float64)
    u, v = uv[hit, 0:1].astype(np.float64), uv[hit, 1:2].astype(np.float64)
    Pb = T[:, 0] * (1 - u - v) + T[:, 1] * u + T[:, 2] * v
    assert np.abs(Ph - Pb).max() < 2e-3
    # occlusion consistency
    tm = np.where(hit, t * np.float32(1.001), np.float32(1.0)).astype(np.float32)
    occ = engine.trace_any(o, d, tm)
    assert np.array_equal(occ[hit], np.ones(int(hit.sum()), np.uint8))
    tm2 = np.where(hit, t * np.float32(0.999), np.float32(0.002)).astype(np.float32)
    occ2 = engine.trace_any(o[hit], d[hit], tm2[hit])
    assert occ2.sum() == 0
    # oracle on a stratified subset
    P = PortOracle(ms["pos"], None, None)
    assert np.array_equal(P.order(), order)
    sel = np.arange(0, n, n // 200000)[:200000]
    rt, rtt, _ = P.trace_closest(o[sel], d[sel])
    assert np.array_equal(tri[sel], rt) and np.array_equal(bits(t[sel]), bits(rtt))
