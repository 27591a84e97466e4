"""Generates the committed golden vectors from the REFERENCE ITSELF (oracle/_ref = the unmodified
reference headers compiled against oracle/glm_shim).  Needs /root/reference (or a prebuilt
oracle/_ref/libref_oracle.so); run from the repo root:

    python tests/golden/make_golden.py

Outputs (small, committed):
  trace_soup.npz      2,000 random triangles, 20,000 rays: reference BVH order, closest-hit triangle
                      ids and hit-distance bit patterns, with tMax = +inf and with per-ray finite tMax.
  trace_cornell.npz   the loader's Cornell scene (room + model), 20,000 rays incl. rays aimed at vertices.
  render_cornell.npz  reference Renderer::render of the Cornell scene, 32x18, 262144 spp, 5 bounces
                      (float framebuffer) + the scene arrays (pre-build order) so that tests do not
                      depend on OBJ parsing.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import RefOracle  # noqa: E402
import path_tracer_ai_b200 as pt  # noqa: E402
from path_tracer_ai_b200 import scenes  # noqa: E402
from conftest import prebuild_from_scene  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    oracle.build("ref")
    rng = np.random.default_rng(20261018)

    pos = scenes.random_soup(2000, 42)
    R = RefOracle(pos)
    o = (rng.random((20000, 3)) * 2.4 - 1.2).astype(np.float32)
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    tmax = (rng.random(20000) * 2.5).astype(np.float32)
    tri, t = R.trace_closest(o, d)
    tri2, t2 = R.trace_closest(o, d, tmax)
    np.savez_compressed(os.path.join(HERE, "trace_soup.npz"), pos=pos, order=R.order(), o=o, d=d, tmax=tmax,
                        tri=tri, t_bits=bits(t), tri_tmax=tri2, t_tmax_bits=bits(t2))
    print("trace_soup: hits", int((tri >= 0).sum()), "hits with tmax", int((tri2 >= 0).sum()))

    tmp = tempfile.mkdtemp()
    obj = scenes.write_cornell_obj(tmp)
    sc = pt.Scene()
    assert sc.loadFromObj(obj)
    Rc = RefOracle(obj_path=obj)
    rp, rn, rm = Rc.triangles()
    assert np.array_equal(bits(rp), bits(sc.pos)) and np.array_equal(rm, sc.mat)
    pre_pos, pre_nrm, pre_mat = prebuild_from_scene(sc)
    # rays: camera-like + aimed at vertices (tie pressure) + random
    V = sc.pos.reshape(-1, 3)
    n = 20000
    o = np.empty((n, 3), np.float32)
    d = np.empty((n, 3), np.float32)
    o[: n // 2] = np.float32([0.0, 2.0, 5.0]) + (rng.random((n // 2, 3)) - 0.5).astype(np.float32) * np.float32(0.5)
    d[: n // 2] = V[rng.integers(0, len(V), n // 2)] - o[: n // 2]
    o[n // 2:] = (rng.random((n - n // 2, 3)) * [6, 4, 6] - [3, 0, 3]).astype(np.float32)
    d[n // 2:] = rng.normal(size=(n - n // 2, 3)).astype(np.float32)
    tri, t = Rc.trace_closest(o, d)
    np.savez_compressed(os.path.join(HERE, "trace_cornell.npz"), pos=pre_pos, nrm=pre_nrm, mat=pre_mat, order=sc.order,
                        o=o, d=d, tri=tri, t_bits=bits(t))
    print("trace_cornell: hits", int((tri >= 0).sum()), "tree", Rc.tree_stats())

    W, H, SPP, B = 32, 18, 262144, 5
    fb, secs = Rc.render(W, H, SPP, B)
    np.savez_compressed(os.path.join(HERE, "render_cornell.npz"), pos=pre_pos, nrm=pre_nrm, mat=pre_mat, order=sc.order,
                        materials8=sc.materials8, fb_ref=fb, spp=SPP, bounces=B)
    print(f"render_cornell: {W}x{H}x{SPP} in {secs:.1f}s mean {fb.mean():.5f}")


if __name__ == "__main__":
    main()
