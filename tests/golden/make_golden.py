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
  render_mesh.npz     reference Renderer::render of a 6,000-triangle tessellated mesh scene (displaced terrain, tori,
                      spheres: diffuse, MIRROR, rough specular, GLASS, smooth normals), 32x18, 6 bounces: the float64 mean of 64
                      independent single-threaded 16384-spp frames (1M spp; render_mesh_golden says why one thread and why
                      frames of 16384) and the means of its four quarters (the reference's own remaining noise).
                      The scene generator's roughness-0.1 object is set to 0.35: the reference's SPECULAR direct term is
                      albedo * D_GGX with no normalisation (renderer.hpp:286-290), peak 1/(pi r^4) = 3183 at r = 0.1, and
                      two 16384-spp runs of the REFERENCE ITSELF then differ by 57-81 % relRMSE per channel and 2.2 % in mean
                      luminance (3.7-4.2 % and 0.05 % at r = 0.35) — no finite-sample gate can be stated on it.

  render_c1.npz       BASELINE.md §4's image-parity run: the reference renderer on BASELINE configs[0] — the loader's Cornell
                      scene at 800x450, 5 bounces — at 4096 spp (1.47 G samples, the better part of an hour on 8 cores),
                      stored as 10x10-pixel block means (80x45x3 floats): at 4096 spp a single pixel of the reference is
                      still ~5 % noisy, a block mean is not.

    python tests/golden/make_golden.py [trace|cornell|mesh|c1 ...]     (default: trace cornell mesh)
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


def _mesh_scene():
    ms = scenes.mesh_scene(6000, seed=11)
    m8 = ms["materials8"].copy()
    m8[5, 4] = 0.35
    return ms, m8


def _mesh_worker(nruns):
    """nruns sequential single-threaded reference frames (own process, own RefOracle); returns their float64 sum."""
    ms, m8 = _mesh_scene()
    R = RefOracle(ms["pos"], ms["nrm"], ms["mat"], m8)
    acc = np.zeros((MESH_H, MESH_W, 3), np.float64)
    for _ in range(nruns):
        acc += R.render(MESH_W, MESH_H, MESH_SPP, MESH_B, nthreads=1)[0]
    return acc


MESH_W, MESH_H, MESH_SPP, MESH_B, MESH_GROUPS, MESH_RUNS_PER_GROUP = 32, 18, 16384, 6, 4, 16


def render_mesh_golden():
    # 64 INDEPENDENT reference frames of 16384 spp, each by Renderer::render on ONE thread, averaged in float64.
    #  * one thread: render() draws tracePath's random numbers from the Renderer's member mt19937, which its OpenMP
    #    loop shares between threads without synchronisation (renderer.hpp:53, :128-130) — a multi-threaded frame is
    #    the reference with a raced generator (duplicated / skipped / out-of-range state reads), not a clean target;
    #  * 16384 spp per frame: render() adds samples in fp32 (renderer.hpp:73); beyond ~10^5 samples the running sum's ulp
    #    reaches the size of a typical sample and the pixel mean drifts (on this scene's highlight pixel the fp32 mean of
    #    4M samples is 1 % below the float64 mean of the same estimator), so a converged target is an AVERAGE OF FRAMES.
    # Stored: the mean of all frames (fb_ref, 1,048,576 spp) and the four group means (fb_runs, 262144 spp each): their
    # scatter is the reference's own remaining noise.
    import multiprocessing as mp
    import time
    ms, m8 = _mesh_scene()
    R = RefOracle(ms["pos"], ms["nrm"], ms["mat"], m8)
    workers = min(mp.cpu_count(), MESH_GROUPS * MESH_RUNS_PER_GROUP)
    per = MESH_GROUPS * MESH_RUNS_PER_GROUP // workers
    assert per * workers == MESH_GROUPS * MESH_RUNS_PER_GROUP and workers % MESH_GROUPS == 0
    t0 = time.time()
    with mp.get_context("spawn").Pool(workers) as pool:
        sums = pool.map(_mesh_worker, [per] * workers)
    secs = time.time() - t0
    g = workers // MESH_GROUPS
    runs = np.stack([sum(sums[k * g:(k + 1) * g]) / (g * per) for k in range(MESH_GROUPS)])
    np.savez_compressed(os.path.join(HERE, "render_mesh.npz"), pos=ms["pos"], nrm=ms["nrm"], mat=ms["mat"], order=R.order(),
                        materials8=m8, fb_ref=runs.mean(0).astype(np.float32), fb_runs=runs.astype(np.float32),
                        spp=MESH_SPP * MESH_GROUPS * MESH_RUNS_PER_GROUP, frame_spp=MESH_SPP, bounces=MESH_B)
    print(f"render_mesh: {MESH_GROUPS * MESH_RUNS_PER_GROUP} x {MESH_W}x{MESH_H}x{MESH_SPP} in {secs:.1f}s mean {runs.mean():.5f}")


def render_c1_golden():
    tmp = tempfile.mkdtemp()
    obj = scenes.write_cornell_obj(tmp)
    sc = pt.Scene()
    assert sc.loadFromObj(obj)
    R = RefOracle(obj_path=obj)
    pre_pos, pre_nrm, pre_mat = prebuild_from_scene(sc)
    W, H, SPP, B, K = 800, 450, 4096, 5, 10
    fb, secs = R.render(W, H, SPP, B)
    blocks = fb.reshape(H // K, K, W // K, K, 3).mean((1, 3)).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "render_c1.npz"), pos=pre_pos, nrm=pre_nrm, mat=pre_mat, order=sc.order,
                        materials8=sc.materials8, blocks_ref=blocks, block=K, width=W, height=H, spp=SPP, bounces=B)
    print(f"render_c1: {W}x{H}x{SPP} in {secs:.1f}s mean {fb.mean():.5f}")


def main():
    oracle.build("ref")
    what = set(sys.argv[1:]) or {"trace", "cornell", "mesh"}
    if "mesh" in what:
        render_mesh_golden()
    if "c1" in what:
        render_c1_golden()
    if not ({"trace", "cornell"} & what):
        return
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
