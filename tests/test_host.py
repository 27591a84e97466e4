"""CPU-only tests of the host side in front of the GPU engine: reference BVH ordering, OBJ/MTL loader and
scene normalisation (vs the reference-harness loader), camera, PNG writer, CLI argument contract."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import oracle
import path_tracer_ai_b200 as pt
from oracle import PortOracle, RefOracle
from path_tracer_ai_b200 import scenes

from conftest import bits, reference_tonemap

needs_ref = pytest.mark.skipif(not (oracle.ref_available() or os.path.isdir("/root/reference/include")),
                               reason="oracle/_ref not built and /root/reference absent")


@pytest.mark.parametrize("n,seed", [(0, 1), (1, 1), (8, 2), (9, 3), (1000, 4), (65537, 5)])
def test_reference_order_matches_oracle(built, n, seed):
    pos = scenes.random_soup(n, seed)
    assert np.array_equal(pt.reference_order(pos), PortOracle(pos).order())


@needs_ref
def test_reference_order_matches_reference(built):
    ms = scenes.mesh_scene(20000, seed=9)
    assert np.array_equal(pt.reference_order(ms["pos"]), RefOracle(ms["pos"]).order())


@needs_ref
def test_loader_matches_reference_loader(built, tmp_path):
    obj = scenes.write_cornell_obj(str(tmp_path))
    sc = pt.Scene()
    assert sc.loadFromObj(obj)
    R = RefOracle(obj_path=obj)
    rp, rn, rm = R.triangles()
    assert len(sc.pos) == 50 and R.tree_stats()["flat"] == 0
    assert np.array_equal(bits(rp), bits(sc.pos)) and np.array_equal(bits(rn), bits(sc.nrm)) and np.array_equal(rm, sc.mat)
    assert np.array_equal(bits(R.materials()), bits(sc.materials8))
    assert sc.lights == [tuple(map(tuple, l[:2])) + (l[2],) for l in pt.REFERENCE_LIGHTS] or len(sc.lights) == 4


def test_loader_obj_features(built, tmp_path):
    """Quads / polygons (fan triangulation), negative indices, v/vt/vn forms, missing MTL, reference name rules."""
    (tmp_path / "m.mtl").write_text(
        "newmtl shiny_red_paint\nKd 0.1 0.2 0.3\n\nnewmtl gold_trim\nKd 0.5 0.5 0.5\n\nnewmtl plain\nKd 0.25 0.5 1.0\n\n"
        "newmtl glass_pane\nKd 1 1 1\nNi 1.33\n\nnewmtl rough0.25_x\nKd 0.4 0.4 0.4\n\nnewmtl darksilver_bolt\nKd 0 0 0\n")
    (tmp_path / "s.obj").write_text(
        "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 2 0.5\nvn 0 0 1\nvt 0.5 0.5\n"
        "usemtl shiny_red_paint\nf 1 2 3 4\n"            # quad -> 2 triangles
        "usemtl gold_trim\nf -5//1 -4//1 -3//1\n"        # negative indices, v//vn
        "usemtl plain\nf 1/1/1 2/1/1 5/1/1\n"
        "usemtl glass_pane\nf 1 3 5\nusemtl rough0.25_x\nf 2 3 5\nusemtl darksilver_bolt\nf 1 2 3 4 5\n"   # pentagon -> 3
        "usemtl nosuch\nf 1 2 3\n")
    sc = pt.Scene()
    assert sc.loadFromObj(str(tmp_path / "s.obj"))
    assert len(sc.pos) == 8 + 2 + 1 + 1 + 1 + 1 + 3 + 1
    m = sc.materials8
    assert m.shape[0] == 2 + 6
    assert m[0, 0] == 1 and np.allclose(m[0, 1:4], [0.9, 0.2, 0.2])          # default red specular (scene.cpp:58-63)
    assert m[1, 0] == 0 and np.allclose(m[1, 1:4], 0.9)                        # wall (scene.cpp:66-71)
    assert m[2, 0] == 1 and np.allclose(m[2, 1:4], [0.9, 0.2, 0.2]) and np.isclose(m[2, 4], 0.1)    # contains "red"
    assert np.allclose(m[3, 1:4], [1.0, 0.8, 0.0]) and np.isclose(m[3, 4], 0.05)                   # "gold"
    want = np.clip(np.power(np.float32([0.25, 0.5, 1.0]), np.float32(0.8)) * np.float32(1.2), 0, 1)
    assert m[4, 0] == 1 and np.allclose(m[4, 1:4], want, atol=1e-6)            # pow(Kd, .8) * 1.2 clamped
    assert m[5, 0] == 2 and np.isclose(m[5, 6], 1.33)                          # extension: glass*
    assert m[6, 0] == 1 and np.isclose(m[6, 4], 0.25)                          # extension: rough<value>*
    assert np.allclose(m[7, 1:4], 0.95)                                        # "silver" (darksilver)
    pre_mat = np.empty_like(sc.mat)
    pre_mat[sc.order] = sc.mat
    assert list(pre_mat[:8]) == [1] * 8                                        # the room comes first
    assert list(pre_mat[8:]) == [2, 2, 3, 4, 5, 6, 7, 7, 7, 2]                 # unknown usemtl -> -1 -> 0 -> +2
    assert not pt.Scene().loadFromObj(str(tmp_path / "missing.obj"))
    # normalisation: model scaled to 3 units, centred, z flipped, lifted 1.8 (scene.cpp:47-52, :236-238)
    pre_pos = np.empty_like(sc.pos)
    pre_pos[sc.order] = sc.pos
    model = pre_pos[8:].reshape(-1, 3)
    assert np.isclose(model[:, 1].max() - model[:, 1].min(), 3.0, atol=1e-5)
    assert np.isclose((model[:, 1].max() + model[:, 1].min()) / 2, 1.8, atol=1e-5)


def test_chunked_obj_parser_equals_line_by_line(built, tmp_path):
    """The multi-threaded chunked parser must return exactly what a sequential read returns: relative (negative)
    indices, `usemtl` state and `mtllib` lookups all depend on earlier lines.  Chunk sizes down to a few bytes put
    every construct on a chunk boundary."""
    from path_tracer_ai_b200.renderer import _host_lib
    L = _host_lib()
    rng = np.random.default_rng(3)
    (tmp_path / "a.mtl").write_text("newmtl diffuse_red\nKd 0.8 0.1 0.1\nnewmtl glass_x\nKd 1 1 1\nNi 1.45\n")
    (tmp_path / "b.mtl").write_text("# second library\nnewmtl mirror_late\nKd 0.9 0.9 0.9\nnewmtl diffuse_red\nKd 0.1 0.8 0.1\n")
    lines = ["# generated", "usemtl diffuse_red   # used before any mtllib: not found", "mtllib a.mtl"]
    nv = nn = nt = 0
    for blk in range(60):
        k = int(rng.integers(3, 9))
        for _ in range(k):
            x, y, z = rng.normal(size=3)
            lines.append(f"v {x:.6f} {y:.6f} {z:.6f}" + (" 1.0" if rng.random() < 0.2 else ""))
            nv += 1
            if rng.random() < 0.7:
                lines.append("vn %.4f %.4f %.4f" % tuple(rng.normal(size=3))); nn += 1
            if rng.random() < 0.5:
                lines.append("vt %.3f %.3f" % tuple(rng.random(2))); nt += 1
        if blk == 20:
            lines.append("mtllib b.mtl")
        if rng.random() < 0.5:
            lines.append("usemtl " + str(rng.choice(["diffuse_red", "glass_x", "mirror_late", "nope"])))
        lines.append("g group%d" % blk)
        for _ in range(int(rng.integers(1, 5))):
            m = int(rng.integers(3, 6))      # triangles, quads, pentagons
            form = int(rng.integers(0, 5))
            toks = []
            for _ in range(m):
                vi = int(rng.integers(1, nv + 1)) if rng.random() < 0.5 else -int(rng.integers(1, min(nv, 6) + 1))
                ti = (int(rng.integers(1, nt + 1)) if rng.random() < 0.5 else -1) if nt else 0
                ni = (int(rng.integers(1, nn + 1)) if rng.random() < 0.5 else -1) if nn else 0
                if form == 0 or (form in (1, 3) and not ti) or (form in (2, 3) and not ni): toks.append(f"{vi}")
                elif form == 1: toks.append(f"{vi}/{ti}")
                elif form == 2: toks.append(f"{vi}//{ni}")
                else: toks.append(f"{vi}/{ti}/{ni}")
            lines.append("f " + " ".join(toks))
        if rng.random() < 0.1:
            lines.append("f 1 2")            # malformed: skipped with a warning
        if rng.random() < 0.1:
            lines.append("   \t  ")
    text = "\r\n".join(lines[:40]) + "\r\n" + "\n".join(lines[40:])   # CRLF part, LF part, no trailing newline
    obj = tmp_path / "mix.obj"
    obj.write_bytes(text.encode())
    for nthreads, chunk in [(1, 1 << 20), (4, 7), (3, 64), (8, 1), (2, 333), (0, 1000)]:
        assert L.b2pt_obj_parser_selfcheck(os.fsencode(str(obj)), nthreads, chunk) == 0, (nthreads, chunk)
    assert L.b2pt_obj_parser_selfcheck(os.fsencode(str(tmp_path / "missing.obj")), 2, 64) < 0
    (tmp_path / "empty.obj").write_bytes(b"")
    assert L.b2pt_obj_parser_selfcheck(os.fsencode(str(tmp_path / "empty.obj")), 4, 1) == 0


def test_binary_scene_cache_roundtrip(built, tmp_path):
    """The scene cache returns exactly what the loader builds (triangles in post-build order, order, materials),
    is rejected once the OBJ changes, and a corrupt cache falls back to parsing."""
    obj = scenes.write_cornell_obj(str(tmp_path))
    a = pt.Scene(); assert a.loadFromObj(obj)
    b = pt.Scene(); assert b.loadFromObj(obj, cache=True)              # writes <obj>.b2ptscene
    assert os.path.exists(obj + ".b2ptscene")
    c = pt.Scene(); assert c.loadFromObj(obj, cache=True)              # reads it
    for x in (b, c):
        assert np.array_equal(bits(x.pos), bits(a.pos)) and np.array_equal(bits(x.nrm), bits(a.nrm))
        assert np.array_equal(x.mat, a.mat) and np.array_equal(x.order, a.order) and np.array_equal(x.materials8, a.materials8)
        assert x.lights == a.lights
    # a changed OBJ invalidates the cache
    text = open(obj).read()
    open(obj, "w").write(text + "v 9 9 9\nv 9.1 9 9\nv 9 9.1 9\nusemtl diffuse_white\nf -3 -2 -1\n")
    d = pt.Scene(); assert d.loadFromObj(obj, cache=True)
    assert len(d.pos) == len(a.pos) + 1
    # garbage in the cache file: ignored
    open(obj + ".b2ptscene", "wb").write(b"not a cache")
    e = pt.Scene(); assert e.loadFromObj(obj, cache=True)
    assert np.array_equal(bits(e.pos), bits(d.pos))


def test_obj_out_of_range_normal_and_texcoord_indices(built, tmp_path):
    """An OBJ whose faces name normals / texture coordinates that do not exist must not crash the loader: those
    references are treated as absent (the loader then uses the face normal), a missing VERTEX fails the load."""
    p = tmp_path / "bad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\n"
                 "f 1//300000000 2//300000000 3//300000000\n"     # normals far out of range
                 "f 1/77/1 2/77/1 3/77/1\n"                        # texcoords out of range, normals fine
                 "f 1//-5 2//-5 3//-5\n")                          # negative index reaching before the first normal
    sc = pt.Scene()
    assert sc.loadFromObj(str(p))
    assert len(sc.pos) == 8 + 3
    assert np.isfinite(sc.nrm).all()
    (tmp_path / "bad2.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    assert not pt.Scene().loadFromObj(str(tmp_path / "bad2.obj"))
    # both parsers agree on the malformed file
    L = pt.load_library()
    from path_tracer_ai_b200 import renderer
    renderer._host_lib()
    assert L.b2pt_obj_parser_selfcheck(str(p).encode(), 4, 7) == 0


def test_python_binding_rejects_mismatched_array_lengths(built):
    """The C ABI takes array lengths on trust; the binding must raise instead of letting it read past a short buffer."""
    from path_tracer_ai_b200 import _capi
    pos = np.zeros((4, 9), np.float32)
    with pytest.raises(ValueError):
        _capi.check_scene_arrays(pos, np.zeros((3, 9), np.float32), None, None, [])
    with pytest.raises(ValueError):
        _capi.check_scene_arrays(pos, None, np.zeros(5, np.int32), None, [])
    with pytest.raises(ValueError):
        _capi.check_scene_arrays(np.zeros(10, np.float32), None, None, None, [])
    with pytest.raises(ValueError):
        _capi.check_scene_arrays(pos, None, None, None, [((0, 0, 0), (1, 1, 1), 1.0)] * 17)
    with pytest.raises(ValueError):
        _capi.check_ray_arrays(np.zeros((5, 3)), np.zeros((4, 3)), None)
    with pytest.raises(ValueError):
        _capi.check_ray_arrays(np.zeros((5, 3)), np.zeros((5, 3)), np.zeros(4))
    with pytest.raises(ValueError):
        pt.Scene().setContents(pos, np.zeros((3, 9), np.float32), None, np.zeros((1, 8), np.float32))
    o, d, tm = _capi.check_ray_arrays(np.zeros((5, 3)), np.ones((5, 3)), np.ones(5))
    assert o.dtype == np.float32 and tm.shape == (5,)


@pytest.mark.parametrize("gamma", [2.2, 1.0, 2.4, 0.45])
def test_tonemap_thresholds_split_the_floats_like_the_reference_maths(built, gamma):
    """b2pt_tonemap_thresholds (host powf bisection): thr[k] is the first float whose reference byte is >= k."""
    from path_tracer_ai_b200 import _capi
    thr = _capi.tonemap_thresholds(gamma)

    def host(v):
        return reference_tonemap(v, gamma).astype(int)

    assert thr[0] == 0.0 and (np.diff(thr) >= 0).all()
    at = host(thr)
    below = host(np.maximum(thr.view(np.uint32).astype(np.int64) - 1, 0).astype(np.uint32).view(np.float32))
    for k in range(1, 256):
        assert at[k] >= k and below[k] < k, k
    x = np.random.default_rng(1).random(20000).astype(np.float32)
    assert np.array_equal(np.searchsorted(thr, x, side="right") - 1, host(x))


def test_camera_matches_oracle(built):
    cam = pt.Camera()
    want = PortOracle.camera()
    got = np.concatenate([cam.getPosition(), cam.getForward(), cam.getRight(), cam.getUp(), [cam.getFOV()]]).astype(np.float32)
    assert np.array_equal(bits(got), bits(want))
    cam2 = pt.Camera((1, 2, 3), (0.5, -1, 0.25), (0.1, 1, 0), 60.0)
    want2 = PortOracle.camera((1, 2, 3), (0.5, -1, 0.25), (0.1, 1, 0), 60.0)
    got2 = np.concatenate([cam2.getPosition(), cam2.getForward(), cam2.getRight(), cam2.getUp(), [60.0]]).astype(np.float32)
    assert np.array_equal(bits(got2), bits(want2))


def test_png_writer_roundtrip(built, tmp_path):
    from path_tracer_ai_b200.renderer import _host_lib
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (13, 37, 3), dtype=np.uint8)
    p = tmp_path / "x.png"
    assert _host_lib().b2pt_write_png(os.fsencode(str(p)), 37, 13, img.ctypes.data) == 0
    data = p.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(data):
        (ln,) = struct.unpack(">I", data[pos:pos + 4])
        typ, body = data[pos + 4:pos + 8], data[pos + 8:pos + 8 + ln]
        (crc,) = struct.unpack(">I", data[pos + 8 + ln:pos + 12 + ln])
        assert zlib.crc32(typ + body) == crc
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        if typ == b"IDAT":
            idat += body
        pos += 12 + ln
    assert ihdr == (37, 13, 8, 2, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(13, 1 + 37 * 3)
    assert (raw[:, 0] == 0).all() and np.array_equal(raw[:, 1:].reshape(13, 37, 3), img)


def cli():
    return os.path.join(os.path.dirname(pt.LIB_PATH), "b2pt_cli")


def test_cli_contract(built, tmp_path):
    """Flag names / defaults / exit codes of the reference's src/main.cpp:13-43, :114-117."""
    r = subprocess.run([cli(), "--help"], capture_output=True, text=True)
    assert r.returncode == 0
    for flag in ("-m, --mode", "-w, --width", "-h, --height", "-s, --samples", "-b, --bounces", "-g, --gamma", "-i, --input",
                 "-o, --output", "--help"):
        assert flag in r.stdout
    for default in ("gpu", "800", "450", "100", "5", "2.2", "IronMan/IronMan.obj", "output.png"):
        assert f"(default: {default})" in r.stdout
    r = subprocess.run([cli(), "--mode", "vulkan"], capture_output=True, text=True)
    assert r.returncode != 0 and "Invalid rendering mode" in r.stderr
    r = subprocess.run([cli(), "-m", "cpu"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    r = subprocess.run([cli(), "-i", str(tmp_path / "nope.obj")], capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to load model" in r.stderr
    r = subprocess.run([cli(), "--frobnicate"], capture_output=True, text=True)
    assert r.returncode != 0
    # extras beyond the reference's flags: camera / light overrides are validated before any GPU work
    for flag in ("--seed", "--camera-pos", "--camera-target", "--fov", "--lights", "--dump-float"):
        assert flag in subprocess.run([cli(), "--help"], capture_output=True, text=True).stdout
    obj = scenes.write_cornell_obj(str(tmp_path))
    r = subprocess.run([cli(), "-i", obj, "--lights", "1,2,3"], capture_output=True, text=True)
    assert r.returncode != 0 and "expects x,y,z,r,g,b,intensity" in r.stderr
    r = subprocess.run([cli(), "-i", obj, "--camera-pos", "north"], capture_output=True, text=True)
    assert r.returncode != 0 and "expects x,y,z" in r.stderr


def test_cli_without_gpu_fails_loudly(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    obj = scenes.write_cornell_obj(str(tmp_path))
    r = subprocess.run([cli(), "-i", obj, "-w", "16", "--height=9", "-s2"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
