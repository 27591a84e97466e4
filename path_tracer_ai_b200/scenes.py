"""Procedural scenes for the BASELINE.json configurations (synthetic data, seeded).

* :func:`write_cornell_obj`  — configs 1/2: a ~42-triangle Cornell-style box (diffuse walls, rough-specular tall
  box, glass short box, diffuse wedge) written as OBJ+MTL for ``Scene::loadFromObj``.  The tall box defaults to
  roughness 0.4: the reference's direct term for SPECULAR is albedo * D_GGX with no normalisation
  (renderer.hpp:286-290), whose peak 1/(pi*rough^4) is ~3000 at roughness 0.1 — an estimator so heavy-tailed that
  the reference does not agree with ITSELF to 30 % at 1024 spp; 0.4 keeps statistical parity tests meaningful.
* :func:`mesh_scene`         — configs 3/5: displaced terrain + tessellated, noise-displaced tori / spheres with
  mixed materials, generated directly as triangle arrays in world space.
* :func:`random_rays`        — config 4: origins uniform in the inflated scene box, directions uniform on the sphere.
* :func:`random_soup`        — small random triangles (unit tests).

Material names use the documented MTL prefix extension (``diffuse*``, ``glass*``, ``mirror*``, ``rough<value>*``;
see DESIGN.md) because the reference loader can only produce SPECULAR materials.
"""
from __future__ import annotations

import os

import numpy as np

DIFFUSE, SPECULAR, DIELECTRIC = 0, 1, 2


def random_soup(n: int, seed: int = 1234, extent: float = 1.0, size: float = 0.08) -> np.ndarray:
    rng = np.random.default_rng(seed)
    c = (rng.random((n, 1, 3)) * 2 - 1) * extent
    pos = c + (rng.random((n, 3, 3)) - 0.5) * size
    return pos.reshape(n, 9).astype(np.float32)


def random_rays(n: int, lo, hi, seed: int = 1234, inflate: float = 0.1):
    """Config 4 rays: origin uniform in the scene AABB inflated by `inflate`, direction uniform on the sphere
    (z = 1 - 2*xi1, phi = 2*pi*xi2)."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    ext = (hi - lo) * inflate
    lo, hi = lo - ext, hi + ext
    o = lo + rng.random((n, 3)) * (hi - lo)
    z = 1.0 - 2.0 * rng.random(n)
    phi = 2.0 * np.pi * rng.random(n)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    return o.astype(np.float32), d.astype(np.float32)


# ------------------------------------------------------------------------------------------------------
# Cornell box (configs 1 and 2)
# ------------------------------------------------------------------------------------------------------
def _box(cx, cz, sx, sy, sz, y0, angle_deg):
    """Axis box rotated about y; returns 8 vertices (post-flip coords) and 12 CCW-outward faces."""
    a = np.radians(angle_deg)
    ca, sa = np.cos(a), np.sin(a)
    v = []
    for dy in (0.0, sy):
        for dx, dz in ((-sx, -sz), (sx, -sz), (sx, sz), (-sx, sz)):
            x = cx + ca * dx + sa * dz
            z = cz - sa * dx + ca * dz
            v.append((x, y0 + dy, z))
    f = [(0, 1, 2), (0, 2, 3),        # bottom (normal -y)
         (4, 6, 5), (4, 7, 6),        # top (+y)
         (0, 5, 1), (0, 4, 5),        # -z side
         (1, 6, 2), (1, 5, 6),        # +x side
         (2, 7, 3), (2, 6, 7),        # +z side
         (3, 4, 0), (3, 7, 4)]        # -x side
    return np.array(v, np.float64), f


def _wedge(cx, cz, s, h, y0, angle_deg):
    a = np.radians(angle_deg)
    ca, sa = np.cos(a), np.sin(a)
    base = [(-s, -s), (s, -s), (0.0, s)]
    v = []
    for dy in (0.0, h):
        for dx, dz in base:
            v.append((cx + ca * dx + sa * dz, y0 + dy, cz - sa * dx + ca * dz))
    f = [(0, 1, 2), (3, 5, 4),
         (0, 4, 1), (0, 3, 4), (1, 5, 2), (1, 4, 5), (2, 3, 0), (2, 5, 3)]
    return np.array(v, np.float64), f


def cornell_geometry(seed: int = 1234, steel: str = "rough0.4_steel"):
    """Returns (vertices[n,3] in post-flip model coords, faces[(i,j,k)], face material names, materials dict)."""
    rng = np.random.default_rng(seed)
    verts, faces, fmat = [], [], []

    def add(v, f, name):
        base = len(verts)
        verts.extend(v.tolist())
        for (i, j, k) in f:
            faces.append((base + i, base + j, base + k))
            fmat.append(name)

    # room shell: x in [-1,1], y in [-1,1], z in [-0.6,0.6]; open toward +z (the camera side)
    X, Y, Z = 1.0, 1.0, 0.6
    shell = np.array([(-X, -Y, -Z), (X, -Y, -Z), (X, -Y, Z), (-X, -Y, Z),
                      (-X, Y, -Z), (X, Y, -Z), (X, Y, Z), (-X, Y, Z)], np.float64)
    add(shell, [(0, 2, 1), (0, 3, 2)], "diffuse_white")      # floor, normal +y
    add(shell, [(4, 5, 6), (4, 6, 7)], "diffuse_white")      # ceiling, normal -y
    add(shell, [(0, 1, 5), (0, 5, 4)], "diffuse_white")      # back wall, normal +z
    add(shell, [(0, 4, 7), (0, 7, 3)], "diffuse_red")        # left wall, normal +x
    add(shell, [(1, 2, 6), (1, 6, 5)], "diffuse_green")      # right wall, normal -x
    v, f = _box(-0.35, -0.15, 0.28, 1.15, 0.22, -Y + 1e-3, 17.0)
    add(v, f, steel)
    v, f = _box(0.38, 0.12, 0.26, 0.55, 0.24, -Y + 1e-3, -19.0)
    add(v, f, "glass_box")
    v, f = _wedge(-0.05, 0.36, 0.16, 0.3, -Y + 1e-3, 33.0)
    add(v, f, "diffuse_blue")

    V = np.array(verts, np.float64)
    # a fraction of a degree about y and x plus a small jitter: no reference BVH node is flat
    ay, ax = np.radians(0.37), np.radians(0.21)
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    V = V @ Ry.T @ Rx.T
    V += (rng.random(V.shape) - 0.5) * 2e-4
    materials = {
        "diffuse_white": dict(Kd=(0.73, 0.73, 0.73)),
        "diffuse_red": dict(Kd=(0.65, 0.05, 0.05)),
        "diffuse_green": dict(Kd=(0.12, 0.45, 0.15)),
        "diffuse_blue": dict(Kd=(0.15, 0.25, 0.7)),
        steel: dict(Kd=(0.8, 0.8, 0.85)),
        "glass_box": dict(Kd=(1.0, 1.0, 1.0), Ni=1.5),
    }
    return V, faces, fmat, materials


def write_cornell_obj(directory: str, seed: int = 1234, name: str = "cornell", steel: str = "rough0.4_steel") -> str:
    """Writes <name>.obj/.mtl.  File z is the NEGATED model z because Scene::loadFromObj flips z
    (reference src/scene.cpp:237): after loading, the box opens toward the camera and face normals
    (computed by the loader from the transformed vertices, :251-256) point outward."""
    V, faces, fmat, materials = cornell_geometry(seed, steel)
    os.makedirs(directory, exist_ok=True)
    obj_path = os.path.join(directory, name + ".obj")
    with open(os.path.join(directory, name + ".mtl"), "w") as f:
        for mname, m in materials.items():
            f.write(f"newmtl {mname}\nKd {m['Kd'][0]:.6f} {m['Kd'][1]:.6f} {m['Kd'][2]:.6f}\n")
            if "Ni" in m:
                f.write(f"Ni {m['Ni']:.4f}\n")
            f.write("\n")
    with open(obj_path, "w") as f:
        f.write(f"mtllib {name}.mtl\n")
        for x, y, z in V:
            f.write(f"v {x:.9g} {y:.9g} {-z:.9g}\n")
        cur = None
        for (i, j, k), m in zip(faces, fmat):
            if m != cur:
                f.write(f"usemtl {m}\n")
                cur = m
            f.write(f"f {i + 1} {j + 1} {k + 1}\n")
    return obj_path


# ------------------------------------------------------------------------------------------------------
# Tessellated mesh scene (configs 3 and 5)
# ------------------------------------------------------------------------------------------------------
def _grid_mesh(P: np.ndarray, wrap_u: bool, wrap_v: bool):
    """P: (nu, nv, 3) vertex grid -> (pos[n,9], nrm[n,9]) with smooth (area-weighted) vertex normals."""
    nu, nv, _ = P.shape
    iu = np.arange(nu if wrap_u else nu - 1)
    iv = np.arange(nv if wrap_v else nv - 1)
    I, J = np.meshgrid(iu, iv, indexing="ij")
    I1, J1 = (I + 1) % nu, (J + 1) % nv
    a = I * nv + J
    b = I1 * nv + J
    c = I1 * nv + J1
    d = I * nv + J1
    tris = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)], 0)
    V = P.reshape(-1, 3)
    fn = np.cross(V[tris[:, 1]] - V[tris[:, 0]], V[tris[:, 2]] - V[tris[:, 0]])
    vn = np.zeros_like(V)
    for k in range(3):
        np.add.at(vn, tris[:, k], fn)
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    vn = vn / np.maximum(ln, 1e-20)
    pos = V[tris].reshape(-1, 9)
    nrm = vn[tris].reshape(-1, 9)
    return pos, nrm


def _noise(u, v, rng, octaves=4):
    out = np.zeros_like(u)
    for o in range(octaves):
        fu, fv = rng.integers(1, 6, 2) * (o + 1)
        ph = rng.random(2) * 2 * np.pi
        out += np.sin(fu * u + ph[0]) * np.cos(fv * v + ph[1]) / (o + 1) ** 1.5
    return out


def mesh_scene(ntri: int = 1_000_000, seed: int = 1234, dielectric_fraction: float = 0.15, room: bool = True):
    """Returns dict(pos[n,9], nrm[n,9], mat[n], materials8[m,8], lo, hi) in world space (pre-build order).

    ~32 % of the triangles form a displaced terrain, the rest 8 displaced tori / spheres.  Material mix by
    object: mostly diffuse, some rough/perfect specular, `dielectric_fraction` of the objects' triangles glass."""
    rng = np.random.default_rng(seed)
    n_terrain = int(ntri * 0.32)
    g = max(2, int(np.sqrt(n_terrain / 2)))
    n_obj = 8
    per_obj = (ntri - 2 * g * g) // n_obj
    m = max(3, int(np.sqrt(per_obj / 2)))
    parts = []
    materials = [
        (DIFFUSE, (0.7, 0.7, 0.68), 0.95, 0.0, 1.5),     # 0 terrain
        (DIFFUSE, (0.75, 0.25, 0.2), 0.95, 0.0, 1.5),    # 1
        (DIFFUSE, (0.2, 0.55, 0.3), 0.95, 0.0, 1.5),     # 2
        (DIFFUSE, (0.25, 0.35, 0.75), 0.95, 0.0, 1.5),   # 3
        (SPECULAR, (0.9, 0.85, 0.7), 0.0, 1.0, 1.5),     # 4 mirror
        (SPECULAR, (0.85, 0.85, 0.9), 0.1, 1.0, 1.5),    # 5 rough 0.1
        (SPECULAR, (0.9, 0.6, 0.3), 0.3, 1.0, 1.5),      # 6 rough 0.3
        (DIELECTRIC, (1.0, 1.0, 1.0), 0.0, 0.0, 1.5),    # 7 glass
    ]
    # terrain over x,z in [-1.5,1.5], y around 0.4
    u = np.linspace(0, 2 * np.pi, g + 1)
    U, Vv = np.meshgrid(u, u, indexing="ij")
    h = 0.12 * _noise(U, Vv, rng)
    P = np.stack([(U / np.pi - 1.0) * 1.5, 0.45 + h, (Vv / np.pi - 1.0) * 1.5], -1)
    P += (rng.random(P.shape) - 0.5) * 2e-5
    pos, nrm = _grid_mesh(P, False, False)
    if nrm.reshape(-1, 3)[:, 1].mean() < 0:   # normals point up
        nrm = -nrm
    parts.append((pos, nrm, 0))
    # dielectric_fraction is of ALL triangles; the 8 objects carry ~68 % of them
    n_glass = min(max(int(round(n_obj * dielectric_fraction / 0.68)), 1), n_obj)
    obj_mats = [7] * n_glass + [4, 5, 6, 1, 2, 3, 1, 2][: n_obj - n_glass]
    for k in range(n_obj):
        ang = 2 * np.pi * k / n_obj + 0.3
        rad = 0.55 + 0.45 * (k % 2)
        centre = np.array([rad * np.cos(ang), 1.1 + 0.55 * ((k * 5) % 4) / 3.0, rad * np.sin(ang)])
        uu = np.linspace(0, 2 * np.pi, m, endpoint=False)
        if k % 2 == 0:   # torus (wraps both ways)
            U, Vv = np.meshgrid(uu, uu, indexing="ij")
            R, r = 0.3, 0.11 * (1.0 + 0.25 * _noise(U, Vv, rng))
            P = np.stack([(R + r * np.cos(Vv)) * np.cos(U), r * np.sin(Vv), (R + r * np.cos(Vv)) * np.sin(U)], -1)
            tilt = rng.random(3) * np.pi
            wrap = (True, True)
        else:            # sphere (wraps in u, poles left open by a hair to avoid degenerate triangles)
            vv = np.linspace(0.02, np.pi - 0.02, m)
            U, Vv = np.meshgrid(uu, vv, indexing="ij")
            r = 0.3 * (1.0 + 0.18 * _noise(U, 2 * Vv, rng))
            P = np.stack([r * np.sin(Vv) * np.cos(U), r * np.cos(Vv), r * np.sin(Vv) * np.sin(U)], -1)
            tilt = rng.random(3) * np.pi
            wrap = (True, False)
        cx, sx = np.cos(tilt[0]), np.sin(tilt[0])
        cz, sz = np.cos(tilt[2]), np.sin(tilt[2])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        P = P @ Rx.T @ Rz.T + centre
        P += (rng.random(P.shape) - 0.5) * 2e-5
        pos, nrm = _grid_mesh(P, *wrap)
        # orient normals outward (away from the object's centre line / centre)
        cen = pos.reshape(-1, 3, 3).mean(1)
        outward = cen - centre
        fn = nrm.reshape(-1, 3, 3).mean(1)
        if np.mean(np.sum(fn * outward, 1)) < 0:
            nrm = -nrm
        parts.append((pos, nrm, obj_mats[k]))
    if room:
        # the 8 room triangles Scene::loadFromObj always adds (reference src/scene.cpp:118-209): floor 16x16 at
        # y=0, back / left / right walls of height 4, diffuse 0.9; first in the pre-build order, as there
        R, Hh = 8.0, 4.0
        quads = [((-R, 0, -R), (R, 0, -R), (R, 0, R), (-R, 0, R), (0, 1, 0)),
                 ((-R, 0, -R), (-R, Hh, -R), (R, Hh, -R), (R, 0, -R), (0, 0, 1)),
                 ((-R, 0, -R), (-R, 0, R), (-R, Hh, R), (-R, Hh, -R), (1, 0, 0)),
                 ((R, 0, -R), (R, Hh, -R), (R, Hh, R), (R, 0, R), (-1, 0, 0))]
        rp, rn = [], []
        for a, b, c, d, n in quads:
            rp += [a + b + c, a + c + d]
            rn += [n * 3, n * 3]
        materials.append((DIFFUSE, (0.9, 0.9, 0.9), 0.95, 0.0, 1.5))
        parts.insert(0, (np.array(rp, np.float64), np.array(rn, np.float64), len(materials) - 1))
    pos = np.concatenate([p for p, _, _ in parts]).astype(np.float32)
    nrm = np.concatenate([n for _, n, _ in parts]).astype(np.float32)
    mat = np.concatenate([np.full(len(p), mid, np.int32) for p, _, mid in parts])
    m8 = np.zeros((len(materials), 8), np.float32)
    for i, (ty, alb, rough, metal, ior) in enumerate(materials):
        m8[i] = (ty, alb[0], alb[1], alb[2], rough, metal, ior, 0.0)
    V = pos[8:].reshape(-1, 3) if room else pos.reshape(-1, 3)   # lo/hi: the mesh itself (ray generators), not the 16-unit room
    return dict(pos=pos, nrm=nrm, mat=mat, materials8=m8, lo=V.min(0), hi=V.max(0))
