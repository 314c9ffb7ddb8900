"""Synthetic scenes in the exporter's object-line vocabulary (plugin/rtsexport.py:312-314).

The reference's large scenes (the Stanford bunny, the million-triangle grid of bunnies, the city)
are not shipped (`.MISSING_LARGE_BLOBS`), so BASELINE.json's configs 2, 3 and 5 are rendered on
stand-ins with the same triangle counts and layout, generated here as `drb_object` arrays and
handed to the library in memory (or written with `write_rts` for the text-ingest path).

Conventions follow the exporter: world "up" is -y (it writes Blender (x, y, z) as (x, -z, y)),
`type` 2, `addional.x` 0, face normal + vertex normals + UVs always present.
Triangle edges stay >= ~0.05 units: the reference rejects |det| < 1e-4 (kernel.cu:291), so smaller
triangles become invisible to grazing rays on both sides (SURVEY.md hard part 5).
"""
from __future__ import annotations

import numpy as np

from . import OBJECT_DTYPE, Settings, default_settings, make_objects


def _tri_objects(v0, v1, v2, col, mat, rough, smooth=0, vn=None, add_x=0.0):
    """Object lines for triangles; v*: (n,3) float arrays; col (3,) or (n,3)."""
    n = len(v0)
    o = make_objects(n)
    o["pos"], o["dim"], o["rot"] = v0, v1, v2
    o["col"] = col
    o["mat"] = mat
    o["add_y"] = rough
    o["add_x"] = add_x
    fn = np.cross(v1 - v0, v2 - v0)
    ln = np.linalg.norm(fn, axis=1, keepdims=True)
    fn = np.where(ln > 0, fn / np.maximum(ln, 1e-30), np.array([0.0, -1.0, 0.0]))
    o["norm"] = fn.astype(np.float32)
    if vn is None:
        o["n1"] = o["n2"] = o["n3"] = o["norm"]
    else:
        o["n1"], o["n2"], o["n3"] = vn
    o["smooth"] = smooth
    o["ncols"] = 38
    return o


def grid_mesh(nu: int, nv: int, fn, wrap_u=True, wrap_v=True):
    """Triangulate a (nu x nv) parameter grid; fn(u, v) -> (positions (.,3), normals (.,3)).
    Returns v0, v1, v2, (n0, n1, n2) with 2*nu*nv triangles when both directions wrap."""
    iu = np.arange(nu if wrap_u else nu + 1)
    iv = np.arange(nv if wrap_v else nv + 1)
    U, V = np.meshgrid(iu / nu, iv / nv, indexing="ij")
    P, N = fn(U.ravel(), V.ravel())
    P = P.reshape(len(iu), len(iv), 3); N = N.reshape(len(iu), len(iv), 3)
    a = np.arange(nu); b = np.arange(nv)
    A, B = np.meshgrid(a, b, indexing="ij")
    A1 = (A + 1) % len(iu); B1 = (B + 1) % len(iv)
    p00, p10, p11, p01 = P[A, B], P[A1, B], P[A1, B1], P[A, B1]
    n00, n10, n11, n01 = N[A, B], N[A1, B], N[A1, B1], N[A, B1]
    v0 = np.concatenate([p00.reshape(-1, 3), p00.reshape(-1, 3)])
    v1 = np.concatenate([p10.reshape(-1, 3), p11.reshape(-1, 3)])
    v2 = np.concatenate([p11.reshape(-1, 3), p01.reshape(-1, 3)])
    vn = (np.concatenate([n00.reshape(-1, 3), n00.reshape(-1, 3)]),
          np.concatenate([n10.reshape(-1, 3), n11.reshape(-1, 3)]),
          np.concatenate([n11.reshape(-1, 3), n01.reshape(-1, 3)]))
    return v0.astype(np.float32), v1.astype(np.float32), v2.astype(np.float32), tuple(x.astype(np.float32) for x in vn)


def bumpy_torus(nu=256, nv=128, R=3.0, r=1.3, bumps=0.18, seed=0):
    """A closed, bumpy torus of 2*nu*nv triangles lying in the x-z plane (the "bunny-class" blob)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 4)

    def fn(u, v):
        a, b = 2 * np.pi * u, 2 * np.pi * v
        rr = r * (1.0 + bumps * np.sin(5 * a + ph[0]) * np.sin(3 * b + ph[1]) + 0.5 * bumps * np.sin(11 * a + 7 * b + ph[2]))
        cx, cz = np.cos(a), np.sin(a)
        x = (R + rr * np.cos(b)) * cx
        z = (R + rr * np.cos(b)) * cz
        y = rr * np.sin(b)
        p = np.stack([x, y, z], 1)
        n = np.stack([np.cos(b) * cx, np.sin(b), np.cos(b) * cz], 1)
        return p, n

    return grid_mesh(nu, nv, fn)


def quad(p00, p10, p11, p01, col, mat=0, rough=0.0):
    p = [np.asarray(x, np.float32)[None] for x in (p00, p10, p11, p01)]
    v0 = np.concatenate([p[0], p[0]]); v1 = np.concatenate([p[1], p[2]]); v2 = np.concatenate([p[2], p[3]])
    return _tri_objects(v0, v1, v2, col, mat, rough)


def instanced_grid_scene(grid=4, nu=256, nv=128, spacing=10.0, width=1920, height=1080, spp=256, max_depth=10, seed=0):
    """BASELINE config 3 stand-in: grid x grid bumpy tori in front of a wall on a floor, flattened
    (the format has no instancing), mixed diffuse / metal like images/millionstris.bmp.
    grid=4, nu=256, nv=128 -> 16 * 65 536 + 4 = 1 048 580 triangles."""
    rng = np.random.default_rng(seed)
    parts = []
    half = (grid - 1) * spacing / 2
    palette = np.array([[0.8, 0.8, 0.8], [0.8, 0.25, 0.2], [0.25, 0.5, 0.8], [0.85, 0.7, 0.3], [0.3, 0.7, 0.35]], np.float32)
    for i in range(grid):
        for j in range(grid):
            v0, v1, v2, vn = bumpy_torus(nu, nv, seed=seed + 17 * (i * grid + j))
            # tilt each instance a little so silhouettes differ
            ang = rng.uniform(-0.6, 0.6)
            c, s = np.cos(ang), np.sin(ang)
            Rm = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float32)
            off = np.array([i * spacing - half, -2.2, j * spacing - half], np.float32)
            v0, v1, v2 = v0 @ Rm.T + off, v1 @ Rm.T + off, v2 @ Rm.T + off
            vn = tuple(x @ Rm.T for x in vn)
            k = i * grid + j
            metal = (k % 3) == 0
            parts.append(_tri_objects(v0, v1, v2, palette[k % len(palette)], 3 if metal else 0, 0.05 if metal else 0.0,
                                      smooth=1 if (k % 2) else 0, vn=vn))
    ext = half + 3 * spacing
    parts.append(quad((-ext, 0, -ext), (ext, 0, -ext), (ext, 0, ext), (-ext, 0, ext), (0.75, 0.75, 0.75)))          # floor (y = 0, up is -y)
    parts.append(quad((-ext, 0, -half - spacing), (ext, 0, -half - spacing), (ext, -ext, -half - spacing), (-ext, -ext, -half - spacing),
                      (0.7, 0.7, 0.75)))                                                                              # back wall
    objs = np.concatenate(parts)
    st = default_settings()
    st.cam[:] = [0.0, -0.55 * (half + spacing) - 6.0, half + 2.2 * spacing]
    st.look[:] = [0.0, -2.0, 0.0]
    st.aperture = 0.01
    st.focus = 3.0
    st.fov = 45
    st.max_depth = max_depth
    st.spp = spp
    st.width, st.height = width, height
    return objs, st


def bunny_class_scene(width=1920, height=1080, spp=64, max_depth=8, nu=320, nv=128):
    """BASELINE config 2 stand-in (samples/sanford.blend.rts is a missing blob): three blobs of
    2*nu*nv triangles each (metal / white / red like images/sanfordnew.bmp) + floor + wall."""
    parts = []
    mats = [(3, 0.02, (0.9, 0.9, 0.9)), (0, 0.0, (0.85, 0.85, 0.85)), (0, 0.0, (0.8, 0.15, 0.12))]
    for k, (mat, rough, col) in enumerate(mats):
        v0, v1, v2, vn = bumpy_torus(nu, nv, seed=100 + k)
        off = np.array([(k - 1) * 9.5, -2.2, 0.0], np.float32)
        parts.append(_tri_objects(v0 + off, v1 + off, v2 + off, np.array(col, np.float32), mat, rough, smooth=1, vn=vn))
    parts.append(quad((-60, 0, -60), (60, 0, -60), (60, 0, 60), (-60, 0, 60), (0.75, 0.75, 0.75)))
    parts.append(quad((-60, 0, -9), (60, 0, -9), (60, -60, -9), (-60, -60, -9), (0.7, 0.7, 0.75)))
    objs = np.concatenate(parts)
    st = default_settings()
    st.cam[:] = [0.0, -9.0, 24.0]
    st.look[:] = [0.0, -2.0, 0.0]
    st.max_depth, st.spp, st.width, st.height = max_depth, spp, width, height
    return objs, st


def city_scene(blocks=200, width=3840, height=2160, spp=1024, max_depth=10, seed=3, target_tris=10_000_000):
    """BASELINE config 5 stand-in: a blocks x blocks grid of box / pyramid buildings whose faces are
    tessellated so that the total is close to `target_tris` (< 2^24, ids stay exact in a float)."""
    rng = np.random.default_rng(seed)
    nb = blocks * blocks
    per_building = max(10, target_tris // nb)
    # a box with 5 visible faces, each face an m x m grid of quads: 10 m^2 triangles
    m = max(1, int(np.sqrt(per_building / 10.0)))
    pitch = 4.0
    gx, gz = np.meshgrid(np.arange(blocks), np.arange(blocks), indexing="ij")
    cx = (gx.ravel() - blocks / 2) * pitch; cz = (gz.ravel() - blocks / 2) * pitch
    hx = rng.uniform(0.9, 1.6, nb); hz = rng.uniform(0.9, 1.6, nb); hy = rng.uniform(2.0, 14.0, nb)
    cols = rng.uniform(0.3, 0.9, (nb, 3)).astype(np.float32)
    matid = np.where(rng.uniform(size=nb) < 0.15, 3, 0)
    a = np.arange(m) / m; b = (np.arange(m) + 1) / m
    A0, B0 = np.meshgrid(a, a, indexing="ij"); A1, B1 = np.meshgrid(b, b, indexing="ij")
    s0, t0, s1, t1 = A0.ravel(), B0.ravel(), A1.ravel(), B1.ravel()          # m*m cells in [0,1]^2

    def face(origin, eu, ev):
        """origin, eu, ev: (nb,3); returns v0,v1,v2 for 2*m*m*nb triangles"""
        def pt(s, t):
            return origin[:, None, :] + s[None, :, None] * eu[:, None, :] + t[None, :, None] * ev[:, None, :]
        p00, p10, p11, p01 = pt(s0, t0), pt(s1, t0), pt(s1, t1), pt(s0, t1)
        v0 = np.concatenate([p00, p00], 1); v1 = np.concatenate([p10, p11], 1); v2 = np.concatenate([p11, p01], 1)
        return v0.reshape(-1, 3), v1.reshape(-1, 3), v2.reshape(-1, 3)

    z = np.zeros(nb)
    lo = np.stack([cx - hx, z, cz - hz], 1); top = -hy                          # up is -y
    ex = np.stack([2 * hx, z, z], 1); ez = np.stack([z, z, 2 * hz], 1); ey = np.stack([z, top, z], 1)
    faces = [
        face(lo + ey, ex, ez),                         # roof
        face(lo, ex, ey), face(lo + ez, ex, ey),       # front / back
        face(lo, ez, ey), face(lo + ex, ez, ey),       # left / right
    ]
    tri_per_face = 2 * m * m
    parts = []
    for v0, v1, v2 in faces:
        col = np.repeat(cols, tri_per_face, 0)
        mat = np.repeat(matid, tri_per_face)
        rough = np.where(mat == 3, 0.08, 0.0).astype(np.float32)
        parts.append(_tri_objects(v0.astype(np.float32), v1.astype(np.float32), v2.astype(np.float32), col, mat, rough))
    ext = blocks * pitch
    parts.append(quad((-ext, 0, -ext), (ext, 0, -ext), (ext, 0, ext), (-ext, 0, ext), (0.5, 0.5, 0.5)))
    objs = np.concatenate(parts)
    st = default_settings()
    st.cam[:] = [0.0, -0.35 * ext, 0.75 * ext]
    st.look[:] = [0.0, -5.0, 0.0]
    st.focus = 3.0
    st.max_depth, st.spp, st.width, st.height = max_depth, spp, width, height
    return objs, st


def heightfield_scene(n=64, size=8.0, amp=0.6, width=256, height=256, spp=4, max_depth=6, seed=1, metal_every=5):
    """Small test scene: an n x n heightfield (2 n^2 triangles) with a few material classes."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 6.28, 3)

    def fn(u, v):
        x = (u - 0.5) * size; z = (v - 0.5) * size
        y = -amp * (np.sin(3 * x + ph[0]) * np.cos(2 * z + ph[1]) + 0.5 * np.sin(5 * z + ph[2]))
        p = np.stack([x, y, z], 1)
        n_ = np.stack([np.zeros_like(x), -np.ones_like(x), np.zeros_like(x)], 1)
        return p, n_

    v0, v1, v2, vn = grid_mesh(n, n, fn, wrap_u=False, wrap_v=False)
    k = np.arange(len(v0))
    mat = np.where(k % metal_every == 0, 3, 0)
    col = np.stack([0.4 + 0.5 * ((k * 7) % 11) / 11, 0.4 + 0.5 * ((k * 3) % 13) / 13, 0.5 + 0 * k], 1).astype(np.float32)
    objs = _tri_objects(v0, v1, v2, col, mat, np.where(mat == 3, 0.1, 0.0).astype(np.float32))
    st = default_settings()
    st.cam[:] = [0.0, -5.0, 7.0]
    st.look[:] = [0.0, 0.0, 0.0]
    st.max_depth, st.spp, st.width, st.height = max_depth, spp, width, height
    return objs, st


def write_test_textures(directory: str):
    """Two small procedural .ppm files (P6, as plugin/rtsexport.py:56-79 writes them): an environment map and a
    colour/roughness map.  Returns their paths."""
    import os
    yy, xx = np.mgrid[0:256, 0:512]
    env = np.stack([120 + 100 * np.sin(xx / 40.0), 140 + 90 * np.cos(yy / 30.0), 200 + 40 * np.sin((xx + yy) / 25.0)], -1)
    env[(yy < 40) & (np.abs(xx - 256) < 40)] = 255                          # a bright "sun"
    tex = np.stack([(xx[:256, :256] // 16 + yy[:256, :256] // 16) % 2 * 180 + 40, 80 + (xx[:256, :256] % 64) * 2, 200 - (yy[:256, :256] % 32) * 4], -1)
    paths = []
    for name, img in (("synth_env.ppm", env), ("synth_tex.ppm", tex)):
        p = os.path.join(directory, name)
        a = np.clip(img, 0, 255).astype(np.uint8)
        with open(p, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (a.shape[1], a.shape[0]) + a.tobytes())
        paths.append(p)
    return paths


def materials_scene(tex_paths, width=1920, height=1080, spp=256, max_depth=10, nu=96, nv=48):
    """BASELINE config 4 stand-in (mats + glass + texer with an environment map): a row of blobs, one per material
    class the reference has -- diffuse, mirror (2), metal (3), glass (4), glossy (5), emissive (1), textured,
    checker, roughness-mapped metal, smooth-shaded -- on a checkered floor under an environment map.
    tex_paths: [environment.ppm, texture.ppm] (write_test_textures)."""
    specs = [
        dict(mat=0, col=(0.8, 0.8, 0.8)), dict(mat=2, col=(0.95, 0.95, 0.95)), dict(mat=3, col=(0.9, 0.7, 0.3), rough=0.15),
        dict(mat=4, col=(1.0, 1.0, 1.0), rough=1.45), dict(mat=5, col=(0.3, 0.6, 0.9), rough=0.05), dict(mat=1, col=(3.0, 2.4, 1.8)),
        dict(mat=0, col=(0.8, 0.8, 0.8), tex=1), dict(mat=0, col=(0.2, 0.7, 0.3), checker=1), dict(mat=3, col=(0.9, 0.9, 0.9), rough=0.3, rtex=1),
        dict(mat=0, col=(0.85, 0.3, 0.3), smooth=1),
    ]
    parts = []
    for k, sp in enumerate(specs):
        v0, v1, v2, vn = bumpy_torus(nu, nv, R=1.6, r=0.7, seed=40 + k)
        off = np.array([(k % 5 - 2) * 5.5, -1.2, (k // 5) * 6.0 - 3.0], np.float32)
        o = _tri_objects(v0 + off, v1 + off, v2 + off, np.array(sp["col"], np.float32), sp["mat"], sp.get("rough", 0.0),
                         smooth=sp.get("smooth", 1 if sp["mat"] == 4 else 0), vn=vn)
        # UVs from the parameter grid so that textures / checker have something to index
        u = (np.arange(len(o)) % (nu * nv)) / float(nu * nv)
        o["t1"] = np.stack([u, (u * 7) % 1.0], 1); o["t2"] = o["t1"] + (0.01, 0.0); o["t3"] = o["t1"] + (0.0, 0.01)
        if sp.get("tex"):
            o["texnum"] = 1
        if sp.get("rtex"):
            o["rtexnum"] = 1
        if sp.get("checker"):
            o["checker"] = 1
        parts.append(o)
    fl = quad((-40, 0, -40), (40, 0, -40), (40, 0, 40), (-40, 0, 40), (0.6, 0.6, 0.6))
    fl["checker"] = 1
    fl["t1"] = (0, 0); fl["t2"] = (4, 0); fl["t3"] = (4, 4)
    parts.append(fl)
    objs = np.concatenate(parts)
    st = default_settings()
    st.cam[:] = [0.0, -7.0, 17.0]
    st.look[:] = [0.0, -1.0, 0.0]
    st.backtex = 0
    st.bg_intensity = 1.0
    st.max_depth, st.spp, st.width, st.height = max_depth, spp, width, height
    return objs, st, list(tex_paths)
