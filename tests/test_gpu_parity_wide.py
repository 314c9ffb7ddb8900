"""Wider parity sweep: every shipped sample scene, and randomised scenes, against the oracle (bit-exact bars as in
test_gpu_parity.py)."""
import os

import numpy as np
import pytest

import dogeray_b200 as drb
from oracle import restated
from conftest import HAVE_REF, SAMPLES, all_sample_scenes, needs_ref, sample
from test_gpu_parity import Oracle, assert_frames_match, assert_ids_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def maybe_ref():
    if HAVE_REF:
        from oracle import refhost
        return refhost.RefHost()
    return None


def renderable_count(path):
    o = drb.HostScene.load(path, SAMPLES).objects()
    return int((((o["type"] == 2) & (o["ncols"] >= 16)) | ((o["type"] == 0) & (o["ncols"] >= 10))).sum())


@needs_ref
@pytest.mark.parametrize("name", all_sample_scenes())
def test_every_sample_scene(maybe_ref, name):
    path = sample(name)
    if os.path.getsize(path) == 0 or renderable_count(path) < 2:
        pytest.skip("the reference's build_bvh does not terminate on fewer than two objects")
    o = drb.HostScene.load(path, SAMPLES).objects()
    if ((o["type"] != 0) & (o["type"] != 2)).any():
        pytest.skip("legacy object types: undefined behaviour in the reference (SURVEY App. B.9)")
    sc = drb.Scene.load(path, SAMPLES)
    st = sc.settings.replace(width=64, height=48, spp=2, max_depth=5)
    orc = Oracle(path, SAMPLES, maybe_ref)
    orc.apply(st, 31)
    ro, rd = sc.primary_rays(st, sample=0, seed=31)
    ids, t = sc.trace_ids(ro, rd)
    oid, ot = orc.hit(ro, rd)
    # junk lines (HIGH.rts, light.rts) leave an indeterminate object in the reference's tree; ids of real objects must agree
    assert_ids_match(ids, t, oid, ot)
    f, fi, rays = orc.frame()
    acc, stats = sc.render(st, seed=31)
    assert stats.rays == rays
    assert_frames_match(acc.transpose(1, 0, 2) * np.float32(255.0) * np.float32(0.5), f)


def random_scene(rng, n, scale, spheres=True, duplicates=True):
    o = drb.make_objects(n)
    c = rng.uniform(-1, 1, (n, 3)) * scale
    o["pos"] = c
    o["dim"] = c + rng.normal(size=(n, 3)) * scale * 0.3
    o["rot"] = c + rng.normal(size=(n, 3)) * scale * 0.3
    o["col"] = rng.uniform(0.1, 1.0, (n, 3))
    o["mat"] = rng.choice([0, 0, 0, 1, 2, 3, 4, 5, 7], n)
    o["add_y"] = np.where(o["mat"] == 4, rng.uniform(1.1, 1.8, n), rng.uniform(0, 0.5, n))
    o["add_x"] = rng.choice([0.0, 0.5], n)
    fn = np.cross(o["dim"] - o["pos"], o["rot"] - o["pos"])
    fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-20)
    has_n = rng.uniform(size=n) < 0.7
    o["norm"][has_n] = fn[has_n]
    vn = fn + rng.normal(size=(n, 3)) * 0.2
    for k in ("n1", "n2", "n3"):
        o[k][has_n] = (vn + rng.normal(size=(n, 3)) * 0.05)[has_n]
    o["smooth"] = rng.integers(0, 2, n)
    o["checker"] = (rng.uniform(size=n) < 0.2).astype(np.int32)
    o["t1"], o["t2"], o["t3"] = rng.uniform(0, 3, (n, 2)), rng.uniform(0, 3, (n, 2)), rng.uniform(0, 3, (n, 2))
    sph = (rng.uniform(size=n) < 0.1) & spheres                       # some spheres
    o["type"][sph] = 0
    o["dim"][sph, 0] = rng.uniform(0.05, 0.4, int(sph.sum())) * scale
    o["mat"][sph & (o["mat"] == 4)] = 3                               # glass spheres need an inside hit the reference does not have
    o["checker"][sph] = 0                                             # getnormal leaves texco uninitialised for spheres (kernel.cu:707-710): UB with the checker
    deg = rng.uniform(size=n) < 0.05                                  # zero-area triangles
    o["rot"][deg] = o["dim"][deg]
    dup = (rng.uniform(size=n) < 0.05) & duplicates                   # exact duplicates (all keys tie, equal t)
    src = rng.integers(0, n, n)
    for k in ("pos", "dim", "rot", "type"):
        o[k][dup] = o[k][src[dup]]
    o["ncols"] = 38
    return o


@pytest.mark.parametrize("seed", range(12))
def test_random_scenes(tmp_path, maybe_ref, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([3, 17, 64, 300, 1500]))
    scale = float(rng.choice([0.5, 3.0, 40.0]))
    # seeds 0-5: triangles only, no exact duplicates -> the strict bar; 6-8: + spheres; 9-11: + duplicated objects
    spheres, duplicates = seed >= 6, seed >= 9
    objs = random_scene(rng, n, scale, spheres, duplicates)
    st = drb.default_settings().replace(cam=(0.3 * scale, -0.2 * scale, 2.8 * scale), look=(0, 0, 0), width=72, height=48, spp=2, max_depth=6,
                                        focus=3.0, aperture=float(rng.choice([0.0, 0.01, 0.2])), fov=int(rng.choice([30, 45, 70])))
    p = str(tmp_path / "r.rts")
    drb.write_rts(p, st, objs)
    sc = drb.Scene.load(p)
    # closest hit: camera rays and random rays against the BVH-independent brute-force definition
    r = restated.Restated(p)
    ro, rd = sc.primary_rays(st, sample=0, seed=seed)
    o2 = rng.uniform(-2, 2, (4000, 3)).astype(np.float32) * scale
    d2 = rng.normal(size=(4000, 3)).astype(np.float32)
    allo = np.concatenate([ro.reshape(-1, 3), o2]); alld = np.concatenate([rd.reshape(-1, 3), d2])
    ids, t = sc.trace_ids(allo, alld)
    bid, bt = r.hit_brute(allo, alld)
    assert_ids_match(ids, t, bid, bt)
    # radiance against the oracle frame (same tree-independent semantics); duplicates make exact-t ties, whose winner
    # depends on the tree, so frames are compared only when the scene has no such tie among primary hits
    orc = Oracle(p, "", maybe_ref)
    orc.apply(st, seed)
    oid, ot = orc.hit(ro, rd)
    f, fi, rays = orc.frame()
    acc, stats = sc.render(st, seed=seed)
    ours = acc.transpose(1, 0, 2) * np.float32(255.0) * np.float32(0.5)
    identical = float(np.mean(np.all(ours == f, axis=-1)))
    rmse = float(np.sqrt(np.mean(((ours - f) / 255.0) ** 2)))
    assert np.isfinite(acc).all()
    if not spheres:
        # nothing tree-dependent, nothing library-dependent: the stated tolerance and the bit-identity bar
        assert stats.rays == rays
        assert identical >= 0.999 and rmse <= 1e-3, (identical, rmse)
    elif not duplicates:
        # hit_sphere squares lengths with powf(len, 2) (kernel.cu:320, 324); the device squares with one multiply, which is
        # the correctly rounded value glibc's powf misses by an ulp now and then -- a path may then take another branch
        # (at 2 spp one such path in a 72x48 image is already 5e-3 of RMSE).  Nearly every pixel is still bit-identical,
        # and the tolerance north_star states for CONVERGED radiance holds with room to spare at 64 spp.
        assert abs(stats.rays - rays) <= 0.002 * rays and identical >= 0.99, (stats.rays, rays, identical)
        st64 = st.replace(spp=64)
        orc.apply(st64, seed)
        f64, _, _ = orc.frame()
        acc64, _ = sc.render(st64, seed=seed)
        rmse64 = float(np.sqrt(np.mean(((acc64.transpose(1, 0, 2) * np.float32(255.0) * np.float32(1.0 / 64) - f64) / 255.0) ** 2)))
        assert rmse64 <= 1e-3, rmse64
    else:
        # exact-t ties on duplicated objects keep the first one VISITED, which depends on the tree (the reference's median
        # split vs the LBVH): the twin may carry another material.  ids were checked above under the tie rule.
        assert identical >= 0.97 and rmse < 0.02, (identical, rmse)
