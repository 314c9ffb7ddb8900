"""BASELINE.json's configurations AT THEIR STATED SIZES.

  * configs 2, 3 and 4 (stand-ins): one full 1920x1080 frame at 1-2 spp and the config's own depth through
    drb_render against the oracle's frame of the same samples (oracle/_ref = the reference's Kernel() compiled for the
    host, else the C restatement): equal ray counts, >= 99.9 % bit-identical pixels, RMSE <= 1e-3;
  * config 5 (10 M triangles, 3840x2160): closest-hit ids and distances against the brute-force definition on primary and
    random incoherent rays -- the scene where the 16-bit box grid is coarsest;
  * size-independent properties of the 1 M-triangle frame: shard additivity, determinism, ray count bookkeeping."""
import os

import numpy as np
import pytest

import dogeray_b200 as drb
from dogeray_b200 import synth
from oracle import restated
from conftest import HAVE_REF
from test_gpu_parity import Oracle, assert_frames_match, assert_ids_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def maybe_ref():
    if HAVE_REF:
        from oracle import refhost
        return refhost.RefHost()
    return None


def frame_against_oracle(tmp_path, maybe_ref, objs, st, seed, tex=(), backtex=None):
    """one full frame of `st` through the text format, drb_render vs the oracle's Kernel(); returns (rmse, identical)"""
    p = str(tmp_path / "scene.rts")
    drb.write_rts(p, st, objs, tex_names=[os.path.basename(t) for t in tex], backtex_name=backtex)
    texdir = str(tmp_path) if tex else ""
    sc = drb.Scene.load(p, texdir or None)
    st = sc.settings                                              # what the text round trip left (6-decimal %f)
    orc = Oracle(p, texdir, maybe_ref)
    orc.apply(st, seed)
    ro, rd = sc.primary_rays(st, sample=0, seed=seed)
    ids, t = sc.trace_ids(ro, rd)
    oid, ot = orc.hit(ro, rd)
    ties = assert_ids_match(ids, t, oid, ot)                      # every primary ray of the frame, bit-exact ids and t
    f, fi, rays = orc.frame()
    acc, stats = sc.render(st, seed=seed)
    assert stats.paths == st.width * st.height * st.spp
    assert stats.rays == rays, (stats.rays, rays)
    ours = acc.transpose(1, 0, 2) * np.float32(255.0) * np.float32(1.0 / st.spp)
    rmse, same = assert_frames_match(ours, f)
    ints = sc.frame_i3(st, 1, seed=seed)
    assert np.mean(ints == fi) >= 0.999
    print("full-size parity: %d x %d, %d spp, depth %d: %d rays, %d id ties, rmse %.3g, %.4f%% pixels bit-identical" %
          (st.width, st.height, st.spp, st.max_depth, rays, ties, rmse, 100 * same))
    return rmse, same


def test_config3_million_triangles_1080p_frame_against_oracle(tmp_path, maybe_ref):
    objs, st = synth.instanced_grid_scene(spp=1, max_depth=10)           # 1 048 580 triangles, 1920x1080, 10 bounces
    frame_against_oracle(tmp_path, maybe_ref, objs, st, seed=5)


def test_config2_bunny_class_1080p_frame_against_oracle(tmp_path, maybe_ref):
    objs, st = synth.bunny_class_scene(spp=2, max_depth=8)               # 245 764 triangles, 1920x1080, 8 bounces
    frame_against_oracle(tmp_path, maybe_ref, objs, st, seed=6)


def test_config4_materials_textures_env_1080p_frame_against_oracle(tmp_path, maybe_ref):
    tex = synth.write_test_textures(str(tmp_path))
    objs, st, tp = synth.materials_scene(tex, spp=2, max_depth=10)       # every material class, textures, env map, 1920x1080
    frame_against_oracle(tmp_path, maybe_ref, objs, st, seed=7, tex=tp, backtex=os.path.basename(tp[0]))


def test_config5_ten_million_triangles_ids_against_brute_force():
    objs, st = synth.city_scene()                                        # 10 000 002 triangles, 3840x2160
    sc = drb.Scene.from_host(drb.HostScene.from_objects(objs, st))
    assert sc.num_prims == len(objs) > 10_000_000 - 8
    bi = sc.build_info
    assert bi.stack_levels <= 3 * bi.wide_levels + 2
    rng = np.random.default_rng(11)
    o, d = sc.primary_rays(st, sample=0, seed=2)
    pick = rng.choice(st.width * st.height, 192, replace=False)
    po, pd = o.reshape(-1, 3)[pick], d.reshape(-1, 3)[pick]
    lo, hi = np.array(list(bi.bounds_min), np.float32), np.array(list(bi.bounds_max), np.float32)
    # incoherent rays: origins inside the scene bounds (a little above the ground plane's side of it), any direction;
    # directions are scaled like bounce rays (unit length) and like camera rays (long), since the reference's epsilons
    # are in ray-parameter units
    io = (lo + rng.uniform(0, 1, (192, 3)).astype(np.float32) * (hi - lo)).astype(np.float32)
    idir = rng.normal(size=(192, 3)).astype(np.float32)
    idir /= np.linalg.norm(idir, axis=1, keepdims=True)
    idir[96:] *= np.float32(300.0)
    allo, alld = np.concatenate([po, io]), np.concatenate([pd, idir])
    ids, t = sc.trace_ids(allo, alld)
    bid, bt = restated.brute_tris(objs["pos"], objs["dim"], objs["rot"], allo, alld)
    assert (bid[:192] >= 0).mean() > 0.5 and (bid[192:] >= 0).mean() > 0.2          # the subsample does exercise the tree
    ties = assert_ids_match(ids, t, bid, bt)
    print("config 5: %d rays against brute force over %d triangles, %d exact-t ties, %d hits" % (len(allo), len(objs), ties, int((bid >= 0).sum())))


@pytest.fixture(scope="module")
def big():
    objs, st = synth.instanced_grid_scene(spp=2, max_depth=6)          # 1 048 580 triangles, 1920x1080
    sc = drb.Scene.from_host(drb.HostScene.from_objects(objs, st))
    return objs, st, sc


def brute_force_ids(objs, o, d):
    """hit_tri (kernel.cu:277-313) over every triangle in float32 numpy, operation for operation"""
    f = np.float32
    v0, e1, e2 = objs["pos"], objs["dim"] - objs["pos"], objs["rot"] - objs["pos"]
    ids = np.full(len(o), -1, np.int32); ts = np.full(len(o), -1, np.float32)

    def cross(a, b):
        return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                         a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1)

    def dot(a, b):
        return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]

    for k in range(len(o)):
        oo, dd = o[k], d[k]
        h = cross(dd[None], e2)
        a = dot(e1, h)
        ok = ~((a > f(-0.0001)) & (a < f(0.0001)))
        with np.errstate(divide="ignore", invalid="ignore"):
            ff = (1.0 / a.astype(np.float64)).astype(np.float32)
        s = oo[None] - v0
        u = ff * dot(s, h)
        ok &= ~((u < 0) | (u > 1))
        q = cross(s, e1)
        v = ff * dot(np.broadcast_to(dd, q.shape), q)
        ok &= ~((v < 0) | (u + v > 1))
        t = ff * dot(e2, q)
        ok &= (t > f(0.0001)) & (t < f(10000))
        if ok.any():
            tt = np.where(ok, t, np.inf)
            ts[k] = tt.min(); ids[k] = int(np.flatnonzero(tt == tt.min())[0])
    return ids, ts


def test_million_triangle_build_and_ids(big):
    objs, st, sc = big
    assert sc.num_prims == 16 * 65536 + 4
    bi = sc.build_info
    assert bi.nnodes == bi.nprims - 1 and 20 <= bi.max_depth <= 96
    o, d = sc.primary_rays(st, sample=0, seed=1)
    ids, t = sc.trace_ids(o, d)
    assert (ids >= 0).mean() > 0.9
    rng = np.random.default_rng(0)
    pick = rng.choice(o.shape[0] * o.shape[1], 96, replace=False)
    bid, bt = brute_force_ids(objs, o.reshape(-1, 3)[pick], d.reshape(-1, 3)[pick])
    got, gt = ids[pick], t[pick]
    for k in np.flatnonzero(got != bid):
        assert gt[k] == bt[k]                                           # exact-t tie only
    assert np.array_equal(gt[got == bid], bt[got == bid])
    assert (got == bid).mean() > 0.95


def test_full_size_render_properties(big):
    objs, st, sc = big
    full, sf = sc.render(st, seed=3)
    assert sf.paths == 1920 * 1080 * 2 and sf.paths < sf.rays <= sf.paths * st.max_depth
    assert np.isfinite(full).all() and full.min() >= 0
    a, sa = sc.render(st, seed=3, sample_base=0, sample_count=1)
    b, sb = sc.render(st, seed=3, sample_base=1, sample_count=1)
    assert sa.rays + sb.rays == sf.rays                                 # disjoint sample shards cover the same paths
    assert np.allclose(a + b, full, rtol=0, atol=1e-5)
    again, _ = sc.render(st, seed=3)
    assert np.array_equal(again, full)
    img = drb.tonemap(full, 2)
    assert img.shape == (1080, 1920, 3) and 20 < img.mean() < 235
