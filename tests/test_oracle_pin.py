"""Pins the plain-C restatement (oracle/dogeray_oracle.c) to the reference: against oracle/_ref (the
reference's own kernel.cu compiled for the host) where it is present, and against the committed golden
vectors that were generated from oracle/_ref (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import dogeray_b200 as drb
from dogeray_b200 import synth
from oracle import restated
from conftest import SAMPLES, needs_ref, sample

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PIN_SCENES = [("cube.rts", {}), ("mats.rts", {}), ("glass.rts", {"max_depth": 8}), ("bolter2.blend.rts", {}), ("rough.blend.rts", {}),
              ("gloss.rts", {}), ("uv2.rts", {}), ("lots.rts", {}), ("cow.rts", {}), ("smoothdiff.rts", {}), ("glasstest.rts", {"max_depth": 8}),
              ("SPERSSSSS.rts", {})]


def settings_from_vector(v):
    st = drb.default_settings()
    st.cam[:] = v[0:3]; st.aperture = v[3]; st.look[:] = v[4:7]; st.focus = v[7]
    st.fov, st.max_depth, st.spp = int(v[8]), int(v[9]), int(v[10])
    st.bg_intensity = v[11]; st.backtex = int(v[12]); st.width, st.height = int(v[13]), int(v[14])
    return st


@needs_ref
@pytest.mark.parametrize("name,over", PIN_SCENES)
def test_restatement_equals_reference(ref, name, over):
    ref.load(sample(name), SAMPLES)
    r = restated.Restated(sample(name), SAMPLES)
    assert r.num_objects == ref.num_objects and r.num_nodes == ref.num_nodes
    s = ref.get_settings()
    assert np.array_equal(s, r.get_settings())
    s[13], s[14], s[10], s[9] = 48, 40, 2, over.get("max_depth", 5)
    ref.set_settings(s); r.set_settings(s); ref.set_seed(9); r.set_seed(9)
    f1, i1, rays1 = ref.frame()
    f2, i2, rays2 = r.frame()
    assert rays1 == rays2
    assert np.array_equal(f1, f2) and np.array_equal(i1, i2)            # bit-identical float frames
    o, d = r.primary_rays(0)
    ids1, t1 = ref.hit(o, d)
    ids2, t2 = r.hit(o, d)
    assert np.array_equal(ids1, ids2) and np.array_equal(t1, t2)
    ids3, t3 = r.hit_brute(o, d)                                        # the tree never changes the answer
    same = ids3 == ids1
    assert np.array_equal(t3[same], t1[same])
    for k in np.flatnonzero(~same):                                     # only exact-t ties may pick another object
        assert t3[k] == t1[k]


@needs_ref
def test_frame_divisor_and_sample_base(ref):
    ref.load(sample("mats.rts"), SAMPLES)
    r = restated.Restated(sample("mats.rts"), SAMPLES)
    s = ref.get_settings(); s[13], s[14], s[10], s[9] = 72, 50, 2, 4        # 50 is not a multiple of 8: partial blocks dropped
    for div, base in ((1, 0), (2, 0), (1, 7)):
        ref.set_settings(s); r.set_settings(s); ref.set_seed(1); r.set_seed(1)
        f1, i1, _ = ref.frame(div, base)
        f2, i2, _ = r.frame(div, base)
        assert np.array_equal(f1, f2) and np.array_equal(i1, i2)
        assert (i1[72 // div // 8 * 8:, :, :] == 0).all()


def _check_golden(g, r):
    st = settings_from_vector(g["settings"])
    r.apply(st); r.set_seed(int(g["seed"]))
    o, d = r.primary_rays(0)
    assert np.array_equal(o, g["origins"]) and np.array_equal(d, g["dirs"])
    ids, t = r.hit(o, d)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(t, g["t"])
    f, i, rays = r.frame()
    assert rays == int(g["rays"]) and np.array_equal(f, g["frame"]) and np.array_equal(i, g["frame_i"])


def test_restatement_against_golden_synthetic(tmp_path):
    """runs everywhere: the scene is regenerated, the expected numbers came from the reference"""
    g = np.load(os.path.join(GOLDEN, "synth_heightfield.npz"))
    objs, st = synth.heightfield_scene(n=24, width=48, height=40, spp=3, max_depth=5)
    p = str(tmp_path / "hf.rts")
    drb.write_rts(p, st, objs)
    _check_golden(g, restated.Restated(p))


def test_restatement_against_golden_materials(tmp_path):
    """every material class, colour + roughness textures, checker, smooth normals, environment map -- expected values
    from the reference, scene and textures regenerated here"""
    g = np.load(os.path.join(GOLDEN, "synth_materials.npz"))
    tex = synth.write_test_textures(str(tmp_path))
    objs, st, tp = synth.materials_scene(tex, width=96, height=56, spp=3, max_depth=6, nu=24, nv=12)
    p = str(tmp_path / "mats.rts")
    drb.write_rts(p, st, objs, tex_names=[os.path.basename(t) for t in tp], backtex_name=os.path.basename(tp[0]))
    assert int(g["settings"][12]) == 0                                  # the environment map is texture 0
    _check_golden(g, restated.Restated(p, str(tmp_path)))


@needs_ref
def test_restatement_against_golden_cube():
    g = np.load(os.path.join(GOLDEN, "cube_frame.npz"))
    _check_golden(g, restated.Restated(sample("cube.rts")))


def test_brute_force_over_arrays_equals_the_scene_brute_force(tmp_path):
    """orc_brute_tris (the 10 M-triangle check's definition of a closest hit) against orc_hit_brute and the tree walk"""
    objs, st = synth.heightfield_scene(n=20, width=32, height=24, spp=1, max_depth=2)
    p = str(tmp_path / "hf.rts")
    drb.write_rts(p, st, objs)
    objs = drb.HostScene.load(p).objects()                               # the floats the text holds
    r = restated.Restated(p)
    r.apply(st); r.set_seed(1)
    o, d = r.primary_rays(0)
    rng = np.random.default_rng(2)
    o2 = rng.uniform(-5, 5, (500, 3)).astype(np.float32); d2 = rng.normal(size=(500, 3)).astype(np.float32)
    allo = np.concatenate([o.reshape(-1, 3), o2]); alld = np.concatenate([d.reshape(-1, 3), d2])
    a_id, a_t = restated.brute_tris(objs["pos"], objs["dim"], objs["rot"], allo, alld, threads=3)
    b_id, b_t = r.hit_brute(allo, alld)
    c_id, c_t = r.hit(allo, alld)
    assert (a_id >= 0).sum() > 300
    assert np.array_equal(a_id, b_id) and np.array_equal(a_t[a_id >= 0], b_t[a_id >= 0])
    assert np.array_equal(a_id, c_id) and np.array_equal(a_t[a_id >= 0], c_t[a_id >= 0])


def test_host_lbvh_is_a_valid_tree():
    rng = np.random.default_rng(4)
    for n in (1, 2, 3, 17, 1000):
        c = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
        if n == 17:
            c[:] = c[0]                                                  # all keys tie: falls back to positions
        e = rng.uniform(0.01, 0.3, (n, 3)).astype(np.float32)
        t = restated.lbvh_host(c - e, c + e)
        assert sorted(t["order"].tolist()) == list(range(n))
        assert (np.diff(t["keys"].astype(np.int64)) >= 0).all()
        if n == 1:
            continue
        seen_leaf = np.zeros(n, int); seen_node = np.zeros(n - 1, int)
        stack = [0]
        while stack:
            k = stack.pop()
            seen_node[k] += 1
            for ch in (t["left"][k], t["right"][k]):
                if ch < 0:
                    seen_leaf[~ch] += 1
                    lo, hi = (c - e)[t["order"][~ch]], (c + e)[t["order"][~ch]]
                else:
                    assert t["parent"][ch] == k
                    stack.append(int(ch))
                    lo, hi = t["node_min"][ch], t["node_max"][ch]
                assert (t["node_min"][k] <= lo).all() and (t["node_max"][k] >= hi).all()
        assert (seen_leaf == 1).all() and (seen_node == 1).all() and t["parent"][0] == -1
        assert np.array_equal(t["node_min"][0], (c - e).min(0)) and np.array_equal(t["node_max"][0], (c + e).max(0))
        # the SAH-guided rebuild over the same sorted leaves is a valid tree too, with root 0 and parents first
        lmin, lmax = (c - e)[t["order"]], (c + e)[t["order"]]
        p = restated.ploc_host(lmin, lmax)
        seen = np.zeros(n, int)
        stack = [0]
        while stack:
            k = stack.pop()
            for ch in (p["left"][k], p["right"][k]):
                if ch < 0:
                    seen[~ch] += 1
                    assert (p["node_min"][k] <= lmin[~ch]).all() and (p["node_max"][k] >= lmax[~ch]).all()
                else:
                    assert ch > k
                    stack.append(int(ch))
        assert (seen == 1).all() and np.array_equal(p["node_min"][0], (c - e).min(0))
        # four-wide collapse: every leaf once, quantised child boxes contain the float boxes they came from
        sb = np.concatenate([(c - e).min(0), (c + e).max(0)])
        w = restated.wide_host(p["left"], p["right"], p["node_min"], p["node_max"], lmin, lmax, sb)
        qs = np.where(sb[3:] - sb[:3] > 0, sb[3:] - sb[:3], 1).astype(np.float32) / np.float32(65527.0)
        qlo = sb[:3] - np.float32(4.0) * qs
        seen = np.zeros(n, int)
        for i in range(len(w["child"])):
            for k in range(4):
                ch = w["child"][i, k]
                if ch == -2 ** 31:
                    assert (w["boxes"][i, k] == 0xFFFF).all()
                    continue
                lo_q = (w["boxes"][i, k] & 0xFFFF).astype(np.float64); hi_q = (w["boxes"][i, k] >> 16).astype(np.float64)
                if ch < 0:
                    seen[~ch] += 1
                    assert (qlo + lo_q * qs <= lmin[~ch]).all() and (qlo + hi_q * qs >= lmax[~ch]).all()
                else:
                    assert i < ch < len(w["child"])
        assert (seen == 1).all()


def _sah(node_min, node_max):
    """surface-area-heuristic cost of the inner nodes, relative to the root (one-primitive leaves cost the same in
    every tree over the same primitives, so they are left out)"""
    e = (node_max - node_min).astype(np.float64)
    a = e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]
    return float(a.sum() / a[0])


def test_sah_guided_rebuild_beats_the_morton_tree():
    """regression guard on tree quality (the GPU build is bit-identical to these mirrors, tests/test_gpu_parity.py):
    on a mesh-like scene the rebuilt tree has a clearly lower SAH cost than the Karras tree over the same order"""
    objs, st = synth.heightfield_scene(n=60)                               # 7 200 triangles
    tri = objs[objs["type"] == 2]
    v = np.stack([tri["pos"], tri["dim"], tri["rot"]], 1).astype(np.float32)
    bmin, bmax = v.min(1), v.max(1)
    t = restated.lbvh_host(bmin, bmax)
    lmin, lmax = bmin[t["order"]], bmax[t["order"]]
    morton = _sah(t["node_min"], t["node_max"])
    costs = {r: _sah(*(lambda p: (p["node_min"], p["node_max"]))(restated.ploc_host(lmin, lmax, radius=r))) for r in (1, 4, 16, 64)}
    assert costs[16] < 0.80 * morton, (morton, costs)
    assert costs[16] <= costs[4] <= costs[1] * 1.001, costs               # a wider search never hurts on this scene
    assert costs[64] > 0.97 * costs[16], costs                            # ... and radius 16 already has nearly all of it
