"""A small frame through every kernel (build, trace, shade with all material classes, resolve, multi-handle render), for
`compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py` (one tool per run)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dogeray_b200 as drb
from dogeray_b200 import synth

d = tempfile.mkdtemp(prefix="drb_san_")
tex = synth.write_test_textures(d)
objs, st, tp = synth.materials_scene(tex, width=64, height=40, spp=3, max_depth=5, nu=12, nv=6)
hs = drb.HostScene.from_objects(objs, st, tp)
sc = drb.Scene.from_host(hs)
img, s = sc.render(st, seed=1)
print("rendered", s.rays, "rays", float(img.mean()))
o, dd = sc.primary_rays(st, 0, seed=1)
ids, t = sc.trace_ids(o, dd)
print("ids", int((ids >= 0).sum()))
pair = drb.create_multi(hs, [0, 0])
img2, s2 = drb.render_multi(pair, st, seed=1, dynamic=True)
assert np.array_equal(img, img2)
img3, _ = drb.render_multi(pair, st, seed=1, shard="samples")
print("multi ok", float(np.abs(img3 - img).max()))
