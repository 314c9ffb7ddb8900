"""Developer diagnostic (test infrastructure: it uses the oracle as the checker): the randomised scenes of
tests/test_gpu_parity_wide.py one by one, printing where ids / distances / radiance differ from the host-compiled
reference.  Run on a GPU box: python tests/dev/fuzz_diag.py"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dogeray_b200 as drb
from oracle import refhost, restated
from test_gpu_parity_wide import random_scene

seed = int(sys.argv[1])
rng = np.random.default_rng(1000 + seed)
n = int(rng.choice([3, 17, 64, 300, 1500])); scale = float(rng.choice([0.5, 3.0, 40.0]))
objs = random_scene(rng, n, scale)
st = drb.default_settings().replace(cam=(0.3 * scale, -0.2 * scale, 2.8 * scale), look=(0, 0, 0), width=72, height=48, spp=2, max_depth=6,
                                    focus=3.0, aperture=float(rng.choice([0.0, 0.01, 0.2])), fov=int(rng.choice([30, 45, 70])))
print("n", n, "scale", scale, "aperture", st.aperture, "fov", st.fov)
td = tempfile.mkdtemp(); p = os.path.join(td, "r.rts"); drb.write_rts(p, st, objs)
sc = drb.Scene.load(p); ref = refhost.RefHost(); ref.load(p, ""); ref.apply(st); ref.set_seed(seed)
o = drb.HostScene.load(p).objects()
for depth in (1, 2, 6):
    s2 = st.replace(max_depth=depth); ref.apply(s2); ref.set_seed(seed)
    f, fi, rays = ref.frame(); acc, stats = sc.render(s2, seed=seed)
    ours = acc.transpose(1, 0, 2) * np.float32(255.0) * np.float32(0.5)
    bad = ~np.all(ours == f, axis=-1)
    print("depth", depth, "identical", 1 - bad.mean(), "rays", stats.rays, rays, "maxdiff", np.abs(ours - f).max())
ro, rd = sc.primary_rays(st, 0, seed=seed); ids, t = sc.trace_ids(ro, rd); ids = ids.reshape(48, 72).T
s2 = st.replace(max_depth=1, spp=1); ref.apply(s2); ref.set_seed(seed)
f, fi, rays = ref.frame(); acc, stats = sc.render(s2, seed=seed)
ours = acc.transpose(1, 0, 2) * np.float32(255.0)
bad = ~np.all(ours == f, axis=-1)
print("depth1 spp1 bad", bad.sum())
hit = ids[bad]
import collections
print("types at bad pixels", collections.Counter(o["type"][hit[hit >= 0]].tolist()), "mats", collections.Counter(o["mat"][hit[hit >= 0]].tolist()), "miss", int((hit < 0).sum()))
print("types overall", collections.Counter(o["type"][ids[ids >= 0]].tolist()), "mats", collections.Counter(o["mat"][ids[ids >= 0]].tolist()))
k = np.argwhere(bad)[:5]
for x, y in k:
    print((x, y), "id", ids[x, y], "type", o["type"][ids[x, y]] if ids[x, y] >= 0 else None, "mat", o["mat"][ids[x, y]] if ids[x, y] >= 0 else None, "ours", ours[x, y], "ref", f[x, y])
