"""The in-process multi-GPU path (drb_scene_create_multi + drb_render_multi) on a box with several GPUs:

    python tools/cli_multi_check.py [spp] [grid1m|city10m|mats]

  * the CLI at --gpus 1/2/4/8 (static tiles, --dynamic, --shard samples): byte-identical BMPs for the tile modes, wall
    clock and the device time the CLI prints;
  * per-handle device times of one static-tile frame through the library (drb_render_multi_times): the static imbalance
    max/mean - 1, i.e. what dynamic balancing could win on equal GPUs."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dogeray_b200 as drb
from dogeray_b200 import synth

spp = sys.argv[1] if len(sys.argv) > 1 else "64"
which = sys.argv[2] if len(sys.argv) > 2 else "grid1m"
cli = os.path.join(os.path.dirname(drb.__file__), "dogeray-b200")
d = tempfile.mkdtemp(prefix="drb_cli_")
tex = []
if which == "city10m":
    objs, st = synth.city_scene()
elif which == "mats":                                    # BASELINE config 4 stand-in: every material class, textures, environment map
    objs, st, tex = synth.materials_scene(synth.write_test_textures(d))
else:
    objs, st = synth.instanced_grid_scene()
res = "%dx%d" % (st.width, st.height)
ndev = drb.device_count()
out, rows = {}, []
use_cli = which == "grid1m"             # the 10 M-triangle scene is 3.3 GB of text: it and the material scene go through the library only
if use_cli:
    drb.write_rts(os.path.join(d, "scene.rts"), st, objs)
for n in (1, 2, 4, 8):
    if n > ndev or not use_cli:
        continue
    for mode, extra in (("tiles", []), ("dynamic", ["--dynamic"]), ("samples", ["--shard", "samples"])):
        if n == 1 and mode != "tiles":
            continue
        name = "g%d_%s.bmp" % (n, mode)
        t = time.time()
        p = subprocess.run([cli, "scene.rts", "--spp", spp, "--gpus", str(n), "--cache", "--out", name] + extra, capture_output=True, text=True, cwd=d)
        wall = time.time() - t
        assert p.returncode == 0, p.stdout + p.stderr
        out[(n, mode)] = open(os.path.join(d, name), "rb").read()
        line = [l for l in p.stdout.splitlines() if l.startswith("Time")][0]
        ms = float(line.split()[2])
        mrays = float(line.split()[-2])
        rows.append({"scene": which, "res": res, "spp": int(spp), "gpus": n, "mode": mode, "device_ms": ms, "mrays_s": mrays, "wall_s": round(wall, 2),
                     "identical_to_1gpu": out[(n, mode)] == out[(1, "tiles")]})
        print(json.dumps(rows[-1]), flush=True)
        if mode != "samples":
            assert out[(n, mode)] == out[(1, "tiles")], "image differs from the 1-GPU image"
# static imbalance through the library
hs = drb.HostScene.load(os.path.join(d, "scene.rts"), d, cache=True) if use_cli else drb.HostScene.from_objects(objs, st, tex)
one = None
for n in (1, 2, 4, 8):
    if n > ndev:
        continue
    t0 = time.time()
    scenes = drb.create_multi(hs, list(range(n)))
    t_create = time.time() - t0
    s1 = st.replace(spp=int(spp))
    drb.render_multi(scenes, s1, seed=0)                             # warm-up (buffers)
    t0 = time.time()
    img, sm = drb.render_multi(scenes, s1, seed=0)
    wall = time.time() - t0
    times = drb.render_multi_times()
    if n == 1:
        one = img
    imgd, sd = drb.render_multi(scenes, s1, seed=0, dynamic=True)
    dtimes = drb.render_multi_times()
    print(json.dumps({"scene": which, "res": res, "spp": int(spp), "gpus": n, "create_multi_s": round(t_create, 3), "render_multi_wall_ms": round(wall * 1e3, 2),
                      "mrays_s_wall": round(sm.rays / wall / 1e6, 1), "static_ms_per_handle": [round(t, 2) for t in times],
                      "static_imbalance": max(times) / (sum(times) / n) - 1, "static_frame_ms": max(times),
                      "dynamic_busy_ms_per_handle": [round(t, 2) for t in dtimes], "dynamic_frame_ms": max(dtimes),
                      "bit_identical_to_1gpu": bool(np.array_equal(img, one)) and bool(np.array_equal(imgd, one))}), flush=True)
    for sc in scenes:
        sc.close()
print("done")
