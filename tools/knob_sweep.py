"""Sweep of the k_trace scheduling knobs (DOGERAY_B200_REFILL / LEAF_BATCH / STEP_MIN, read when the library is
loaded, hence one process per setting) on the bench frame.  No torch: host-buffer API, device times from drb_stats.
    python tools/knob_sweep.py            # on a GPU box"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

if len(sys.argv) > 2 and sys.argv[1] == "--child":
    import dogeray_b200 as drb
    objs = np.load(sys.argv[2])
    st = drb.Settings.from_buffer_copy(open(sys.argv[2] + ".settings", "rb").read())
    sc = drb.Scene.from_host(drb.HostScene.from_objects(objs, st))
    best = None
    for _ in range(2):
        _, s = sc.render(st, seed=0)
        best = s if best is None or s.total_ms < best.total_ms else best
    print("%.1f %.1f" % (best.total_ms, best.trace_ms))
    sys.exit(0)

from dogeray_b200 import synth
objs, st = synth.instanced_grid_scene()
tmp = tempfile.mkdtemp(prefix="drb_knobs_")
path = os.path.join(tmp, "objs.npy")
np.save(path, objs)
open(path + ".settings", "wb").write(bytes(st))
base = {"REFILL": 24, "LEAF_BATCH": 12, "STEP_MIN": 20}
runs = [dict(base)]
for k, vals in (("REFILL", (16, 20, 28, 32)), ("LEAF_BATCH", (6, 8, 16, 20)), ("STEP_MIN", (12, 16, 24, 28))):
    for v in vals:
        d = dict(base); d[k] = v; runs.append(d)
for d in runs:
    env = dict(os.environ)
    for k, v in d.items():
        env["DOGERAY_B200_" + k] = str(v)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path], capture_output=True, text=True, env=env, timeout=120)
    print("refill %2d leaf_batch %2d step_min %2d : total/trace ms = %s" % (d["REFILL"], d["LEAF_BATCH"], d["STEP_MIN"], r.stdout.strip() or r.stderr[-200:]), flush=True)
os.remove(path); os.remove(path + ".settings"); os.rmdir(tmp)
