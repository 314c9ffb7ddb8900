"""Step-to-step spread of the resident frame time (1 M triangles, 1080p, 256 spp): several renders per scene
handle, several handles, with and without handing the cached device blocks back in between."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dogeray_b200 as drb
from dogeray_b200 import synth

objs, st = synth.instanced_grid_scene()
hs = drb.HostScene.from_objects(objs, st)
import torch
acc = torch.zeros(st.height, st.width, 3, device="cuda")
for rep in range(4):
    sc = drb.Scene.from_host(hs)
    line = []
    for k in range(4):
        s = sc.render_device(acc.data_ptr(), st, seed=0, want_stats=True)
        line.append("%.1f/%.1f" % (s.total_ms, s.trace_ms))
    print("handle %d%s: total/trace ms = %s" % (rep, " (after trim)" if rep >= 2 else "", "  ".join(line)), flush=True)
    sc.close()
    if rep >= 1:
        drb.trim(0)
