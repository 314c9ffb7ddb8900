"""Scene creation timing (upload + GPU tree build), repeated: `python tools/build_timing.py [grid1m|city10m] [n]`.
Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel list of one build."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dogeray_b200 as drb
from dogeray_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "grid1m"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
objs, st = synth.city_scene() if name == "city10m" else synth.instanced_grid_scene()
hs = drb.HostScene.from_objects(objs, st)
hs.num_renderable
for k in range(n):
    t0 = time.perf_counter()
    sc = drb.Scene.from_host(hs)
    t1 = time.perf_counter()
    bi = sc.build_info
    print("create %d: wall %.2f ms, upload %.2f ms, build %.2f ms (%d rounds, %d wide levels, %d wide nodes, stack %d)" %
          (k, (t1 - t0) * 1e3, bi.upload_ms, bi.build_ms, bi.rebuild_iterations, bi.wide_levels, bi.nwide, bi.stack_levels), flush=True)
    sc.close()
