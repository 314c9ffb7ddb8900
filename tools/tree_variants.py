"""Traversal cost of a scene under the build variants: `python tools/tree_variants.py [city10m|grid1m] [spp]`.
Prints k_trace time and rays for the default tree and the Karras-only tree (DRB_BUILD_LBVH_ONLY); compile-time variants
(DRB_PLOC_RADIUS, DRB_SORTED_PUSH, ...) are compared by rebuilding the library with DRB_NVCC_EXTRA and running this again."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dogeray_b200 as drb
from dogeray_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "city10m"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
objs, st = synth.city_scene() if name == "city10m" else synth.instanced_grid_scene()
st = st.replace(spp=spp)
hs = drb.HostScene.from_objects(objs, st)
for label, flags in (("default", 0), ("lbvh_only", drb.BUILD_LBVH_ONLY)):
    sc = drb.Scene.from_host(hs, build_flags=flags)
    bi = sc.build_info
    sc.render(st, seed=0)
    _, s = sc.render(st, seed=0)
    print("%s %s: build %.1f ms, %d wide nodes, %d levels, stack %d | %d rays, k_trace %.1f ms (%.0f Mrays/s), frame %.1f ms" %
          (name, label, bi.build_ms, bi.nwide, bi.wide_levels, bi.stack_levels, s.rays, s.trace_ms, s.rays / s.trace_ms / 1e3, s.total_ms), flush=True)
    sc.close()
