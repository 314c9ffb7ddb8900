"""Half the tiles against half the samples of the 1 M-triangle frame on one GPU (what a rank of a 2-GPU run
traces under --shard tiles resp. --shard samples).  python tools/tile_timing.py"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import dogeray_b200 as drb
from dogeray_b200 import synth
objs, st = synth.instanced_grid_scene()
sc = drb.Scene.from_host(drb.HostScene.from_objects(objs, st))
acc = torch.zeros(st.height, st.width, 3, device='cuda')
stream = torch.cuda.current_stream().cuda_stream
def run(**kw):
    sc.render_device(acc.data_ptr(), st, seed=0, stream=stream, want_stats=True, **kw)
    s = sc.render_device(acc.data_ptr(), st, seed=0, stream=stream, want_stats=True, **kw)
    return s
for kw in (dict(sample_count=128), dict(tile_rank=0, tile_count=2), dict(tile_rank=1, tile_count=2), dict(sample_count=256)):
    s = run(**kw)
    print(kw, "total %.1f ms trace %.1f ms rays %d paths %d launches %d" % (s.total_ms, s.trace_ms, s.rays, s.paths, s.kernel_launches))
