"""Developer check, run on a B200 through gpurun: new CUDA path vs the host-compiled reference.
Not a test (tests/ holds those); prints a summary so a first bring-up can be read from one log."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dogeray_b200 as drb  # noqa: E402
from oracle import refhost  # noqa: E402


def check_scene(ref, name, W=256, H=256, spp=4, depth=4, cam=None):
    path = os.path.join(refhost.SAMPLES, name)
    t0 = time.time()
    sc = drb.Scene.load(path, refhost.SAMPLES)
    bi = sc.build_info
    print("== %s: %d prims, %d nodes, height %d, upload %.2f ms, build %.2f ms, load wall %.3f s" %
          (name, bi.nprims, bi.nnodes, bi.max_depth, bi.upload_ms, bi.build_ms, time.time() - t0))
    ref.load(path, refhost.SAMPLES)
    st = sc.settings.replace(width=W, height=H, spp=spp, max_depth=depth)
    if cam is not None:
        st = st.replace(cam=cam)
    ref.apply(st)
    # ids on identical rays
    o, d = sc.primary_rays(st, sample=0, seed=7)
    ids, t = sc.trace_ids(o, d)
    rid, rt = ref.hit(o, d)
    mism = np.flatnonzero(ids != rid)
    print("   primary ids: %d rays, %d hits, mismatches %d, t bit-equal on agreeing hits: %s" %
          (len(ids), int((rid >= 0).sum()), len(mism), bool(np.array_equal(t[(ids == rid) & (rid >= 0)], rt[(ids == rid) & (rid >= 0)]))))
    for k in mism[:5]:
        print("     ray %d: ours id %d t %r, ref id %d t %r" % (k, ids[k], t[k], rid[k], rt[k]))
    # frame
    ref.set_seed(7)
    rf, ri, rrays = ref.frame(1, 0)
    acc, stats = sc.render(st, seed=7)
    ours = acc.transpose(1, 0, 2) * 255.0 * np.float32(1.0 / spp)    # (W,H,3) like outputr
    diff = np.abs(ours - rf)
    print("   frame %dx%d spp %d depth %d: rays ours %d ref %d; max|diff| %.4g, mean|diff| %.4g (255 scale), pixels>0.5: %d of %d" %
          (W, H, spp, depth, stats.rays, rrays, diff.max(), diff.mean(), int((diff.max(axis=2) > 0.5).sum()), W * H))
    fi = sc.frame_i3(st, 1, seed=7)
    print("   frame_i3 equal to reference ints: %.4f%% of entries" % (100.0 * np.mean(fi == ri)))
    print("   stats:", stats.as_dict())
    sc.close()


def main():
    print("devices:", drb.device_count())
    ref = refhost.RefHost()
    check_scene(ref, "cube.rts", cam=(6.0, -5.0, 9.0))
    check_scene(ref, "cube.rts")
    check_scene(ref, "mats.rts")
    check_scene(ref, "glass.rts", depth=8)
    check_scene(ref, "bolter2.blend.rts", W=128, H=128, spp=2, depth=6)
    check_scene(ref, "rough.blend.rts", W=128, H=128, spp=2, depth=6)
    check_scene(ref, "gloss.rts", W=128, H=128, spp=2, depth=6)
    check_scene(ref, "uv2.rts", W=128, H=128, spp=2, depth=6)
    check_scene(ref, "lots.rts", W=128, H=128, spp=2, depth=6)
    check_scene(ref, "SPERSSSSS.rts", W=128, H=128, spp=2, depth=6)


if __name__ == "__main__":
    main()
