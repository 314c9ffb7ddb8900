"""Regenerate the golden vectors from the REFERENCE ITSELF (oracle/_ref, built by oracle/make_ref.py from
/root/reference/raygpu/kernel.cu).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

Vectors (all small, committed):
  synth_heightfield.npz  a 2*24*24-triangle synthetic scene (dogeray_b200.synth.heightfield_scene, written with
                         the product's .rts writer, loaded by the reference's read()): primary rays, closest-hit
                         ids / t from the reference's hit(), and a 48x40 float frame from the reference's Kernel()
  cube_frame.npz         samples/cube.rts, camera (6,-5,9), 40x32, 3 spp, depth 4, seed 11: ids / t / float frame
  synth_materials.npz    dogeray_b200.synth.materials_scene at 24x12 per blob (every material class of the reference,
                         colour + roughness textures, checker, smooth normals, environment map; textures from
                         synth.write_test_textures), 96x56, 3 spp, depth 6, seed 13: ids / t / float frame
  bmp_headers.json       the file + BITMAPV4 headers (first 122 bytes, hex) of every screenshot the reference ships in
                         images/*.bmp (SDL_SaveBMP output, kernel.cu:2505-2513) with their sizes: the only golden
                         artefacts the reference holds for the image writer
  philox_kat.json        Random123 known-answer vectors for Philox4x32-10 (published with the algorithm)
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dogeray_b200 as drb  # noqa: E402
from dogeray_b200 import synth  # noqa: E402
from oracle import refhost, restated  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def bmp_headers(out):
    import glob
    import struct
    rec = {}
    for p in sorted(glob.glob("/root/reference/images/*.bmp")):
        with open(p, "rb") as f:
            b = f.read()
        off = struct.unpack_from("<I", b, 10)[0]
        w, h = struct.unpack_from("<ii", b, 18)
        rec[os.path.basename(p)] = {"file_size": len(b), "width": w, "height": h, "pixel_offset": off, "header_hex": b[:off].hex(),
                                    "alpha_values": sorted(set(b[off + 3::4]))}
    with open(out, "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print("wrote", out, len(rec), "headers")


def golden_for(ref, rts, texdir, st, seed, out):
    """ids / frame straight from the reference; rays from the restated camera (pinned to the reference's by the frame)."""
    ref.load(rts, texdir)
    ref.apply(st)
    ref.set_seed(seed)
    r = restated.Restated(rts, texdir)
    r.apply(st); r.set_seed(seed)
    o, d = r.primary_rays(0)
    ids, t = ref.hit(o, d)
    f, i, rays = ref.frame(1, 0)
    np.savez_compressed(out, origins=o, dirs=d, ids=ids, t=t, frame=f, frame_i=i, rays=np.int64(rays),
                        settings=np.array([st.cam[0], st.cam[1], st.cam[2], st.aperture, st.look[0], st.look[1], st.look[2], st.focus,
                                           st.fov, st.max_depth, st.spp, st.bg_intensity, st.backtex, st.width, st.height], np.float64),
                        seed=np.int64(seed))
    print(out, "hits", int((ids >= 0).sum()), "of", len(ids), "rays", rays)


def main():
    bmp_headers(os.path.join(HERE, "bmp_headers.json"))
    if "--bmp-only" in sys.argv:
        return
    ref = refhost.RefHost()
    objs, st = synth.heightfield_scene(n=24, width=48, height=40, spp=3, max_depth=5)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "hf.rts")
        drb.write_rts(p, st, objs)
        golden_for(ref, p, "", st, 5, os.path.join(HERE, "synth_heightfield.npz"))
    st = drb.HostScene.load(os.path.join(refhost.SAMPLES, "cube.rts")).settings.replace(cam=(6.0, -5.0, 9.0), width=40, height=32, spp=3, max_depth=4)
    golden_for(ref, os.path.join(refhost.SAMPLES, "cube.rts"), "", st, 11, os.path.join(HERE, "cube_frame.npz"))
    with tempfile.TemporaryDirectory() as td:
        tex = synth.write_test_textures(td)
        objs, st, tp = synth.materials_scene(tex, width=96, height=56, spp=3, max_depth=6, nu=24, nv=12)
        p = os.path.join(td, "mats.rts")
        drb.write_rts(p, st, objs, tex_names=[os.path.basename(t) for t in tp], backtex_name=os.path.basename(tp[0]))
        st = drb.HostScene.load(p, td).settings                      # backtex resolved against the directory
        golden_for(ref, p, td, st, 13, os.path.join(HERE, "synth_materials.npz"))
    kat = [
        {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
        {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
        {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
         "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
    ]
    with open(os.path.join(HERE, "philox_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)


if __name__ == "__main__":
    main()
