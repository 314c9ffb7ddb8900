"""On a box with >= 2 GPUs: the CLI's --gpus 2 against --gpus 1 on the 1 M-triangle scene (byte-identical BMP,
wall-clock and device times as the CLI prints them).  python tools/cli_multi_check.py [spp]"""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dogeray_b200 as drb
from dogeray_b200 import synth

spp = sys.argv[1] if len(sys.argv) > 1 else "64"
cli = os.path.join(os.path.dirname(drb.__file__), "dogeray-b200")
d = tempfile.mkdtemp(prefix="drb_cli_")
objs, st = synth.instanced_grid_scene()
drb.write_rts(os.path.join(d, "scene.rts"), st, objs)
out = {}
for n in range(1, drb.device_count() + 1):
    if n not in (1, 2, 4, 8):
        continue
    t = time.time()
    p = subprocess.run([cli, "scene.rts", "--spp", spp, "--gpus", str(n), "--cache", "--out", "g%d.bmp" % n], capture_output=True, text=True, cwd=d)
    wall = time.time() - t
    assert p.returncode == 0, p.stdout + p.stderr
    out[n] = open(os.path.join(d, "g%d.bmp" % n), "rb").read()
    print("gpus=%d wall %.2f s | %s" % (n, wall, " | ".join(l for l in p.stdout.splitlines() if l.startswith(("Time", "scene cache", "Done")))), flush=True)
    assert out[n] == out[1], "image differs from the 1-GPU image"
print("all images byte-identical:", sorted(out))
