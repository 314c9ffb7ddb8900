"""The headless CLI (drop-in for `raygpu.exe scene.rts`, kernel.cu:2041-2051, 2501-2516)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import dogeray_b200 as drb
from dogeray_b200 import synth
from conftest import ROOT

CLI = os.path.join(ROOT, "dogeray_b200", "dogeray-b200")


def test_cli_help_and_bad_option():
    p = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and "usage: dogeray-b200" in p.stdout
    p = subprocess.run([CLI, "--bogus"], capture_output=True, text=True)
    assert p.returncode == 2


def test_cli_missing_scene_fails_loudly(tmp_path):
    p = subprocess.run([CLI], capture_output=True, text=True, cwd=tmp_path)       # default scene.rts does not exist here
    assert p.returncode == 1 and "cannot open scene file" in p.stderr


def test_cli_without_gpu_has_no_fallback(tmp_path):
    if drb.device_count() > 0:
        pytest.skip("a GPU is present")
    objs, st = synth.heightfield_scene(n=4, width=16, height=8, spp=1)
    drb.write_rts(str(tmp_path / "scene.rts"), st, objs)
    p = subprocess.run([CLI], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 1 and "no CPU path" in p.stderr
    assert not os.path.exists(tmp_path / "scene.rts.bmp")


def test_cli_cache_flag_writes_then_hits(tmp_path):
    """--cache goes through drb_host_scene_load_cached before any CUDA call, so it is observable without a GPU"""
    objs, st = synth.heightfield_scene(n=4, width=16, height=8, spp=1)
    drb.write_rts(str(tmp_path / "scene.rts"), st, objs)
    p = subprocess.run([CLI, "--cache"], capture_output=True, text=True, cwd=tmp_path)
    assert "scene cache: miss (written)" in p.stdout and os.path.exists(tmp_path / "scene.rts.drbcache")
    p = subprocess.run([CLI, "--cache"], capture_output=True, text=True, cwd=tmp_path)
    assert "scene cache: hit" in p.stdout and "%d tris" % len(objs) in p.stdout
    assert p.returncode == (0 if drb.device_count() > 0 else 1)


@pytest.mark.gpu
def test_cli_renders_the_same_image_as_the_library(tmp_path):
    objs, st = synth.heightfield_scene(n=16, width=64, height=40, spp=3, max_depth=4)
    rng = np.random.default_rng(0)
    tex = rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)
    with open(tmp_path / "sky.ppm", "wb") as f:                                      # found through the CWD scan
        f.write(b"P6\n8 8\n255\n" + tex.tobytes())
    drb.write_rts(str(tmp_path / "scene.rts"), st, objs, backtex_name="sky.ppm")
    p = subprocess.run([CLI, "--seed", "9"], capture_output=True, text=True, cwd=tmp_path)     # no path: opens scene.rts
    assert p.returncode == 0, p.stdout + p.stderr
    assert "exported image:scene.rts.bmp" in p.stdout and "1 textures total" in p.stdout
    raw = open(tmp_path / "scene.rts.bmp", "rb").read()
    px = np.frombuffer(raw[122:], np.uint8).reshape(40, 64, 4)[::-1, :, 2::-1]       # bottom-up BGRA -> top-down RGB
    sc = drb.Scene.load(str(tmp_path / "scene.rts"), str(tmp_path))
    assert sc.settings.backtex == 0
    acc, _ = sc.render(sc.settings, seed=9)
    assert np.array_equal(px, drb.tonemap(acc, 3))
    p = subprocess.run([CLI, "scene.rts", "--spp", "1", "--res", "32x16", "--out", "o.ppm"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 0 and open(tmp_path / "o.ppm", "rb").read().startswith(b"P6\n32 16\n255\n")


@pytest.mark.gpu
def test_cli_progressive_snapshots_and_resume(tmp_path):
    """chunked accumulation + checkpoint/resume reproduce the one-shot image (disjoint Philox sample indices)"""
    objs, st = synth.heightfield_scene(n=12, width=48, height=32, spp=6, max_depth=4)
    drb.write_rts(str(tmp_path / "s.rts"), st, objs)
    run = lambda *a: subprocess.run([CLI, "s.rts", "--seed", "4"] + list(a), capture_output=True, text=True, cwd=tmp_path)
    p = run("--out", "full.ppm"); assert p.returncode == 0, p.stderr
    p = run("--out", "prog.ppm", "--snapshot-every", "2"); assert p.returncode == 0 and p.stdout.count("samples -> prog.ppm") == 3
    p = run("--spp", "4", "--out", "a.ppm", "--save-acc", "ck.acc"); assert p.returncode == 0
    p = run("--spp", "2", "--out", "b.ppm", "--resume", "ck.acc"); assert p.returncode == 0 and "resumed 4 samples" in p.stdout
    full = np.frombuffer(open(tmp_path / "full.ppm", "rb").read()[len(b"P6\n48 32\n255\n"):], np.uint8).astype(int)
    for name in ("prog.ppm", "b.ppm"):
        img = np.frombuffer(open(tmp_path / name, "rb").read()[len(b"P6\n48 32\n255\n"):], np.uint8).astype(int)
        assert np.array_equal(img, full)                          # chunks and resumes continue one running sum per pixel


@pytest.mark.gpu
def test_cli_device_list_renders_the_one_gpu_image(tmp_path):
    """--devices 0,0,0: three device scenes, three host threads, interleaved tiles -> byte-identical BMP"""
    objs, st = synth.heightfield_scene(n=12, width=52, height=30, spp=4, max_depth=4)
    drb.write_rts(str(tmp_path / "s.rts"), st, objs)
    run = lambda *a: subprocess.run([CLI, "s.rts", "--seed", "6"] + list(a), capture_output=True, text=True, cwd=tmp_path)
    p = run("--out", "one.bmp"); assert p.returncode == 0, p.stdout + p.stderr
    p = run("--devices", "0,0,0", "--out", "three.bmp"); assert p.returncode == 0, p.stdout + p.stderr
    assert "3 device scenes" in p.stdout
    assert open(tmp_path / "one.bmp", "rb").read() == open(tmp_path / "three.bmp", "rb").read()
    # progressive chunks, one handle or two: the float accumulators themselves are identical
    p = run("--gpus", "1", "--snapshot-every", "2", "--out", "prog1.ppm", "--save-acc", "p1.acc"); assert p.returncode == 0, p.stdout + p.stderr
    p = run("--devices", "0,0", "--snapshot-every", "2", "--out", "prog2.ppm", "--save-acc", "p2.acc"); assert p.returncode == 0, p.stdout + p.stderr
    assert open(tmp_path / "p1.acc", "rb").read() == open(tmp_path / "p2.acc", "rb").read()            # the float accumulators
    assert open(tmp_path / "prog1.ppm", "rb").read() == open(tmp_path / "prog2.ppm", "rb").read()
    p = run("--devices", "0,99"); assert p.returncode == 1 and "cannot create device scene" in p.stderr

