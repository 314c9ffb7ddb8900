"""Host-side code under AddressSanitizer + UndefinedBehaviorSanitizer (the reference has no such checks, SURVEY.md 5;
its loader reads past short lines and leaves fields indeterminate -- this one must not)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import dogeray_b200 as drb
from dogeray_b200 import synth
from conftest import ROOT, SAMPLES

CSRC = os.path.join(ROOT, "dogeray_b200", "csrc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("san") / "sanitize_host")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, os.path.join(ROOT, "tests", "native", "sanitize_host_main.cpp"),
           os.path.join(CSRC, "rts_loader.cpp"), os.path.join(CSRC, "image_io.cpp"), "-o", out, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + r.stderr[-300:])
    return out


def run(harness, tmp_path, files):
    env = dict(os.environ, DRB_SAN_TMP=str(tmp_path), ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([harness] + files, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "bad 0" in r.stdout and "ERROR" not in r.stderr and "runtime error" not in r.stderr


def test_loader_cache_and_writer_are_clean_on_written_scenes(harness, tmp_path):
    files = []
    for k, n in enumerate((1, 7, 40)):
        objs, st = synth.heightfield_scene(n, seed=k)
        p = str(tmp_path / ("h%d.rts" % k))
        drb.write_rts(p, st, objs)
        files.append(p)
    big, st = synth.heightfield_scene(160, seed=9)                      # > 20000 objects: the threaded parse
    p = str(tmp_path / "big.rts"); drb.write_rts(p, st, big); files.append(p)
    ragged = str(tmp_path / "ragged.rts")
    open(ragged, "wb").write(b"/c\r\n*,1,2\r\n2\n2,1\n\n,,,\n0,0,0,0,1,1,1,0,0,1" + b",9" * 60 + b"\nr,r,r,2,r,r,r,r,r,r,r,r,r,r,r,r")
    files.append(ragged)
    empty = str(tmp_path / "empty.rts"); open(empty, "wb").close(); files.append(empty)
    run(harness, tmp_path, files)


@pytest.mark.skipif(not os.path.isdir(SAMPLES), reason="sample scenes live under /root/reference")
def test_loader_cache_and_writer_are_clean_on_every_shipped_scene(harness, tmp_path):
    files = sorted(os.path.join(SAMPLES, n) for n in os.listdir(SAMPLES) if n.endswith(".rts"))
    run(harness, tmp_path, files)


def test_threaded_parse_hash_and_writer_are_race_free(tmp_path):
    """ThreadSanitizer over the parallel paths: line-range parse threads, chunk hash threads, the writer's workers"""
    out = str(tmp_path / "tsan_host")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
           os.path.join(ROOT, "tests", "native", "sanitize_host_main.cpp"), os.path.join(CSRC, "rts_loader.cpp"),
           os.path.join(CSRC, "image_io.cpp"), "-o", out, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("tsan runtime not available: " + r.stderr[-300:])
    objs, st = synth.heightfield_scene(200, seed=1)                     # 80 000 triangles, 26 MB of text: every pool is used
    p = str(tmp_path / "t.rts")
    drb.write_rts(p, st, objs)
    r = subprocess.run([out, p], capture_output=True, text=True, env=dict(os.environ, DRB_SAN_TMP=str(tmp_path)), timeout=600)
    if "unexpected memory mapping" in r.stderr:
        pytest.skip("tsan cannot run in this container")
    assert r.returncode == 0 and "bad 0" in r.stdout and "WARNING: ThreadSanitizer" not in r.stderr, (r.stdout + r.stderr)[-3000:]
