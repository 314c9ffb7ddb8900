"""The C-ABI library loads without a GPU and exports every symbol include/dogeray_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import dogeray_b200 as drb
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dogeray_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(drb.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libdogeray_b200.so does not export %s" % n
    assert sorted(drb.EXPORTED_SYMBOLS) == names


def test_abi_version_and_struct_sizes():
    assert drb._lib.drb_abi_version() == 2
    assert ctypes.sizeof(drb.Settings) == 60
    assert drb.OBJECT_DTYPE.itemsize == 156
    assert ctypes.sizeof(drb.Opts) == 40
    assert ctypes.sizeof(drb.Stats) == 32


def test_default_settings_are_reference_defaults():
    s = drb.default_settings()      # raygpu/kernel.cu:29-30, 119-132
    assert list(s.cam) == [0, 0, 2] and list(s.look) == [0, 0, 0]
    assert s.aperture == np.float32(0.01) and s.focus == 3 and s.fov == 45
    assert s.max_depth == 50 and s.spp == 1 and s.bg_intensity == 1 and s.backtex == -1
    assert (s.width, s.height) == (1280, 720)


def test_no_cpu_fallback_without_device(tmp_path):
    if drb.device_count() > 0:
        pytest.skip("a GPU is present")
    objs = drb.make_objects(1)
    hs = drb.HostScene.from_objects(objs)
    with pytest.raises(drb.DogerayError) as e:
        drb.Scene.from_host(hs)
    assert e.value.status == drb.ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_error_codes():
    with pytest.raises(drb.DogerayError) as e:
        drb.HostScene.load("/nonexistent/scene.rts")
    assert e.value.status == drb.ERR_IO
    with pytest.raises(drb.DogerayError) as e:
        drb.HostScene.parse(b"1,2,abc,2,0,0,0\n")
    assert e.value.status == drb.ERR_PARSE and "column 2" in str(e.value)
    with pytest.raises(drb.DogerayError) as e:
        drb.read_ppm("/nonexistent.ppm")
    assert e.value.status == drb.ERR_IO


def test_render_multi_rejects_bad_arguments_before_touching_cuda():
    import ctypes as C
    st = drb.Settings(); drb._lib.drb_settings_default(C.byref(st))
    buf = (C.c_float * (st.width * st.height * 3))()
    none = (C.c_void_p * 2)(None, None)
    assert drb._lib.drb_render_multi(none, 0, C.byref(st), None, buf, None) == drb.ERR_ARG
    assert drb._lib.drb_render_multi(None, 2, C.byref(st), None, buf, None) == drb.ERR_ARG
    assert drb._lib.drb_render_multi(none, 2, C.byref(st), None, buf, None) == drb.ERR_ARG and b"scene 0 is null" in drb._lib.drb_last_error()
    assert drb._lib.drb_render_multi(none, 2, C.byref(st), None, None, None) == drb.ERR_ARG



def _build_c_consumer(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_consumer")
    libdir = os.path.dirname(drb.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "native", "abi_consumer.c"), "-o", exe, "-L" + libdir, "-ldogeray_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr                                   # the header is valid, warning-free C99
    return subprocess.run([exe], capture_output=True, text=True, timeout=300)


def test_header_is_plain_c_and_struct_sizes_agree(tmp_path):
    r = _build_c_consumer(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sizes settings=%d object=%d opts=%d stats=%d build_info=%d abi=2" % (
        ctypes.sizeof(drb.Settings), drb.OBJECT_DTYPE.itemsize, ctypes.sizeof(drb.Opts), ctypes.sizeof(drb.Stats), ctypes.sizeof(drb.BuildInfo)) in r.stdout
    assert "parsed objects=2 width=16 height=8" in r.stdout
    if drb.device_count() == 0:
        assert "no device: drb_scene_create -> -4" in r.stdout and "no CPU path" in r.stdout


@pytest.mark.gpu
def test_c_consumer_renders_through_the_abi(tmp_path):
    r = _build_c_consumer(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rendered paths=256" in r.stdout and "multi paths=256" in r.stdout and "identical" in r.stdout
