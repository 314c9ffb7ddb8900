"""TEST INFRASTRUCTURE -- ctypes access to oracle/_ref/libdogeray_ref_host.so, the reference's own
functions compiled for the host by oracle/make_ref.py.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(REF_DIR, "libdogeray_ref_host.so")
SAMPLES = os.path.join(REF_DIR, "samples")

# numpy view of the reference's `singleobject` (raygpu/kernel.cu:48-74; 164 bytes, offsets SURVEY.md row a1)
SINGLEOBJECT = np.dtype({
    "names": ["type", "pos", "rot", "norm", "n1", "n2", "n3", "t1", "t2", "t3", "smooth", "tex", "mat", "dim", "col",
              "texnum", "rtexnum", "addional"],
    "formats": ["<i4", ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3),
                ("<f4", 3), "u1", "u1", "<i4", ("<f4", 3), ("<f4", 3), "<i4", "<i4", ("<f4", 3)],
    "offsets": [0, 4, 16, 28, 40, 52, 64, 76, 88, 100, 112, 113, 116, 120, 132, 144, 148, 152],
    "itemsize": 164,
})


def available() -> bool:
    return os.path.exists(LIB)


class RefHost:
    """One process-wide reference scene (the reference keeps its scene in file-scope globals)."""

    def __init__(self):
        if not available():
            raise FileNotFoundError(LIB + " -- run `python oracle/make_ref.py` where /root/reference exists")
        L = C.CDLL(LIB)
        L.ref_load.argtypes = [C.c_char_p, C.c_char_p]; L.ref_load.restype = C.c_int
        L.ref_objects.restype = C.c_void_p
        L.ref_nodes.restype = C.c_void_p
        L.ref_texture_path.argtypes = [C.c_int]; L.ref_texture_path.restype = C.c_char_p
        L.ref_get_settings.argtypes = [C.c_void_p]
        L.ref_set_settings.argtypes = [C.c_void_p]
        L.ref_set_seed.argtypes = [C.c_ulonglong]
        L.ref_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_singlehit.argtypes = [C.c_void_p, C.c_void_p, C.c_int]; L.ref_singlehit.restype = C.c_float
        L.ref_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_int]
        L.ref_last_rays.restype = C.c_ulonglong
        L.ref_raycolor.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.L = L
        assert L.ref_sizeof_object() == SINGLEOBJECT.itemsize

    def load(self, rts_path: str, tex_dir: str = "") -> int:
        n = self.L.ref_load(os.fsencode(rts_path), os.fsencode(tex_dir or ""))
        if n < 0:
            raise IOError("reference could not open %s" % rts_path)
        return n

    @property
    def num_objects(self) -> int:
        return max(self.L.ref_num_objects(), 0)

    def objects(self) -> np.ndarray:
        n = self.num_objects
        if n == 0:
            return np.zeros(0, SINGLEOBJECT)
        buf = (C.c_char * (n * 164)).from_address(self.L.ref_objects())
        return np.frombuffer(buf, SINGLEOBJECT).copy()

    @property
    def num_nodes(self) -> int:
        return self.L.ref_num_nodes()

    def texture_paths(self):
        return [os.fsdecode(self.L.ref_texture_path(i)) for i in range(self.L.ref_num_textures())]

    # settings vector: cam xyz, aperture, look xyz, focus, fov, depth, spp, bg, backtex, W, H, 0
    def get_settings(self) -> np.ndarray:
        s = np.zeros(16, np.float32)
        self.L.ref_get_settings(s.ctypes.data)
        return s

    def set_settings(self, s):
        s = np.ascontiguousarray(s, np.float32)
        self.L.ref_set_settings(s.ctypes.data)

    def apply(self, st):
        """Push a dogeray_b200.Settings into the reference's globals."""
        self.set_settings([st.cam[0], st.cam[1], st.cam[2], st.aperture, st.look[0], st.look[1], st.look[2], st.focus,
                           st.fov, st.max_depth, st.spp, st.bg_intensity, st.backtex, st.width, st.height, 0])

    def set_seed(self, seed: int):
        self.L.ref_set_seed(seed)

    def hit(self, origins, dirs, threads=0):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        t = np.empty(len(o), np.float32); ids = np.empty(len(o), np.int32)
        self.L.ref_hit(o.ctypes.data, d.ctypes.data, len(o), t.ctypes.data, ids.ctypes.data, threads or (os.cpu_count() or 1))
        return ids, t

    def singlehit(self, o, d, obj: int) -> float:
        o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
        return float(self.L.ref_singlehit(o.ctypes.data, d.ctypes.data, int(obj)))

    def frame(self, divisor=1, sample_base=0, threads=0):
        """Kernel() over the launched grid.  Returns (float32 (W,H,3) = 255*mean unquantised, int32 (W,H,3), rays)."""
        s = self.get_settings()
        W, H = int(s[13]), int(s[14])
        f = np.zeros((W, H, 3), np.float32); i = np.zeros((W, H, 3), np.int32)
        self.L.ref_frame(f.ctypes.data, i.ctypes.data, divisor, sample_base, threads or (os.cpu_count() or 1))
        return f, i, int(self.L.ref_last_rays())

    def raycolor(self, origins, dirs, xy, samples, depth):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        sm = np.ascontiguousarray(samples, np.uint32).reshape(-1)
        rgb = np.empty((len(o), 3), np.float32)
        self.L.ref_raycolor(o.ctypes.data, d.ctypes.data, xy.ctypes.data, sm.ctypes.data, len(o), depth, rgb.ctypes.data)
        return rgb
