"""oracle/ref_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/_ref/libmalevich_ref_<W>x<H>.so, i.e. over the REFERENCE's own pipeline
code compiled by oracle/build_ref.sh (kind = "reference", not a port). One library per resolution
because the reference fixes WIDTH/HEIGHT at compile time (main.c:21-22).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

TRIANGLE_DTYPE = np.dtype([("p_attributes", "<u8"), ("min_bounds", "<i4", (2,)), ("max_bounds", "<i4", (2,)),
                           ("edges", "<i4", (3, 3)), ("reciprocal_ws", "<f4", (3,)), ("one_over_area", "<f4"), ("max_depth", "<f4")])
COMPACTED_BIN_DTYPE = np.dtype([("num_triangles_self", "<u4"), ("num_triangles_upto", "<u4"), ("bin_index", "<u4")])
TILE_INFO_DTYPE = np.dtype([("triangle_id", "<u4"), ("_pad", "<u4"), ("fragment_mask", "<u8")])
STATS_FIELDS = ["vertex_count", "input_triangle_count", "assembled_triangle_count", "active_bin_count", "total_triangle_count_in_bins"]


def available_resolutions():
    out = []
    if os.path.isdir(REF_DIR):
        for f in os.listdir(REF_DIR):
            if f.startswith("libmalevich_ref_") and f.endswith(".so"):
                w, h = f[len("libmalevich_ref_"):-3].split("x")
                out.append((int(w), int(h)))
    return sorted(out)


def lib_path(width: int, height: int) -> str:
    return os.path.join(REF_DIR, f"libmalevich_ref_{width}x{height}.so")


def have(width: int, height: int) -> bool:
    return os.path.exists(lib_path(width, height))


class RefOracle:
    """The reference renderer at one fixed resolution."""

    _loaded = {}

    def __init__(self, width: int, height: int, threads: int = 1):
        path = lib_path(width, height)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run oracle/build_ref.sh {width}x{height} (needs /root/reference)")
        lib = RefOracle._loaded.get(path)
        if lib is None:
            lib = C.CDLL(path)
            vp, u32, i32, f32 = C.c_void_p, C.c_uint32, C.c_int, C.c_float
            lib.ref_camera.restype, lib.ref_camera.argtypes = vp, [f32] * 5
            lib.ref_set_cb0.argtypes = [vp]
            lib.ref_clear.argtypes = [vp, f32]
            for n in ("ref_draw", "ref_draw_staged"):
                getattr(lib, n).argtypes = [vp, vp, u32, i32, i32, vp, u32, u32]
            for n in ("ref_staged_vs_out", "ref_staged_triangles", "ref_staged_triangle_ids", "ref_staged_compacted_bins",
                      "ref_suprematist_vb", "ref_suprematist_ib", "ref_fullscreen_vb", "ref_fullscreen_ib"):
                getattr(lib, n).restype, getattr(lib, n).argtypes = vp, [C.POINTER(u32)]
            for n in ("ref_staged_attributes", "ref_staged_tile_infos", "ref_colors", "ref_depths", "ref_tile_min_depths", "ref_stats"):
                getattr(lib, n).restype = vp
            lib.ref_prof_get.argtypes = [i32, C.POINTER(C.c_char_p), C.POINTER(C.c_double)]
            lib.ref_rsqrt.restype, lib.ref_rsqrt.argtypes = f32, [f32]
            lib.ref_texture_srgb_to_linear.argtypes = [vp, u32, u32]
            RefOracle._loaded[path] = lib
        self.lib = lib
        self.width, self.height = lib.ref_width(), lib.ref_height()
        assert (self.width, self.height) == (width, height)
        self.set_threads(threads)
        self._keep = []

    def set_threads(self, n: int):
        self.lib.ref_set_threads(int(n))

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    def camera(self, pos=(3.5, 1.0, 1.0), yaw=0.0, pitch=0.0) -> np.ndarray:
        p = self.lib.ref_camera(pos[0], pos[1], pos[2], yaw, pitch)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (3, 4, 4)).copy()

    def begin_frame(self, cb: np.ndarray, clear_color=None, clear_depth=0.0):
        cb = np.ascontiguousarray(cb, dtype=np.float32)
        assert cb.nbytes == 192
        self.lib.ref_set_cb0(cb.ctypes.data_as(C.c_void_p))
        self.lib.ref_begin_frame()
        if clear_color is not None:
            cc = (C.c_float * 4)(*[float(np.float32(x)) for x in clear_color])
            self.lib.ref_clear(cc, float(clear_depth))

    def draw(self, vb, ib, vs_id, ps_id, tex=None, staged=False):
        vb = np.ascontiguousarray(vb)
        ib = np.ascontiguousarray(ib, dtype=np.uint32)
        self._keep = [vb, ib, tex]
        fn = self.lib.ref_draw_staged if staged else self.lib.ref_draw
        if tex is None:
            fn(vb.ctypes.data_as(C.c_void_p), ib.ctypes.data_as(C.c_void_p), ib.shape[0], vs_id, ps_id, None, 0, 0)
        else:
            fn(vb.ctypes.data_as(C.c_void_p), ib.ctypes.data_as(C.c_void_p), ib.shape[0], vs_id, ps_id,
               tex.p_data.ctypes.data_as(C.c_void_p), tex.width, tex.height)

    def render(self, scene, staged_last=False):
        """Mirrors malevich_b200.scenes.render for a Scene object (clears + every draw)."""
        from malevich_b200 import scenes as S
        self.begin_frame(scene.per_frame_cb, S.CLEAR_COLOR, S.CLEAR_DEPTH)
        for i, o in enumerate(scene.objects):
            self.draw(o.vertex_buffer, o.index_buffer, o.vertex_shader.vs_main, o.pixel_shader.ps_main, o.texture,
                      staged=staged_last and i == len(scene.objects) - 1)

    def _arr(self, ptr, ctype, shape):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape).copy()

    def colors(self) -> np.ndarray:
        return self._arr(self.lib.ref_colors(), C.c_uint32, (self.height, self.width))

    def depths(self) -> np.ndarray:
        return self._arr(self.lib.ref_depths(), C.c_float, (self.height, self.width))

    def tile_min_depths(self) -> np.ndarray:
        return self._arr(self.lib.ref_tile_min_depths(), C.c_float, ((self.height // 8) * (self.width // 8),))

    def stats(self) -> dict:
        a = self._arr(self.lib.ref_stats(), C.c_uint32, (6,))
        return dict(zip(STATS_FIELDS, [int(x) for x in a[1:]]))

    # staged intermediates of the last draw(staged=True)
    def _struct_arr(self, ptr, dtype, n):
        if n == 0:
            return np.empty(0, dtype=dtype)
        buf = (C.c_char * (dtype.itemsize * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).copy()

    def staged_vs_out(self) -> np.ndarray:
        n = C.c_uint32()
        p = self.lib.ref_staged_vs_out(C.byref(n))
        return self._arr(p, C.c_float, (n.value, 12))

    def staged_triangles(self):
        n = C.c_uint32()
        p = self.lib.ref_staged_triangles(C.byref(n))
        tris = self._struct_arr(p, TRIANGLE_DTYPE, n.value)
        attrs = self._arr(self.lib.ref_staged_attributes(), C.c_float, (n.value, 9, 4)) if n.value else np.empty((0, 9, 4), np.float32)
        return tris, attrs

    def staged_bins(self):
        n, m = C.c_uint32(), C.c_uint32()
        pi = self.lib.ref_staged_triangle_ids(C.byref(n))
        pb = self.lib.ref_staged_compacted_bins(C.byref(m))
        ids = self._arr(pi, C.c_uint32, (n.value,)) if n.value else np.empty(0, np.uint32)
        return ids, self._struct_arr(pb, COMPACTED_BIN_DTYPE, m.value)

    def staged_tile_infos(self) -> np.ndarray:
        n = C.c_uint32()
        self.lib.ref_staged_triangle_ids(C.byref(n))
        return self._struct_arr(self.lib.ref_staged_tile_infos(), TILE_INFO_DTYPE, n.value)

    def prof_reset(self):
        self.lib.ref_prof_reset()

    def prof(self) -> dict:
        out, i = {}, 0
        name, ms = C.c_char_p(), C.c_double()
        while self.lib.ref_prof_get(i, C.byref(name), C.byref(ms)):
            out[name.value.decode()] = ms.value
            i += 1
        return out

    def rsqrt(self, x: float) -> float:
        return float(self.lib.ref_rsqrt(float(x)))

    def texture_srgb_to_linear(self, tex_u32: np.ndarray) -> np.ndarray:
        t = np.ascontiguousarray(tex_u32, dtype=np.uint32).copy()
        self.lib.ref_texture_srgb_to_linear(t.ctypes.data_as(C.c_void_p), t.shape[1], t.shape[0])
        return t


def fnv64_words(a: np.ndarray) -> str:
    """word-wise FNV-1a over u32 words (h = 0xcbf29ce484222325; h ^= w; h *= 0x100000001b3)."""
    w = np.ascontiguousarray(a).view(np.uint32).ravel()
    h = 0xcbf29ce484222325
    for x in w.tolist():
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
