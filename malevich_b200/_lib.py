"""ctypes binding of the C-ABI in include/malevich_b200.h (no torch types anywhere in the boundary)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLV_LIB_PATH") or os.path.join(_HERE, "csrc", "libmalevich_b200.so")  # (the override serves kernel experiments: variant builds side by side)

# every symbol include/malevich_b200.h declares (tests check the library exports each of them)
EXPORTED_SYMBOLS = [
    "mlv_last_error_string", "mlv_create_device", "mlv_destroy_device", "mlv_finish", "mlv_get_stream",
    "mlv_create_buffer", "mlv_update_buffer", "mlv_update_buffer_range", "mlv_buffer_device_ptr", "mlv_get_copy_stream", "mlv_buffer_mark_updated", "mlv_release_buffer", "mlv_create_texture2d", "mlv_release_texture", "mlv_texture_srgb_to_linear", "mlv_texture_generate_mips", "mlv_texture_mip_levels", "mlv_read_texture_mip", "mlv_read_texture",
    "mlv_ia_set_vertex_buffer", "mlv_ia_set_index_buffer", "mlv_ia_set_index_format", "mlv_ia_set_input_layout", "mlv_ia_set_primitive_topology",
    "mlv_vs_set_shader", "mlv_vs_set_constant_buffer", "mlv_vs_set_shader_resource", "mlv_rs_set_viewport",
    "mlv_ps_set_shader", "mlv_ps_set_shader_resource",
    "mlv_clear_render_target_view", "mlv_clear_depth_stencil_view", "mlv_draw_indexed", "mlv_draw_indexed_ex", "mlv_draw",
    "mlv_begin_command_list", "mlv_finish_command_list", "mlv_execute_command_list", "mlv_command_list_set_constants", "mlv_command_list_info", "mlv_release_command_list",
    "mlv_present_readback", "mlv_present_readback_async", "mlv_present_wait", "mlv_get_stats", "mlv_reset_stats", "mlv_get_work_counters",
    "mlv_resolve", "mlv_resolved_color_device_ptr", "mlv_resolved_depth_device_ptr",
    "mlv_composite_peer_export", "mlv_composite_peer_attach", "mlv_composite_broadcast", "mlv_composite_wait", "mlv_composite_broadcast_async", "mlv_composite_join", "mlv_composite_readback_async", "mlv_present_owned_rows_async", "mlv_register_host_memory", "mlv_unregister_host_memory", "mlv_composite_layout", "mlv_composite_pack", "mlv_composite_unpack",
    "mlv_debug_read_vs_out", "mlv_debug_read_triangles", "mlv_debug_read_bins", "mlv_debug_read_masks",
    "mlv_debug_read_tile_min_depths", "mlv_debug_read_keys", "mlv_read_bin_lists", "mlv_fnv64_words", "mlv_profile_begin", "mlv_profile_end", "mlv_profile_read_events", "mlv_kernel_launch_count",
    "mlv_timeline_begin", "mlv_timeline_end", "mlv_timeline_reset", "mlv_timeline_read",
]
STAGE_NAMES = ["clear", "geometry", "bin_count", "bin_scan", "bin_fill", "tile", "resolve", "composite", "vertex_cache", "clip", "geometry_back"]

MLV_OK = 0
MLV_ERR_INVALID_ARGUMENT, MLV_ERR_CUDA, MLV_ERR_OUT_OF_MEMORY, MLV_ERR_CAPACITY, MLV_ERR_STATE = 1, 2, 3, 4, 5
PRIMITIVE_TOPOLOGY_UNDEFINED, PRIMITIVE_TOPOLOGY_TRIANGLELIST = 0, 1
VS_PASSTHROUGH, VS_BASIC, VS_VERTEX_LIGHTING, VS_FULLSCREEN = 0, 1, 2, 3
PS_PASSTHROUGH, PS_BASIC, PS_ENV_LIGHTING, PS_BASIC_TRILINEAR = 0, 1, 2, 3
FORMAT_R8G8B8A8_UNORM, FORMAT_R32G32B32A32_FLOAT = 0, 1
BUFFER_VERTEX, BUFFER_INDEX = 0, 1
INDEX_U32, INDEX_U16 = 0, 1
DEVICE_DEBUG_CAPTURE, DEVICE_GROUP_SAME_GPU, DEVICE_GROUP_NCCL, DEVICE_GROUP_PEER_EXCHANGE = 1, 2, 4, 8
ALL_DRAWS = 0xFFFFFFFF


class WorkCounters(C.Structure):
    _fields_ = [("records_written", C.c_uint64), ("pairs_listed", C.c_uint64), ("tiles_visited", C.c_uint64)]


class ProfileEvent(C.Structure):
    _fields_ = [("stage", C.c_int32), ("start_ms", C.c_float), ("duration_ms", C.c_float)]


class TimelineEvent(C.Structure):
    _fields_ = [("stage", C.c_int32), ("draw", C.c_int32), ("resident_us", C.c_double), ("start_us", C.c_double), ("end_us", C.c_double)]


class Viewport(C.Structure):
    _fields_ = [("top_left_x", C.c_float), ("top_left_y", C.c_float), ("width", C.c_float), ("height", C.c_float),
                ("min_depth", C.c_float), ("max_depth", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("frame_time", C.c_float), ("vertex_count", C.c_uint32), ("input_triangle_count", C.c_uint32),
                ("assembled_triangle_count", C.c_uint32), ("active_bin_count", C.c_uint32),
                ("total_triangle_count_in_bins", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "frame_time"}


class PeerInfo(C.Structure):  # mlv_peer_info
    _fields_ = [("ipc_color", (C.c_ubyte * 64) * 2), ("ipc_flags", C.c_ubyte * 64), ("color", C.c_void_p * 2), ("flags", C.c_void_p),
                ("cuda_device", C.c_int32), ("reserved", C.c_int32)]


class DeviceDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("cuda_device", C.c_int32), ("num_ranks", C.c_uint32),
                ("rank", C.c_uint32), ("stripe_height_tiles", C.c_uint32), ("max_pairs_per_draw", C.c_uint64),
                ("flags", C.c_uint32), ("num_gpus", C.c_uint32)]


class MalevichError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"malevich_b200 error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Loads the CUDA extension. Fails loudly when it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C malevich_b200/csrc` or __graft_entry__.build(); "
                          "malevich_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u32, i32, f32, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_size_t
    P = C.POINTER
    sig = {
        "mlv_last_error_string": (C.c_char_p, []),
        "mlv_create_device": (i32, [P(DeviceDesc), P(vp)]),
        "mlv_destroy_device": (None, [vp]),
        "mlv_finish": (i32, [vp]),
        "mlv_get_stream": (vp, [vp]),
        "mlv_create_buffer": (i32, [vp, vp, sz, i32, P(vp)]),
        "mlv_update_buffer": (i32, [vp, vp, vp, sz]),
        "mlv_update_buffer_range": (i32, [vp, vp, sz, vp, sz]),
        "mlv_buffer_device_ptr": (vp, [vp]),
        "mlv_get_copy_stream": (vp, [vp]),
        "mlv_buffer_mark_updated": (i32, [vp, vp, vp]),
        "mlv_release_buffer": (None, [vp, vp]),
        "mlv_create_texture2d": (i32, [vp, vp, u32, u32, i32, P(vp)]),
        "mlv_release_texture": (None, [vp, vp]),
        "mlv_texture_srgb_to_linear": (i32, [vp, vp]),
        "mlv_read_texture": (i32, [vp, vp, vp]),
        "mlv_texture_generate_mips": (i32, [vp, vp]),
        "mlv_texture_mip_levels": (i32, [vp, P(u32)]),
        "mlv_read_texture_mip": (i32, [vp, vp, u32, vp]),
        "mlv_ia_set_vertex_buffer": (i32, [vp, vp]),
        "mlv_ia_set_index_buffer": (i32, [vp, vp]),
        "mlv_ia_set_index_format": (i32, [vp, i32]),
        "mlv_ia_set_input_layout": (i32, [vp, u32]),
        "mlv_ia_set_primitive_topology": (i32, [vp, i32]),
        "mlv_vs_set_shader": (i32, [vp, i32]),
        "mlv_vs_set_constant_buffer": (i32, [vp, u32, vp, sz]),
        "mlv_vs_set_shader_resource": (i32, [vp, u32, vp]),
        "mlv_rs_set_viewport": (i32, [vp, P(Viewport)]),
        "mlv_ps_set_shader": (i32, [vp, i32]),
        "mlv_ps_set_shader_resource": (i32, [vp, u32, vp]),
        "mlv_clear_render_target_view": (i32, [vp, P(f32)]),
        "mlv_clear_depth_stencil_view": (i32, [vp, f32]),
        "mlv_draw_indexed": (i32, [vp, u32]),
        "mlv_draw_indexed_ex": (i32, [vp, u32, u32, C.c_int32]),
        "mlv_draw": (i32, [vp, u32]),
        "mlv_begin_command_list": (i32, [vp]),
        "mlv_finish_command_list": (i32, [vp, P(vp)]),
        "mlv_execute_command_list": (i32, [vp, vp]),
        "mlv_command_list_set_constants": (i32, [vp, vp, u32, vp, sz]),
        "mlv_command_list_info": (i32, [vp, P(u32), P(C.c_uint64)]),
        "mlv_release_command_list": (None, [vp, vp]),
        "mlv_present_readback": (i32, [vp, vp, vp]),
        "mlv_present_readback_async": (i32, [vp, vp, vp]),
        "mlv_present_wait": (i32, [vp]),
        "mlv_get_stats": (i32, [vp, P(Stats)]),
        "mlv_reset_stats": (i32, [vp]),
        "mlv_get_work_counters": (i32, [vp, P(WorkCounters)]),
        "mlv_resolve": (i32, [vp]),
        "mlv_resolved_color_device_ptr": (vp, [vp]),
        "mlv_resolved_depth_device_ptr": (vp, [vp]),
        "mlv_composite_layout": (i32, [vp, P(vp), P(sz)]),
        "mlv_composite_peer_export": (i32, [vp, vp]),
        "mlv_composite_peer_attach": (i32, [vp, vp, i32]),
        "mlv_composite_broadcast": (i32, [vp]),
        "mlv_composite_wait": (i32, [vp]),
        "mlv_composite_broadcast_async": (i32, [vp]),
        "mlv_composite_join": (i32, [vp]),
        "mlv_composite_readback_async": (i32, [vp, vp]),
        "mlv_present_owned_rows_async": (i32, [vp, vp]),
        "mlv_register_host_memory": (i32, [vp, vp, C.c_size_t]),
        "mlv_unregister_host_memory": (i32, [vp, vp]),
        "mlv_composite_pack": (i32, [vp]),
        "mlv_composite_unpack": (i32, [vp]),
        "mlv_debug_read_vs_out": (i32, [vp, vp, P(u32)]),
        "mlv_debug_read_triangles": (i32, [vp, vp, vp, P(u32)]),
        "mlv_debug_read_bins": (i32, [vp, vp, P(u32), vp, P(u32)]),
        "mlv_debug_read_masks": (i32, [vp, vp, P(u32)]),
        "mlv_debug_read_tile_min_depths": (i32, [vp, vp]),
        "mlv_debug_read_keys": (i32, [vp, vp, P(u32)]),
        "mlv_read_bin_lists": (i32, [vp, vp, P(u32), vp, P(u32)]),
        "mlv_fnv64_words": (C.c_uint64, [vp, sz]),
        "mlv_profile_begin": (i32, [vp]),
        "mlv_profile_end": (i32, [vp, P(C.c_double), P(u32)]),
        "mlv_profile_read_events": (i32, [vp, vp, u32, P(u32)]),
        "mlv_timeline_begin": (i32, [vp]),
        "mlv_timeline_end": (i32, [vp]),
        "mlv_timeline_reset": (i32, [vp]),
        "mlv_timeline_read": (i32, [vp, vp, u32, P(u32)]),
        "mlv_kernel_launch_count": (C.c_uint64, [vp]),
    }
    assert sorted(sig) == sorted(EXPORTED_SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != MLV_OK:
        raise MalevichError(code, load().mlv_last_error_string().decode("utf-8", "replace"))


def fnv64_words(a) -> str:
    """Frame hash (word-wise 64-bit FNV-1a over u32 words) as 16 hex digits -- the format of tests/golden/golden.json."""
    import numpy as np
    w = np.ascontiguousarray(a).view(np.uint32).ravel()
    return "%016x" % load().mlv_fnv64_words(w.ctypes.data_as(C.c_void_p), w.size)
