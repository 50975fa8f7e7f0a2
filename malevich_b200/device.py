"""Host-side mirror of the reference's D3D11-mimicking interface (reference source/main.c).

The reference host (`render()`, main.c:1265-1299) writes the fields of a global `Pipeline
graphics_pipeline` (main.c:71-115,222) and then calls `clear_render_target_view` (:1191),
`clear_depth_stencil_view` (:1204) and `draw_indexed` (:1219).  `Device` keeps exactly that shape:
`dev.graphics_pipeline.{ia,vs,rs,ps}` carry the same field names, the three entry points have the
same names and argument meaning, and the reference's asserts surface as `MalevichError`.
Underneath, every call goes through the C-ABI of include/malevich_b200.h; host arrays bound as
vertex/index buffers and textures are uploaded once and cached by identity, like D3D11 resources.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib as L


@dataclass(frozen=True)
class VertexShader:
    """VertexShader descriptor, common_shader_core.h:10-14 ({in_vertex_size, out_vertex_size, vs_main})."""
    in_vertex_size: int
    out_vertex_size: int
    vs_main: int  # device shader id instead of a host function pointer


@dataclass(frozen=True)
class PixelShader:
    """PixelShader descriptor, common_shader_core.h:16-18."""
    ps_main: int


# the seven descriptors the reference exports (main.c:46-52); sizes are sizeof(Vs_Input/Vs_Output) of 8-wide SoA blocks
passthrough_vs = VertexShader(256, 384, L.VS_PASSTHROUGH)
basic_vs = VertexShader(256, 384, L.VS_BASIC)
vertex_lighting_vs = VertexShader(256, 384, L.VS_VERTEX_LIGHTING)
fullscreen_vs = VertexShader(256, 384, L.VS_FULLSCREEN)
passthrough_ps = PixelShader(L.PS_PASSTHROUGH)
basic_ps = PixelShader(L.PS_BASIC)
env_lighting_ps = PixelShader(L.PS_ENV_LIGHTING)
# extension: basic_ps with SRV0 sampled trilinearly from a device-built mip chain
basic_trilinear_ps = PixelShader(L.PS_BASIC_TRILINEAR)

VECTOR_WIDTH = 8  # main.c:26


class Texture2D:
    """Texture2D, common_shader_core.h:20-24. `p_data` is uint32 [h, w] (R8G8B8A8) or float32 [h, w, 4].
    `is_in_srgb` is load_texture's argument (main.c:538): the device copy is re-quantised to linear on the GPU."""

    def __init__(self, p_data: np.ndarray, is_in_srgb: bool = False, generate_mips: bool = False):
        self.is_in_srgb = bool(is_in_srgb)
        self.generate_mips = bool(generate_mips)  # extension: mlv_texture_generate_mips after the upload
        if p_data.dtype == np.uint32 and p_data.ndim == 2:
            self.format = L.FORMAT_R8G8B8A8_UNORM
        elif p_data.dtype == np.float32 and p_data.ndim == 3 and p_data.shape[2] == 4:
            self.format = L.FORMAT_R32G32B32A32_FLOAT
        else:
            raise ValueError("Texture2D wants uint32 [h,w] or float32 [h,w,4]")
        self.p_data = np.ascontiguousarray(p_data)
        self.height, self.width = int(p_data.shape[0]), int(p_data.shape[1])


@dataclass
class Viewport:  # main.c:85-92
    top_left_x: float = 0.0
    top_left_y: float = 0.0
    width: float = 0.0
    height: float = 0.0
    min_depth: float = 0.0
    max_depth: float = 1.0


@dataclass
class IA:  # main.c:71-76
    p_index_buffer: Optional[np.ndarray] = None
    p_vertex_buffer: Optional[np.ndarray] = None
    input_layout: int = 0
    primitive_topology: int = L.PRIMITIVE_TOPOLOGY_UNDEFINED


@dataclass
class VS:  # main.c:78-83
    shader: Optional[VertexShader] = None
    output_register_count: int = 0
    p_constant_buffers: list = field(default_factory=lambda: [None] * 16)
    p_shader_resource_views: list = field(default_factory=lambda: [None] * 16)


@dataclass
class RS:  # main.c:94-96
    viewport: Viewport = field(default_factory=Viewport)


@dataclass
class PS:  # main.c:98-101
    shader: Optional[PixelShader] = None
    p_shader_resource_views: list = field(default_factory=lambda: [None] * 16)


@dataclass
class Pipeline:  # main.c:109-115 (om.p_colors / om.p_depth are owned by the device here)
    ia: IA = field(default_factory=IA)
    vs: VS = field(default_factory=VS)
    rs: RS = field(default_factory=RS)
    ps: PS = field(default_factory=PS)


REF_TRIANGLE_DTYPE = np.dtype([("p_attributes", "<u8"), ("min_bounds", "<i4", (2,)), ("max_bounds", "<i4", (2,)),
                               ("edges", "<i4", (3, 3)), ("reciprocal_ws", "<f4", (3,)), ("one_over_area", "<f4"),
                               ("max_depth", "<f4")])
REF_COMPACTED_BIN_DTYPE = np.dtype([("num_triangles_self", "<u4"), ("num_triangles_upto", "<u4"), ("bin_index", "<u4")])
REF_TILE_INFO_DTYPE = np.dtype([("triangle_id", "<u4"), ("_pad", "<u4"), ("fragment_mask", "<u8")])
assert REF_TRIANGLE_DTYPE.itemsize == 80 and REF_COMPACTED_BIN_DTYPE.itemsize == 12 and REF_TILE_INFO_DTYPE.itemsize == 16


class Device:
    """One B200 running the draw pipeline for a width x height render target."""

    def __init__(self, width: int, height: int, cuda_device: int = -1, num_ranks: int = 1, rank: int = 0,
                 stripe_height_tiles: int = 1, debug_capture: bool = False, max_pairs_per_draw: int = 0,
                 num_gpus: int = 0, group_same_gpu: bool = False, group_nccl: bool = False, group_peer_exchange: bool = False):
        """num_gpus > 1: a device GROUP -- one host thread, num_gpus CUDA devices in this process, the sort-first fan-out
        and the exchange of the stripes behind the same calls (include/malevich_b200.h, DEVICE GROUPS); stripe_height_tiles
        0 = one contiguous band per GPU. group_same_gpu puts every rank on one CUDA device (tests on a single-GPU box)."""
        self._lib = L.load()
        self.width, self.height = int(width), int(height)
        self.num_ranks, self.rank, self.stripe_height_tiles = num_ranks, rank, stripe_height_tiles
        desc = L.DeviceDesc(width=width, height=height, cuda_device=cuda_device, num_ranks=num_ranks, rank=rank,
                            stripe_height_tiles=stripe_height_tiles, max_pairs_per_draw=max_pairs_per_draw,
                            flags=(L.DEVICE_DEBUG_CAPTURE if debug_capture else 0) | (L.DEVICE_GROUP_SAME_GPU if group_same_gpu else 0) |
                            (L.DEVICE_GROUP_NCCL if group_nccl else 0) | (L.DEVICE_GROUP_PEER_EXCHANGE if group_peer_exchange else 0), num_gpus=num_gpus)
        self._h = C.c_void_p()
        L.check(self._lib.mlv_create_device(C.byref(desc), C.byref(self._h)))
        self.graphics_pipeline = Pipeline()
        self._buffers = {}   # id(array) -> (array, handle)
        self._textures = {}  # id(Texture2D) -> (tex, handle)
        self._sent = {}      # pipeline state as last sent through the C-ABI (see _bind)

    # ---- lifetime ------------------------------------------------------------------------------
    def close(self):
        if self._h:
            for _, h in self._buffers.values():
                self._lib.mlv_release_buffer(self._h, h)
            for _, h in self._textures.values():
                self._lib.mlv_release_texture(self._h, h)
            self._buffers.clear()
            self._textures.clear()
            self._lib.mlv_destroy_device(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- resources -----------------------------------------------------------------------------
    def _buffer(self, arr: np.ndarray, kind: int):
        key = (id(arr), kind)
        hit = self._buffers.get(key)
        if hit is not None and hit[0] is arr:
            return hit[1]
        a = np.ascontiguousarray(arr)
        h = C.c_void_p()
        L.check(self._lib.mlv_create_buffer(self._h, a.ctypes.data_as(C.c_void_p), a.nbytes, kind, C.byref(h)))
        self._buffers[key] = (arr, h)
        return h

    def adopt_buffer(self, arr: np.ndarray, kind: int, capacity_bytes: int):
        """Device copy of `arr` in an allocation of `capacity_bytes` >= arr.nbytes (a padded size for in-place collectives
        over the buffer, see mlv_update_buffer_range). Later draws that bind `arr` use it."""
        a = np.ascontiguousarray(arr)
        h = C.c_void_p()
        L.check(self._lib.mlv_create_buffer(self._h, None, max(int(capacity_bytes), a.nbytes), kind, C.byref(h)))
        L.check(self._lib.mlv_update_buffer(self._h, h, a.ctypes.data_as(C.c_void_p), a.nbytes))
        self._buffers[(id(arr), kind)] = (arr, h)
        return h

    @property
    def copy_stream(self) -> int:
        return int(self._lib.mlv_get_copy_stream(self._h) or 0)

    def composite_readback_async(self, colors: np.ndarray):
        """Non-blocking read-back of the composited image (after composite_wait / composite_join) into a page-locked array."""
        L.check(self._lib.mlv_composite_readback_async(self._h, colors.ctypes.data_as(C.c_void_p)))

    def _texture(self, tex: Texture2D):
        hit = self._textures.get(id(tex))
        if hit is not None and hit[0] is tex:
            return hit[1]
        h = C.c_void_p()
        L.check(self._lib.mlv_create_texture2d(self._h, tex.p_data.ctypes.data_as(C.c_void_p), tex.width, tex.height, tex.format, C.byref(h)))
        if tex.is_in_srgb:
            L.check(self._lib.mlv_texture_srgb_to_linear(self._h, h))
        if tex.generate_mips:
            L.check(self._lib.mlv_texture_generate_mips(self._h, h))
        self._textures[id(tex)] = (tex, h)
        return h

    def read_texture(self, tex: Texture2D) -> np.ndarray:
        """The device copy of a texture (after the sRGB re-quantisation, if any)."""
        out = np.empty_like(tex.p_data)
        L.check(self._lib.mlv_read_texture(self._h, self._texture(tex), out.ctypes.data_as(C.c_void_p)))
        return out

    def read_texture_mips(self, tex: Texture2D) -> list:
        """Every level of the device's mip chain of `tex` (level 0 first) as uint32 [h, w] arrays."""
        h = self._texture(tex)
        n = C.c_uint32()
        L.check(self._lib.mlv_texture_mip_levels(h, C.byref(n)))
        out = []
        for level in range(n.value):
            a = np.empty((max(1, tex.height >> level), max(1, tex.width >> level)), dtype=np.uint32)
            L.check(self._lib.mlv_read_texture_mip(self._h, h, level, a.ctypes.data_as(C.c_void_p)))
            out.append(a)
        return out

    def upload(self, *objs):
        """Optional: create device copies ahead of the first draw (keeps uploads out of a timed region)."""
        for o in objs:
            if isinstance(o, Texture2D):
                self._texture(o)
            elif isinstance(o, np.ndarray):
                self._buffer(o, L.BUFFER_INDEX if o.dtype in (np.uint32, np.uint16) and o.ndim == 1 else L.BUFFER_VERTEX)

    def invalidate(self, obj):
        """Forget the device copy of a host array / texture whose contents changed."""
        self._sent.clear()
        if isinstance(obj, Texture2D):
            hit = self._textures.pop(id(obj), None)
            if hit:
                self._lib.mlv_release_texture(self._h, hit[1])
        else:
            for kind in (L.BUFFER_VERTEX, L.BUFFER_INDEX):
                hit = self._buffers.pop((id(obj), kind), None)
                if hit:
                    self._lib.mlv_release_buffer(self._h, hit[1])

    # ---- the reference's three entry points ---------------------------------------------------
    def clear_render_target_view(self, p_clear_color):  # main.c:1191
        c = (C.c_float * 4)(*[float(np.float32(x)) for x in p_clear_color])
        L.check(self._lib.mlv_clear_render_target_view(self._h, c))

    def clear_depth_stencil_view(self, depth: float):  # main.c:1204
        L.check(self._lib.mlv_clear_depth_stencil_view(self._h, float(depth)))

    def _bind(self, need_indices: bool):
        """Snapshot `graphics_pipeline` into the device, like the C shim does on every draw (host/malevich_compat.c).
        Only state that differs from what was last sent crosses the C-ABI: a draw that re-binds the same objects costs one
        ctypes call (the draw itself), which matters once the GPU side of a draw is tens of microseconds."""
        gp, lib, h, sent = self.graphics_pipeline, self._lib, self._h, self._sent
        v = int(gp.ia.primitive_topology)
        if sent.get("topology") != v:
            L.check(lib.mlv_ia_set_primitive_topology(h, v))
            sent["topology"] = v
        v = int(gp.ia.input_layout)
        if sent.get("layout") != v:
            L.check(lib.mlv_ia_set_input_layout(h, v))
            sent["layout"] = v
        if gp.ia.p_vertex_buffer is None:
            raise L.MalevichError(L.MLV_ERR_STATE, "ia.p_vertex_buffer is not set")
        hb = self._buffer(gp.ia.p_vertex_buffer, L.BUFFER_VERTEX)
        if sent.get("vb") is not hb:
            L.check(lib.mlv_ia_set_vertex_buffer(h, hb))
            sent["vb"] = hb
        if need_indices:
            ib = gp.ia.p_index_buffer
            if ib is None:
                raise L.MalevichError(L.MLV_ERR_STATE, "ia.p_index_buffer is not set")
            fmt = L.INDEX_U16 if ib.dtype == np.uint16 else L.INDEX_U32  # u16 = the reference's TODO (main.c:72)
            if sent.get("index_format") != fmt:
                L.check(lib.mlv_ia_set_index_format(h, fmt))
                sent["index_format"] = fmt
            hb = self._buffer(ib, L.BUFFER_INDEX)
            if sent.get("ib") is not hb:
                L.check(lib.mlv_ia_set_index_buffer(h, hb))
                sent["ib"] = hb
        if gp.vs.shader is None or gp.ps.shader is None:
            raise L.MalevichError(L.MLV_ERR_STATE, "vs.shader / ps.shader is not set")
        if gp.vs.output_register_count != 3:
            # the reference hard-codes three registers in interpolation and VS scatter (main.c:714-727,1124-1126)
            raise L.MalevichError(L.MLV_ERR_STATE, "vs.output_register_count must be 3")
        v = gp.vs.shader.vs_main
        if sent.get("vs") != v:
            L.check(lib.mlv_vs_set_shader(h, v))
            sent["vs"] = v
        v = gp.ps.shader.ps_main
        if sent.get("ps") != v:
            L.check(lib.mlv_ps_set_shader(h, v))
            sent["ps"] = v
        for slot, cb in enumerate(gp.vs.p_constant_buffers):
            if cb is not None:
                raw = np.ascontiguousarray(cb).tobytes()  # the contents, not the object: hosts rewrite the camera in place every frame
                if sent.get(("cb", slot)) != raw:
                    L.check(lib.mlv_vs_set_constant_buffer(h, slot, raw, len(raw)))
                    sent[("cb", slot)] = raw
        sent_vs, sent_ps = sent.setdefault("vs_srv", [None] * 16), sent.setdefault("ps_srv", [None] * 16)
        vs_srv, ps_srv = gp.vs.p_shader_resource_views, gp.ps.p_shader_resource_views
        for slot in range(16):
            t = vs_srv[slot]
            if t is not sent_vs[slot]:
                L.check(lib.mlv_vs_set_shader_resource(h, slot, self._texture(t) if t is not None else None))
                sent_vs[slot] = t
            t = ps_srv[slot]
            if t is not sent_ps[slot]:
                L.check(lib.mlv_ps_set_shader_resource(h, slot, self._texture(t) if t is not None else None))
                sent_ps[slot] = t
        v = gp.rs.viewport
        v = (v.top_left_x, v.top_left_y, v.width, v.height, v.min_depth, v.max_depth)
        if sent.get("viewport") != v:
            vp = L.Viewport(*v)
            L.check(lib.mlv_rs_set_viewport(h, C.byref(vp)))
            sent["viewport"] = v

    def draw_indexed(self, index_count: int, start_index_location: int = 0, base_vertex_location: int = 0):  # main.c:1219 (+ its TODO arguments)
        self._bind(True)
        if start_index_location or base_vertex_location:
            L.check(self._lib.mlv_draw_indexed_ex(self._h, int(index_count), int(start_index_location), int(base_vertex_location)))
        else:
            L.check(self._lib.mlv_draw_indexed(self._h, int(index_count)))

    def draw(self, vertex_count: int):
        self._bind(False)
        L.check(self._lib.mlv_draw(self._h, int(vertex_count)))

    # ---- command lists (D3D11 deferred-context pattern; one CUDA-graph launch per frame) ---------
    def begin_command_list(self):
        """Start recording: clears / draws / resolve issued from now on are captured instead of executed."""
        L.check(self._lib.mlv_begin_command_list(self._h))

    def finish_command_list(self) -> "CommandList":
        h = C.c_void_p()
        L.check(self._lib.mlv_finish_command_list(self._h, C.byref(h)))
        return CommandList(self, h)

    def record(self, fn) -> "CommandList":
        """Records whatever `fn()` issues on this device into a command list."""
        self.begin_command_list()
        try:
            fn()
        except Exception:
            try:
                self.finish_command_list().release()
            except Exception:
                pass
            raise
        return self.finish_command_list()

    def execute_command_list(self, cl: "CommandList"):
        L.check(self._lib.mlv_execute_command_list(self._h, cl._h))

    # ---- results ---------------------------------------------------------------------------------
    def present(self, want_depth: bool = True):
        """-> (colors uint32 [H, W], depths float32 [H, W] or None); replaces the GDI blit (main.c:286-357)."""
        colors = np.empty((self.height, self.width), dtype=np.uint32)
        depths = np.empty((self.height, self.width), dtype=np.float32) if want_depth else None
        L.check(self._lib.mlv_present_readback(self._h, colors.ctypes.data_as(C.c_void_p),
                                               depths.ctypes.data_as(C.c_void_p) if want_depth else None))
        return colors, depths

    def present_into(self, colors: np.ndarray, depths: Optional[np.ndarray] = None):
        L.check(self._lib.mlv_present_readback(self._h, colors.ctypes.data_as(C.c_void_p),
                                               depths.ctypes.data_as(C.c_void_p) if depths is not None else None))

    def present_owned_rows_async(self, frame_colors: np.ndarray):
        """Composite in host memory: this rank copies the rows it owns into the FULL frame `frame_colors` (pinned, shared by all
        ranks) over its own PCIe link; present_wait() completes it."""
        L.check(self._lib.mlv_present_owned_rows_async(self._h, frame_colors.ctypes.data_as(C.c_void_p)))

    def register_host_memory(self, address: int, nbytes: int):
        """Page-lock caller-owned host memory (e.g. the shared mapping of hostframe.SharedHostFrames) for asynchronous read-backs."""
        L.check(self._lib.mlv_register_host_memory(self._h, C.c_void_p(address), nbytes))

    def unregister_host_memory(self, address: int):
        L.check(self._lib.mlv_unregister_host_memory(self._h, C.c_void_p(address)))

    def present_async(self, colors: np.ndarray, depths: Optional[np.ndarray] = None):
        """Non-blocking present into (page-locked) arrays; `present_wait` or `finish` says when they are complete."""
        L.check(self._lib.mlv_present_readback_async(self._h, colors.ctypes.data_as(C.c_void_p),
                                                     depths.ctypes.data_as(C.c_void_p) if depths is not None else None))

    def present_wait(self):
        L.check(self._lib.mlv_present_wait(self._h))

    def finish(self):
        L.check(self._lib.mlv_finish(self._h))

    def stats(self) -> dict:
        s = L.Stats()
        L.check(self._lib.mlv_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def work_counters(self) -> dict:
        """What the kernels really processed since reset_stats (Stats count the reference's work, hidden or not)."""
        w = L.WorkCounters()
        L.check(self._lib.mlv_get_work_counters(self._h, C.byref(w)))
        return {"records_written": int(w.records_written), "pairs_listed": int(w.pairs_listed), "tiles_visited": int(w.tiles_visited)}

    def reset_stats(self):  # memset(&stats, 0) main.c:1268
        L.check(self._lib.mlv_reset_stats(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.mlv_get_stream(self._h) or 0)

    @property
    def kernel_launch_count(self) -> int:
        return int(self._lib.mlv_kernel_launch_count(self._h))

    def profile_begin(self):
        L.check(self._lib.mlv_profile_begin(self._h))

    def profile_end(self) -> dict:
        """-> {stage: (milliseconds, launches)} summed since profile_begin (CUDA events on the device stream)."""
        ms = (C.c_double * len(L.STAGE_NAMES))()
        n = (C.c_uint32 * len(L.STAGE_NAMES))()
        L.check(self._lib.mlv_profile_end(self._h, ms, n))
        return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(L.STAGE_NAMES)}

    def profile_events(self) -> list:
        """Launch-by-launch records of the last profiled region: [(stage name, start ms, duration ms)]."""
        n = C.c_uint32()
        L.check(self._lib.mlv_profile_read_events(self._h, None, 0, C.byref(n)))
        ev = (L.ProfileEvent * max(n.value, 1))()
        L.check(self._lib.mlv_profile_read_events(self._h, ev, n.value, C.byref(n)))
        return [(L.STAGE_NAMES[e.stage], float(e.start_ms), float(e.duration_ms)) for e in ev[:n.value]]

    def timeline_begin(self):
        """Launches issued or recorded from now on stamp %globaltimer into their slot (works inside command lists)."""
        L.check(self._lib.mlv_timeline_begin(self._h))

    def timeline_end(self):
        L.check(self._lib.mlv_timeline_end(self._h))

    def timeline_reset(self):
        L.check(self._lib.mlv_timeline_reset(self._h))

    def timeline_read(self) -> list:
        """[(stage name, draw, resident us, start us, end us)] per slot in issue order, relative to the earliest stamp."""
        n = C.c_uint32()
        L.check(self._lib.mlv_timeline_read(self._h, None, 0, C.byref(n)))
        ev = (L.TimelineEvent * max(n.value, 1))()
        L.check(self._lib.mlv_timeline_read(self._h, ev, n.value, C.byref(n)))
        return [(L.STAGE_NAMES[e.stage], int(e.draw), float(e.resident_us), float(e.start_us), float(e.end_us)) for e in ev[:n.value]]

    def resolve(self):
        L.check(self._lib.mlv_resolve(self._h))

    def resolved_color_ptr(self) -> int:
        return int(self._lib.mlv_resolved_color_device_ptr(self._h))

    def composite_layout(self):
        p, n = C.c_void_p(), C.c_size_t()
        L.check(self._lib.mlv_composite_layout(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def composite_pack(self):
        L.check(self._lib.mlv_composite_pack(self._h))

    def composite_unpack(self):
        L.check(self._lib.mlv_composite_unpack(self._h))

    # peer-memory compositing (fused resolve + all-gather over NVLink; see include/malevich_b200.h)
    def composite_peer_export(self) -> bytes:
        """This rank's mlv_peer_info as bytes, to be exchanged between the ranks by any transport."""
        info = L.PeerInfo()
        L.check(self._lib.mlv_composite_peer_export(self._h, C.byref(info)))
        return bytes(info)

    def composite_peer_attach(self, infos, same_process: bool = False):
        """`infos`: the exported bytes of every rank, in rank order."""
        arr = (L.PeerInfo * len(infos))(*[L.PeerInfo.from_buffer_copy(b) for b in infos])
        L.check(self._lib.mlv_composite_peer_attach(self._h, arr, 1 if same_process else 0))

    def composite_broadcast(self):
        L.check(self._lib.mlv_composite_broadcast(self._h))

    def composite_wait(self):
        L.check(self._lib.mlv_composite_wait(self._h))

    def composite_broadcast_async(self):
        """Broadcast + wait of this frame on the exchange stream; the device carries on with the next frame."""
        L.check(self._lib.mlv_composite_broadcast_async(self._h))

    def composite_join(self):
        """The device stream waits for the last composite_broadcast_async; resolved_color_ptr() is that frame."""
        L.check(self._lib.mlv_composite_join(self._h))

    # ---- debug read-back of the last draw ------------------------------------------------------
    def debug_vs_out(self) -> np.ndarray:
        n = C.c_uint32()
        L.check(self._lib.mlv_debug_read_vs_out(self._h, None, C.byref(n)))
        out = np.empty((n.value, 12), dtype=np.float32)
        L.check(self._lib.mlv_debug_read_vs_out(self._h, out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out

    def debug_triangles(self):
        n = C.c_uint32()
        L.check(self._lib.mlv_debug_read_triangles(self._h, None, None, C.byref(n)))
        tris = np.empty(n.value, dtype=REF_TRIANGLE_DTYPE)
        attrs = np.empty((n.value, 9, 4), dtype=np.float32)
        L.check(self._lib.mlv_debug_read_triangles(self._h, tris.ctypes.data_as(C.c_void_p), attrs.ctypes.data_as(C.c_void_p), C.byref(n)))
        return tris, attrs

    def debug_bins(self):
        npairs, nbins = C.c_uint32(), C.c_uint32()
        L.check(self._lib.mlv_debug_read_bins(self._h, None, C.byref(npairs), None, C.byref(nbins)))
        ids = np.empty(npairs.value, dtype=np.uint32)
        bins = np.empty(nbins.value, dtype=REF_COMPACTED_BIN_DTYPE)
        L.check(self._lib.mlv_debug_read_bins(self._h, ids.ctypes.data_as(C.c_void_p), C.byref(npairs), bins.ctypes.data_as(C.c_void_p), C.byref(nbins)))
        return ids, bins

    def debug_masks(self) -> np.ndarray:
        n = C.c_uint32()
        L.check(self._lib.mlv_debug_read_masks(self._h, None, C.byref(n)))
        infos = np.empty(n.value, dtype=REF_TILE_INFO_DTYPE)
        L.check(self._lib.mlv_debug_read_masks(self._h, infos.ctypes.data_as(C.c_void_p), C.byref(n)))
        return infos

    def debug_keys(self) -> np.ndarray:
        """keys[i] = device key of the reference's assembled triangle id i of the last draw (debug capture)."""
        n = C.c_uint32()
        L.check(self._lib.mlv_debug_read_keys(self._h, None, C.byref(n)))
        keys = np.empty(n.value, dtype=np.uint32)
        L.check(self._lib.mlv_debug_read_keys(self._h, keys.ctypes.data_as(C.c_void_p), C.byref(n)))
        return keys

    def bin_lists(self):
        """(keys, bins) of the last draw as the production path built them: Hi-Z-rejected pairs removed, arrival order."""
        npairs, nbins = C.c_uint32(), C.c_uint32()
        L.check(self._lib.mlv_read_bin_lists(self._h, None, C.byref(npairs), None, C.byref(nbins)))
        keys = np.empty(npairs.value, dtype=np.uint32)
        bins = np.empty(nbins.value, dtype=REF_COMPACTED_BIN_DTYPE)
        L.check(self._lib.mlv_read_bin_lists(self._h, keys.ctypes.data_as(C.c_void_p), C.byref(npairs), bins.ctypes.data_as(C.c_void_p), C.byref(nbins)))
        return keys, bins

    def debug_tile_min_depths(self) -> np.ndarray:
        out = np.empty((self.height // 8) * (self.width // 8), dtype=np.float32)
        L.check(self._lib.mlv_debug_read_tile_min_depths(self._h, out.ctypes.data_as(C.c_void_p)))
        return out


class CommandList:
    """A recorded frame (mlv_command_list). `set_constants` replaces the vertex-shader constant buffer of one draw or of
    all of them without re-recording -- the per-frame camera update of the reference's loop (main.c:1480-1562, 1595-1597)."""

    def __init__(self, dev: Device, handle):
        self._dev, self._h = dev, handle

    def execute(self):
        self._dev.execute_command_list(self)

    def set_constants(self, cb: np.ndarray, draw_index: int = L.ALL_DRAWS):
        raw = np.ascontiguousarray(cb, dtype=np.float32)
        L.check(self._dev._lib.mlv_command_list_set_constants(self._dev._h, self._h, int(draw_index), raw.ctypes.data_as(C.c_void_p), raw.nbytes))

    @property
    def info(self) -> dict:
        d, k = C.c_uint32(), C.c_uint64()
        L.check(self._dev._lib.mlv_command_list_info(self._h, C.byref(d), C.byref(k)))
        return {"draws": int(d.value), "kernel_launches": int(k.value)}

    def release(self):
        if self._h and self._dev._h:
            self._dev._lib.mlv_release_command_list(self._dev._h, self._h)
        self._h = None
