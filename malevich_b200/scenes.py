"""Scene table + `render()` -- the caller of the hot path (reference main.c:214-220, 1265-1299, 1317-1420).

The five BASELINE.json configs are restated here as concrete inputs: real meshes and
the real irradiance panorama where the reference checkout ships them, deterministic stand-ins
(malevich_b200.assets) for the 12 assets it does not.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib as L
from . import assets, camera
from .device import (Device, PixelShader, Texture2D, VertexShader, VECTOR_WIDTH, basic_ps, basic_vs, env_lighting_ps,
                     fullscreen_vs, passthrough_ps, passthrough_vs, vertex_lighting_vs)

# render() main.c:1271: { (f32)227/255, (f32)223/255, (f32)216/255, 0.f }
CLEAR_COLOR = tuple(float(np.float32(v) / np.float32(255)) for v in (227, 223, 216)) + (0.0,)
CLEAR_DEPTH = 0.0  # reversed Z: far = 0 (main.c:1273)


@dataclass
class SceneObject:  # one entry of Scene.a_meshes/a_textures/a_vertex_shaders/a_pixel_shaders (main.c:214-220)
    vertex_buffer: np.ndarray
    index_buffer: np.ndarray
    vertex_shader: VertexShader
    pixel_shader: PixelShader
    texture: Optional[Texture2D] = None
    name: str = ""

    @property
    def index_count(self) -> int:
        return int(self.index_buffer.shape[0])


@dataclass
class Scene:
    name: str
    width: int
    height: int
    objects: List[SceneObject]
    per_frame_cb: np.ndarray = field(default_factory=lambda: np.zeros((3, 4, 4), np.float32))
    camera_pose: tuple = ((3.5, 1.0, 1.0), 0.0, 0.0)  # pos, yaw, pitch (main.c:1423-1425)

    @property
    def input_triangles(self) -> int:
        return sum(o.index_count // 3 for o in self.objects)


def render(dev: Device, scene: Scene, clear: bool = True) -> None:
    """render() main.c:1265-1299, statement for statement."""
    if clear:
        dev.clear_render_target_view(CLEAR_COLOR)
        dev.clear_depth_stencil_view(CLEAR_DEPTH)
    gp = dev.graphics_pipeline
    gp.ia.primitive_topology = L.PRIMITIVE_TOPOLOGY_TRIANGLELIST
    gp.rs.viewport.top_left_x, gp.rs.viewport.top_left_y = 0.0, 0.0
    gp.rs.viewport.width, gp.rs.viewport.height = float(scene.width), float(scene.height)
    gp.rs.viewport.min_depth, gp.rs.viewport.max_depth = 0.0, 1.0
    gp.vs.p_constant_buffers[0] = scene.per_frame_cb
    for obj in scene.objects:
        gp.ia.input_layout = obj.vertex_shader.in_vertex_size // VECTOR_WIDTH
        gp.vs.output_register_count = obj.vertex_shader.out_vertex_size // (16 * VECTOR_WIDTH)
        gp.vs.shader = obj.vertex_shader
        gp.ps.shader = obj.pixel_shader
        gp.ia.p_index_buffer = obj.index_buffer
        gp.ia.p_vertex_buffer = obj.vertex_buffer
        gp.vs.p_shader_resource_views[0] = obj.texture
        gp.ps.p_shader_resource_views[0] = obj.texture
        dev.draw_indexed(obj.index_count)


def upload(dev: Device, scene: Scene) -> None:
    for obj in scene.objects:
        dev.upload(obj.vertex_buffer, obj.index_buffer)
        if obj.texture is not None:
            dev.upload(obj.texture)


# ---------------------------------------------------------------------------------------------
FTM_SCREENSHOT_POSE = ((-8.8964, 5.61089, 0.9198), -2.92499, 0.05)  # camera shown in the reference's screenshot.png
FTM_MESHES = ["ftm_piedras_mesh", "ftm_madera_mesh", "ftm_leaves_mesh", "ftm_dec_mesh", "ftm_roof_mesh", "ftm_ground_mesh", "ftm_sky_mesh"]  # init() order main.c:1319-1359

_cache = {}


def _tex(seed: int) -> Texture2D:
    key = ("tex", seed)
    if key not in _cache:
        _cache[key] = Texture2D(assets.standin_texture(seed))
    return _cache[key]


def _irradiance() -> Texture2D:
    if "irr" not in _cache:
        _cache["irr"] = Texture2D(assets.load_irradiance())
    return _cache["irr"]


def _radiance() -> Texture2D:
    if "rad" not in _cache:
        _cache["rad"] = Texture2D(assets.standin_radiance(assets.load_irradiance()))
    return _cache["rad"]


def _scene(name, width, height, objects, pose, cb) -> Scene:
    if cb is None:
        cb = camera.per_frame_cb(width, height, pose[0], pose[1], pose[2])
    return Scene(name, width, height, objects, np.ascontiguousarray(cb, dtype=np.float32).reshape(3, 4, 4), pose)


def suprematism(width=1200, height=720, cb=None) -> Scene:
    """SUPREMATISM (main.c:232-254,1381-1389): asset-free known-answer scene."""
    vb, ib = assets.suprematist_scene()
    return _scene("suprematism", width, height, [SceneObject(vb, ib, passthrough_vs, passthrough_ps, None, "suprematist")], ((3.5, 1.0, 1.0), 0.0, 0.0), cb)


def toon(width=1280, height=720, cb=None) -> Scene:
    """Config 1: TOON scene (main.c:1364-1379), basic_vs/basic_ps, default camera."""
    objs = []
    for i, m in enumerate(["toon_house_mesh", "toon_sky_mesh"]):
        vb, ib = assets.load_mesh(m)
        objs.append(SceneObject(vb, ib, basic_vs, basic_ps, _tex(i), m))
    return _scene("toon", width, height, objs, ((3.5, 1.0, 1.0), 0.0, 0.0), cb)


def ftm(width=1920, height=1080, cb=None) -> Scene:
    """Config 2: FTM scene (main.c:1317-1362), 7 draws, screenshot camera."""
    objs = []
    for i, m in enumerate(FTM_MESHES):
        vb, ib = assets.load_mesh(m)
        objs.append(SceneObject(vb, ib, basic_vs, basic_ps, _tex(i), m))
    return _scene("ftm", width, height, objs, FTM_SCREENSHOT_POSE, cb)


def emily(width=1920, height=1080, cb=None, n_lat=256, n_lon=512) -> Scene:
    """Config 3: EMILY scene (main.c:1391-1409). emily_head_mesh is missing -> UV sphere; draw 2 is the
    reference's fullscreen quad with fullscreen_vs + env_lighting_ps on the (stand-in) radiance panorama."""
    vb, ib = assets.uv_sphere(n_lat=n_lat, n_lon=n_lon)
    fvb, fib = assets.fullscreen_quad()
    objs = [SceneObject(vb, ib, basic_vs, env_lighting_ps, _irradiance(), "emily_standin_sphere"),
            SceneObject(fvb, fib, fullscreen_vs, env_lighting_ps, _radiance(), "fullscreen_quad")]
    return _scene("emily", width, height, objs, ((3.5, 1.0, 1.0), 0.0, 0.0), cb)


def locomotive(width=3840, height=2160, cb=None, n_u=4096, n_v=128) -> Scene:
    """Config 4: LOCOMOTIVE scene (main.c:1411-1420). locomotive_mesh is missing -> torus knot."""
    vb, ib = assets.torus_knot(n_u=n_u, n_v=n_v)
    objs = [SceneObject(vb, ib, vertex_lighting_vs, passthrough_ps, _irradiance(), "locomotive_standin_torus_knot")]
    return _scene("locomotive", width, height, objs, ((3.5, 1.0, 1.0), 0.0, 0.0), cb)


def synthetic(width=3840, height=2160, cb=None, layers=8, nx=1250, ny=500) -> Scene:
    """Config 5: 8 draws x (nx*ny*2) triangles = 10 M at the defaults, front-to-back."""
    objs = []
    for layer in range(layers):
        vb, ib = assets.synthetic_grid_layer(layer, width, height, nx, ny)
        objs.append(SceneObject(vb, ib, basic_vs, basic_ps, _tex(layer), f"grid_layer_{layer}"))
    return _scene("synthetic", width, height, objs, ((3.5, 1.0, 1.0), 0.0, 0.0), cb)


CONFIGS = {1: toon, 2: ftm, 3: emily, 4: locomotive, 5: synthetic}
