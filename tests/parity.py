"""Shared parity checks: GPU path (through the C-ABI) vs the reference oracle, stage by stage.

Bars (BASELINE.json north_star): coverage masks, depth buffer, assembled triangles, per-tile triangle
order and Stats are BIT-EXACT; shaded colour is within 1/255 per channel on >= 99.9 % of pixels and
never off by more than 2/255.
"""
from __future__ import annotations

import numpy as np

from malevich_b200 import scenes as S

COLOR_TOL_FRACTION = 0.999  # >= 99.9 % of pixels within 1/255 per channel
COLOR_TOL_MAX = 2           # none off by more than 2/255


def host_vrsqrtps_matches_table(oracle) -> bool:
    """The live oracle executes the HOST's vrsqrtps (math.h:278), a hardware-defined approximation; the
    GPU reproduces the Intel table committed in csrc/rsqrt_lut.inc. On a host whose instruction differs (non-Intel)
    the live oracle is not the canonical reference for anything that goes through a normal; the committed golden frames
    (generated on an Intel host) remain valid."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "malevich_b200", "csrc", "rsqrt_lut.inc")).read()
    lut = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", txt)]
    for parity_bit in (0, 1):
        for i in range(0, 1024, 7):
            x = np.array([((127 + parity_bit) << 23) | (i << 13)], dtype=np.uint32).view(np.float32)[0]
            got = int(np.float32(oracle.rsqrt(float(x))).view(np.uint32))
            if got != lut[parity_bit * 1024 + i]:
                return False
    return True


def channel_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """max over the four bytes of |a - b| per pixel."""
    d = np.zeros(a.shape, dtype=np.int32)
    for sh in (0, 8, 16, 24):
        ca = ((a >> sh) & 0xFF).astype(np.int32)
        cb = ((b >> sh) & 0xFF).astype(np.int32)
        d = np.maximum(d, np.abs(ca - cb))
    return d


def compare_frames(gpu_colors, gpu_depths, ref_colors, ref_depths) -> dict:
    d = channel_diff(gpu_colors, ref_colors)
    n = d.size
    return {
        "depth_bit_exact": bool(np.array_equal(gpu_depths.view(np.uint32), ref_depths.view(np.uint32))),
        "depth_mismatch_pixels": int(np.count_nonzero(gpu_depths.view(np.uint32) != ref_depths.view(np.uint32))),
        "color_exact_fraction": float(np.count_nonzero(d == 0)) / n,
        "color_within1_fraction": float(np.count_nonzero(d <= 1)) / n,
        "color_max_diff": int(d.max()) if n else 0,
        "color_over2_pixels": int(np.count_nonzero(d > COLOR_TOL_MAX)),
    }


def assert_frames_match(gpu_colors, gpu_depths, ref_colors, ref_depths, what=""):
    r = compare_frames(gpu_colors, gpu_depths, ref_colors, ref_depths)
    assert r["depth_bit_exact"], f"{what}: depth differs on {r['depth_mismatch_pixels']} pixels"
    assert r["color_within1_fraction"] >= COLOR_TOL_FRACTION, f"{what}: only {r['color_within1_fraction']:.6f} of pixels within 1/255"
    assert r["color_max_diff"] <= COLOR_TOL_MAX, f"{what}: colour off by {r['color_max_diff']}/255 on {r['color_over2_pixels']} pixels"
    return r


def compare_staged_draw(dev, oracle, vs_id=1) -> dict:
    """Compares every intermediate of the LAST draw (GPU device must be in debug-capture mode, oracle
    draw must have been staged). Returns a dict of booleans/counters; everything must be True/0."""
    out = {}
    # vertex shader output: r0, r1 and r2.x are defined by basic_vs / vertex_lighting_vs; passthrough_vs and
    # fullscreen_vs leave UV (r1.w, r2.x) unwritten; r2.yzw is never written (uninitialised stack in the
    # reference, f256 vertex_output[12] main.c:711) -- undefined lanes are excluded from the comparison
    ndef = 9 if vs_id in (1, 2) else 7
    g_vs, r_vs = dev.debug_vs_out(), oracle.staged_vs_out()
    out["vs_count"] = g_vs.shape[0] == r_vs.shape[0]
    if out["vs_count"] and g_vs.shape[0]:
        with np.errstate(all="ignore"):
            out["vs_out_max_abs_diff"] = float(np.nanmax(np.abs(g_vs[:, :ndef].astype(np.float64) - r_vs[:, :ndef].astype(np.float64))))
        if vs_id == 2:  # colour lanes 4..6 to a few ulp (see below), everything else bit-exact
            exact = [0, 1, 2, 3, 7, 8]
            out["vs_out_bit_exact"] = bool(np.array_equal(g_vs[:, exact].view(np.uint32), r_vs[:, exact].view(np.uint32)) and out["vs_out_max_abs_diff"] <= 4e-6)
        else:
            out["vs_out_bit_exact"] = bool(np.array_equal(g_vs[:, :ndef].view(np.uint32), r_vs[:, :ndef].view(np.uint32)))
    g_tris, g_attrs = dev.debug_triangles()
    r_tris, r_attrs = oracle.staged_triangles()
    out["tri_count"] = (len(g_tris), len(r_tris))
    if len(g_tris) == len(r_tris) and len(g_tris):
        for f in ("min_bounds", "max_bounds", "edges"):
            out[f"tri_{f}"] = bool(np.array_equal(g_tris[f], r_tris[f]))
        for f in ("reciprocal_ws", "one_over_area", "max_depth"):
            out[f"tri_{f}"] = bool(np.array_equal(g_tris[f].view(np.uint32), r_tris[f].view(np.uint32)))
        # attributes: v{0,1,2} x {r0, r1, r2}; compare r0, r1 fully and r2.x
        ga = g_attrs.reshape(-1, 3, 3, 4).view(np.uint32)
        ra = r_attrs.reshape(-1, 3, 3, 4).view(np.uint32)
        # bit-exact; a NaN (Inf * 0 on a vertex with infinite coordinates) compares equal to a NaN: x86 produces the negative
        # default NaN, the GPU the positive one, and nothing downstream tells them apart (snap() maps both to 0x80000000)
        gf0, rf0 = g_attrs.reshape(-1, 3, 3, 4)[:, :, 0], r_attrs.reshape(-1, 3, 3, 4)[:, :, 0]
        out["attr_r0"] = bool(np.all((ga[:, :, 0] == ra[:, :, 0]) | (np.isnan(gf0) & np.isnan(rf0))))
        if vs_id == 2:
            # vertex_lighting_vs computes COLOR with acos/exp/pow (Intel SVML in the reference, glibc in the oracle,
            # CUDA libdevice here): "parity unpinned" arithmetic, compared to a few ulp; UV stays bit-exact
            gf, rf = g_attrs.reshape(-1, 3, 3, 4), r_attrs.reshape(-1, 3, 3, 4)
            out["attr_r1_color_max_abs_diff"] = float(np.abs(gf[:, :, 1, :3].astype(np.float64) - rf[:, :, 1, :3]).max())
            out["attr_r1"] = bool(out["attr_r1_color_max_abs_diff"] <= 4e-6 and np.array_equal(ga[:, :, 1, 3], ra[:, :, 1, 3]))
            out["attr_r2x"] = bool(np.array_equal(ga[:, :, 2, 0], ra[:, :, 2, 0]))
        elif ndef == 9:
            out["attr_r1"] = bool(np.array_equal(ga[:, :, 1], ra[:, :, 1]))
            out["attr_r2x"] = bool(np.array_equal(ga[:, :, 2, 0], ra[:, :, 2, 0]))
        else:
            out["attr_r1"] = bool(np.array_equal(ga[:, :, 1, :3], ra[:, :, 1, :3]))
    g_ids, g_bins = dev.debug_bins()
    r_ids, r_bins = oracle.staged_bins()
    out["pair_count"] = (len(g_ids), len(r_ids))
    out["bin_count"] = (len(g_bins), len(r_bins))
    if len(g_bins) == len(r_bins):
        out["compacted_bins"] = bool(np.array_equal(g_bins, r_bins))
    if len(g_ids) == len(r_ids):
        out["triangle_id_order"] = bool(np.array_equal(g_ids, r_ids))
    g_infos, r_infos = dev.debug_masks(), oracle.staged_tile_infos()
    if len(g_infos) == len(r_infos):
        out["tile_info_ids"] = bool(np.array_equal(g_infos["triangle_id"], r_infos["triangle_id"]))
        out["coverage_masks"] = bool(np.array_equal(g_infos["fragment_mask"], r_infos["fragment_mask"]))
    g_tm, r_tm = dev.debug_tile_min_depths(), oracle.tile_min_depths()
    out["tile_min_depths"] = bool(np.array_equal(g_tm, r_tm))  # float equality (+0 == -0)
    return out


def staged_ok(res: dict) -> bool:
    for k, v in res.items():
        if isinstance(v, bool) and not v:
            return False
        if isinstance(v, tuple) and v[0] != v[1]:
            return False
    return True


def render_both_staged(dev, oracle, scene, clear=True):
    """Draw-by-draw render on both sides with a staged comparison after each draw."""
    import malevich_b200._lib as L
    results = []
    gp = dev.graphics_pipeline
    if clear:
        dev.clear_render_target_view(S.CLEAR_COLOR)
        dev.clear_depth_stencil_view(S.CLEAR_DEPTH)
    oracle.begin_frame(scene.per_frame_cb, S.CLEAR_COLOR if clear else None, S.CLEAR_DEPTH)
    dev.reset_stats()
    gp.ia.primitive_topology = L.PRIMITIVE_TOPOLOGY_TRIANGLELIST
    gp.rs.viewport.width, gp.rs.viewport.height = float(scene.width), float(scene.height)
    gp.rs.viewport.top_left_x = gp.rs.viewport.top_left_y = 0.0
    gp.rs.viewport.min_depth, gp.rs.viewport.max_depth = 0.0, 1.0
    gp.vs.p_constant_buffers[0] = scene.per_frame_cb
    for o in scene.objects:
        gp.ia.input_layout = o.vertex_shader.in_vertex_size // 8
        gp.vs.output_register_count = o.vertex_shader.out_vertex_size // 128
        gp.vs.shader, gp.ps.shader = o.vertex_shader, o.pixel_shader
        gp.ia.p_index_buffer, gp.ia.p_vertex_buffer = o.index_buffer, o.vertex_buffer
        gp.vs.p_shader_resource_views[0] = o.texture
        gp.ps.p_shader_resource_views[0] = o.texture
        dev.draw_indexed(o.index_count)
        oracle.draw(o.vertex_buffer, o.index_buffer, o.vertex_shader.vs_main, o.pixel_shader.ps_main, o.texture, staged=True)
        results.append((o.name, compare_staged_draw(dev, oracle, o.vertex_shader.vs_main)))
    return results
