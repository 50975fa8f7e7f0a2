"""CPU tests (`-m "not gpu"`): the oracle against every known answer the reference offers, the host
logic, and the C-ABI library's exported symbols. No compute call needs a GPU here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import cases
from conftest import ROOT, have_ref
from malevich_b200 import _lib as L
from malevich_b200 import assets, camera, scenes

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
needs_ref = pytest.mark.skipif(not have_ref(320, 200), reason="oracle/_ref not built (needs /root/reference at build time)")


def _oracle(w, h, threads=1):
    from oracle.ref_oracle import RefOracle
    return RefOracle(w, h, threads=threads)


# ---- oracle pinned against the reference's own known answers ---------------
@needs_ref
def test_oracle_suprematism_known_pixel_histogram():
    """SUPREMATISM (embedded in main.c:232-254): exactly three colours with the surveyed pixel counts."""
    from oracle.ref_oracle import fnv64_words
    sc = scenes.suprematism(1200, 720)
    o = _oracle(1200, 720)
    o.render(sc)
    col = o.colors()
    vals, counts = np.unique(col, return_counts=True)
    got = {hex(int(v)): int(c) for v, c in zip(vals, counts)}
    assert got == GOLDEN["_reference_known_answers"]["suprematism_1200x720_pixel_histogram"]
    assert sorted(np.unique(o.depths()).tolist()) == [0.25, 0.5, 0.75]
    assert fnv64_words(col) == GOLDEN["sup_1200x720"]["color_fnv"]
    assert fnv64_words(o.depths()) == GOLDEN["sup_1200x720"]["depth_fnv"]


@needs_ref
def test_oracle_embedded_scene_matches_reference_arrays():
    """assets.suprematist_scene()/fullscreen_quad() restate main.c:232-270; compare with the arrays compiled from it."""
    o = _oracle(320, 200)
    n = ctypes.c_uint32()
    for getter_vb, getter_ib, mine in ((o.lib.ref_suprematist_vb, o.lib.ref_suprematist_ib, assets.suprematist_scene()),
                                       (o.lib.ref_fullscreen_vb, o.lib.ref_fullscreen_ib, assets.fullscreen_quad())):
        p = getter_vb(ctypes.byref(n))
        vb = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), (n.value // 32, 8))
        assert np.array_equal(vb.view(np.uint32), mine[0].view(np.uint32))
        p = getter_ib(ctypes.byref(n))
        ib = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), (n.value,))
        assert np.array_equal(ib, mine[1])


@needs_ref
def test_oracle_ftm_screenshot_stats():
    """screenshot.png overlay (FTM, 1200x720): 143 808 vertices, 47 936 triangles in, 23 606 assembled, 27 481 active
    bins. The gcc build of the reference gives 23 605 assembled (+-1 across compilers)."""
    ref_cb = _oracle(1200, 720).camera(*scenes.FTM_SCREENSHOT_POSE)
    sc = scenes.ftm(1200, 720, cb=ref_cb)
    o = _oracle(1200, 720)
    o.render(sc)
    st = o.stats()
    known = GOLDEN["_reference_known_answers"]["screenshot_png_overlay_ftm_1200x720"]
    assert st["vertex_count"] == known["vertex_count"]
    assert st["input_triangle_count"] == known["input_triangle_count"]
    assert st["active_bin_count"] == known["active_bin_count"]
    assert abs(st["assembled_triangle_count"] - known["assembled_triangle_count"]) <= 1
    assert abs(st["total_triangle_count_in_bins"] / st["active_bin_count"] - known["avg_triangles_per_bin"]) < 1e-3
    assert st == GOLDEN["ftm_screenshot_refcam_1200x720"]["stats"]


@needs_ref
@pytest.mark.parametrize("name", sorted(cases.SMALL))
def test_oracle_reproduces_committed_golden_frames(name):
    from oracle.ref_oracle import fnv64_words
    frames = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))
    sc = cases.SMALL[name]()
    o = _oracle(sc.width, sc.height)
    o.render(sc)
    assert np.array_equal(o.depths().view(np.uint32), frames[name + "/depths"].view(np.uint32))
    assert np.array_equal(o.colors(), frames[name + "/colors"])
    assert o.stats() == GOLDEN[name]["stats"]
    assert fnv64_words(o.depths()) == GOLDEN[name]["depth_fnv"]


@needs_ref
def test_oracle_thread_count_only_affects_z_ties():
    """Assembled order is schedule dependent with >1 OMP thread (main.c:874); depth is not."""
    sc = scenes.toon(320, 200)
    o1 = _oracle(320, 200, threads=1)
    o1.render(sc)
    d1, s1 = o1.depths(), o1.stats()
    o8 = _oracle(320, 200, threads=4)
    o8.render(sc)
    assert np.array_equal(d1.view(np.uint32), o8.depths().view(np.uint32))
    assert s1 == o8.stats()
    o8.set_threads(1)


# ---- host-side restatements vs the reference ------------------------------------------------------
@needs_ref
def test_srgb_to_linear_lut_matches_reference_load_path():
    """assets.srgb_texture_to_linear restates main.c:546-558; all 256 byte values, all four channels."""
    o = _oracle(320, 200)
    allb = (np.arange(256, dtype=np.uint32) * np.uint32(0x01010101)).reshape(16, 16)
    assert np.array_equal(o.texture_srgb_to_linear(allb), assets.srgb_texture_to_linear(allb))
    t = assets.standin_texture_srgb(3, 128)
    assert np.array_equal(o.texture_srgb_to_linear(t), assets.srgb_texture_to_linear(t))


@needs_ref
def test_camera_matches_reference_camera_to_ulps():
    """camera.py restates update() (main.c:1480-1562) with a generic inverse instead of the reference's closed form
    (math.h:282-320). For the pose of BASELINE configs 1, 3, 4 and 5 (no yaw, no pitch) the 48 floats are the reference's
    value for value (a zero may carry the other sign); with yaw and pitch (FTM, config 2) a handful of elements differ in
    the last place or two, and elements that are rounding residue of a cancellation (|x| < 1e-8) differ freely."""
    for w, h in ((3840, 2160), (1280, 720), (1920, 1080)):
        ref = np.asarray(_oracle(w, h).camera((3.5, 1.0, 1.0), 0.0, 0.0), dtype=np.float32)
        assert np.array_equal(ref, camera.per_frame_cb(w, h, (3.5, 1.0, 1.0), 0.0, 0.0))
    for w, h, pose in ((1200, 720, scenes.FTM_SCREENSHOT_POSE), (1920, 1080, ((1.0, -2.0, 0.5), 1.2, -0.3))):
        ref = np.asarray(_oracle(w, h).camera(*pose), dtype=np.float32).ravel()
        mine = camera.per_frame_cb(w, h, *pose).ravel()
        differing = ref != mine
        assert differing.sum() <= 8
        big = np.abs(ref) >= 1e-8
        assert np.all(np.abs(ref[big] - mine[big]) <= 4 * np.spacing(np.abs(ref[big])))  # <= 4 ulp
        assert np.all(np.abs(ref[~big] - mine[~big]) <= 1e-8)
        assert np.allclose(ref, mine, rtol=2e-5, atol=2e-5)


@needs_ref
def test_rsqrt_table_reproduces_host_vrsqrtps():
    """the committed 2x1024 table reproduces vrsqrtps bit-exactly on an Intel host."""
    o = _oracle(320, 200)
    txt = open(os.path.join(ROOT, "malevich_b200", "csrc", "rsqrt_lut.inc")).read()
    lut = np.array([int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", txt)], dtype=np.uint32)
    assert lut.shape == (2048,)
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(1e-30, 1e30, 2000), rng.uniform(0.0, 4.0, 2000), [1.0, 2.0, 4.0, 0.5, 3.9999]]).astype(np.float32)
    mismatches = 0
    for x in xs:
        u = int(np.float32(x).view(np.uint32))
        e = (u >> 23) & 0xFF
        parity = (e - 127) & 1
        half = (e - 127 - parity) // 2
        expect = (int(lut[parity * 1024 + ((u >> 13) & 1023)]) - (half << 23)) & 0xFFFFFFFF
        got = int(np.float32(o.rsqrt(float(x))).view(np.uint32))
        mismatches += expect != got
    if mismatches:
        pytest.skip(f"host vrsqrtps differs from the committed Intel table on {mismatches} inputs (non-Intel host?)")


def test_octrn_reader_and_assets():
    vb, ib = assets.load_mesh("toon_house_mesh")
    assert vb.shape == (22056, 8) and ib.shape == (22056,) and np.array_equal(ib, np.arange(22056, dtype=np.uint32))
    total = sum(assets.load_mesh(m)[1].shape[0] for m in scenes.FTM_MESHES)
    assert total == 143808  # screenshot.png: vertex count 143 808
    irr = assets.load_irradiance()
    assert irr.shape == (128, 256, 4) and irr.dtype == np.float32 and np.all(irr[..., 3] == 1.0)
    with pytest.raises(assets.OctrnError):
        assets.read_octrn_image(os.path.join(assets.ASSET_DIR, "toon_sky_mesh.octrn"))


def test_standins_are_deterministic_and_well_formed():
    a, b = assets.standin_texture(2, 64), assets.standin_texture(2, 64)
    assert np.array_equal(a, b) and a.dtype == np.uint32
    for build in (lambda: assets.uv_sphere(n_lat=8, n_lon=16), lambda: assets.torus_knot(n_u=32, n_v=8), lambda: assets.synthetic_grid_layer(1, 320, 200, 12, 6)):
        vb, ib = build()
        assert vb.dtype == np.float32 and vb.shape[1] == 8 and ib.dtype == np.uint32
        assert ib.max() < vb.shape[0] and ib.shape[0] % 3 == 0
    vb, ib = assets.synthetic_grid_layer(0, 3840, 2160)
    assert ib.shape[0] == 3750000 and ib.shape[0] % 8 == 0 and vb.shape[0] == 1251 * 501
    assert assets.pad_indices_to_8(np.arange(6, dtype=np.uint32)).shape[0] == 24


# ---- the C-ABI library ------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "malevich_b200.h")).read()
    declared = sorted(set(re.findall(r"MLV_API\s+[\w \*]+?\b(mlv_\w+)\s*\(", header)))
    assert declared == sorted(L.EXPORTED_SYMBOLS), "header and binding disagree"
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by {L.LIB_PATH}"


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The header is the contract: compile it with the C compiler and compare sizeof / offsetof of every plain-data struct with
    the ctypes mirrors in malevich_b200/_lib.py and the numpy record dtypes in device.py (an ABI drift would otherwise only
    show up as garbage on the GPU box)."""
    import ctypes as C
    import shutil
    import subprocess
    from malevich_b200 import _lib as L, device as D
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"mlv_viewport": L.Viewport, "mlv_stats": L.Stats, "mlv_device_desc": L.DeviceDesc, "mlv_peer_info": L.PeerInfo,
               "mlv_work_counters": L.WorkCounters, "mlv_profile_event": L.ProfileEvent, "mlv_timeline_event": L.TimelineEvent}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "malevich_b200.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    for cname in ("mlv_ref_triangle", "mlv_ref_compacted_bin", "mlv_ref_tile_info"):
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
    lines.append('printf("MLV_PS_COUNT %d\\nMLV_VS_COUNT %d\\nMLV_STAGE_COUNT %d\\nMLV_MAX_PEERS %d\\n", MLV_PS_COUNT, MLV_VS_COUNT, MLV_STAGE_COUNT, MLV_MAX_PEERS);')
    lines += ["return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)  # the header is plain C
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"
    assert int(got["mlv_ref_triangle"]) == D.REF_TRIANGLE_DTYPE.itemsize == 80        # Triangle main.c:134-139
    assert int(got["mlv_ref_compacted_bin"]) == D.REF_COMPACTED_BIN_DTYPE.itemsize == 12  # CompactedBin main.c:146-150
    assert int(got["mlv_ref_tile_info"]) == D.REF_TILE_INFO_DTYPE.itemsize == 16        # TileInfo main.c:159-162
    assert int(got["MLV_STAGE_COUNT"]) == len(L.STAGE_NAMES)
    assert int(got["MLV_PS_COUNT"]) == 1 + max(L.PS_PASSTHROUGH, L.PS_BASIC, L.PS_ENV_LIGHTING, L.PS_BASIC_TRILINEAR)
    assert int(got["MLV_VS_COUNT"]) == 1 + max(L.VS_PASSTHROUGH, L.VS_BASIC, L.VS_VERTEX_LIGHTING, L.VS_FULLSCREEN)


def test_c_abi_argument_errors_without_gpu():
    lib = L.load()
    h = ctypes.c_void_p()
    bad = L.DeviceDesc(width=1201, height=720, cuda_device=-1)
    assert lib.mlv_create_device(ctypes.byref(bad), ctypes.byref(h)) == L.MLV_ERR_INVALID_ARGUMENT
    assert b"multiples of 8" in lib.mlv_last_error_string()
    assert lib.mlv_draw_indexed(None, 24) == L.MLV_ERR_INVALID_ARGUMENT
    buf = (ctypes.c_uint32 * 16)()
    for call in (lambda: lib.mlv_present_owned_rows_async(None, buf), lambda: lib.mlv_register_host_memory(None, buf, 64), lambda: lib.mlv_unregister_host_memory(None, buf),
                 lambda: lib.mlv_present_readback_async(None, buf, None), lambda: lib.mlv_execute_command_list(None, None)):
        assert call() == L.MLV_ERR_INVALID_ARGUMENT
    import torch
    if not torch.cuda.is_available():
        ok = L.DeviceDesc(width=320, height=200, cuda_device=-1)
        assert lib.mlv_create_device(ctypes.byref(ok), ctypes.byref(h)) == L.MLV_ERR_CUDA  # fails loudly: no CPU fallback
        assert b"no CPU fallback" in lib.mlv_last_error_string()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "malevich_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/harness.c", ""), f
