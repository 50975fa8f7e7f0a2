"""The reference-language host: host/render_host.c (the reference's render() + scene table in C) linked against the
C-ABI through host/malevich_compat.c (the reference's three entry points and global pipeline state)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import cases
import parity
from conftest import ROOT, have_ref

HOST_DIR = os.path.join(ROOT, "host")
HOST_BIN = os.path.join(HOST_DIR, "render_host")


def write_scene(path, sc):
    with open(path, "wb") as f:
        f.write(b"MLVSCENE")
        f.write(struct.pack("<IIII", sc.width, sc.height, len(sc.objects), 0))
        f.write(np.ascontiguousarray(sc.per_frame_cb, dtype=np.float32).tobytes())
        for o in sc.objects:
            tex = o.texture
            kind = 0 if tex is None else (1 if tex.p_data.dtype == np.uint32 else 2)
            f.write(struct.pack("<IIIIIIII", o.vertex_shader.vs_main, o.pixel_shader.ps_main, o.vertex_buffer.shape[0], o.index_count, kind,
                                tex.width if tex is not None else 0, tex.height if tex is not None else 0, 0))
            f.write(np.ascontiguousarray(o.vertex_buffer, dtype=np.float32).tobytes())
            f.write(np.ascontiguousarray(o.index_buffer, dtype=np.uint32).tobytes())
            if tex is not None:
                f.write(tex.p_data.tobytes())


def read_frame(path):
    blob = open(path, "rb").read()
    w, h = struct.unpack_from("<II", blob, 0)
    n = w * h
    col = np.frombuffer(blob, dtype=np.uint32, count=n, offset=8).reshape(h, w)
    dep = np.frombuffer(blob, dtype=np.float32, count=n, offset=8 + 4 * n).reshape(h, w)
    st = struct.unpack_from("<fIIIII", blob, 8 + 8 * n)
    keys = ["vertex_count", "input_triangle_count", "assembled_triangle_count", "active_bin_count", "total_triangle_count_in_bins"]
    return col, dep, dict(zip(keys, st[1:]))


def test_c_host_builds_and_links_the_c_abi():
    subprocess.run(["make", "-C", HOST_DIR], check=True, capture_output=True)
    assert os.path.exists(HOST_BIN)
    syms = subprocess.run(["nm", "-D", "--undefined-only", HOST_BIN], check=True, capture_output=True, text=True).stdout
    for name in ("mlv_create_device", "mlv_draw_indexed", "mlv_clear_render_target_view", "mlv_clear_depth_stencil_view", "mlv_present_readback"):
        assert name in syms


def _fnv64(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in np.frombuffer(data, dtype=np.uint8).tolist():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def test_c_octrn_reader_agrees_with_the_python_reader(tmp_path):
    """host/octrn.c (the reference's octarine_*_read_from_file, main.c:526-559) over the committed assets and over files
    written by the Python writer: extents and payload hashes equal what malevich_b200.assets reads."""
    import shutil
    from malevich_b200 import assets
    subprocess.run(["make", "-C", HOST_DIR], check=True, capture_output=True)
    d = tmp_path / "assets"
    d.mkdir()
    names = ["toon_sky_mesh", "ftm_ground_mesh", "ninomaru_teien_panorama_irradiance"]
    for n in names:
        shutil.copy(os.path.join(assets.ASSET_DIR, n + ".octrn"), d / (n + ".octrn"))
    tex = assets.standin_texture_srgb(1, size=64)
    assets.write_octrn_image(str(d / "standin_tex.octrn"), tex)
    vb, ib = assets.suprematist_scene()
    assets.write_octrn_mesh(str(d / "sup_mesh.octrn"), vb, ib)
    names += ["standin_tex", "sup_mesh"]
    out = subprocess.run([HOST_BIN, "--list-assets", str(d)] + names, check=True, capture_output=True, text=True).stdout.split("\n")
    got = {l.split()[1]: l.split() for l in out if l}
    for n in names:
        path = str(d / (n + ".octrn"))
        if got[n][0] == "mesh":
            v, i = assets.read_octrn_mesh(path)
            assert [int(got[n][2]), int(got[n][3])] == [v.shape[0], i.shape[0]]
            assert int(got[n][4], 16) == _fnv64(v.tobytes() + i.tobytes())
        else:
            t, w, h, fmt = assets.read_octrn_image(path)
            assert [int(got[n][2]), int(got[n][3]), int(got[n][4])] == [w, h, fmt]
            assert int(got[n][5], 16) == _fnv64(t.tobytes())
    assert got["sup_mesh"][0] == "mesh" and got["standin_tex"][0] == "image"
    # a truncated file is refused, not read past its end
    blob = open(d / "toon_sky_mesh.octrn", "rb").read()
    open(d / "bad_mesh.octrn", "wb").write(blob[:-5])
    r = subprocess.run([HOST_BIN, "--list-assets", str(d), "bad_mesh"], capture_output=True, text=True)
    assert r.returncode != 0 and "bad_mesh.octrn" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["toon", "suprematism"])
def test_c_host_loads_the_reference_scene_table(name, tmp_path):
    """render_host in the reference's own start-up form: init()'s scene table filled by load_mesh / load_texture from
    .octrn files (the stand-in sRGB textures are written to disk first; load_texture linearises them like main.c:546-558),
    the PerFrameCB supplied as bytes so that the frame can be compared with the committed golden frame; the PPM
    (replacement of the GDI blit) holds the same pixels."""
    import json
    import shutil
    from malevich_b200 import assets, camera
    subprocess.run(["make", "-C", HOST_DIR], check=True, capture_output=True)
    d = tmp_path / "assets"
    d.mkdir()
    for i, part in enumerate(["house", "sky"]):
        shutil.copy(os.path.join(assets.ASSET_DIR, f"toon_{part}_mesh.octrn"), d / f"toon_{part}_mesh.octrn")
        assets.write_octrn_image(str(d / f"toon_{part}_tex.octrn"), assets.standin_texture_srgb(i))
    cb_path, out_path, ppm_path = str(tmp_path / "cb.bin"), str(tmp_path / "out.bin"), str(tmp_path / "frame.ppm")
    camera.per_frame_cb(320, 200).astype(np.float32).tofile(cb_path)
    r = subprocess.run([HOST_BIN, "--assets", str(d), "--scene", name, "--size", "320x200", "--cb", cb_path, "--ppm", ppm_path, out_path, "2"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    col, dep, st = read_frame(out_path)
    key = {"toon": "toon_320x200", "suprematism": "sup_320x200"}[name]
    frames = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))
    parity.assert_frames_match(col, dep, frames[key + "/colors"], frames[key + "/depths"], key)
    assert st == json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))[key]["stats"]
    ppm = open(ppm_path, "rb").read()
    head = b"P6\n320 200\n255\n"
    assert ppm.startswith(head) and len(ppm) == len(head) + 320 * 200 * 3
    rgb = np.frombuffer(ppm, dtype=np.uint8, offset=len(head)).reshape(200, 320, 3)
    assert np.array_equal(rgb[..., 0], (col >> 16) & 0xFF) and np.array_equal(rgb[..., 1], (col >> 8) & 0xFF) and np.array_equal(rgb[..., 2], col & 0xFF)
    # the camera of init() + update() computed in C (no --cb) draws the same triangles
    r = subprocess.run([HOST_BIN, "--assets", str(d), "--scene", name, "--size", "320x200", out_path], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    col2, dep2, st2 = read_frame(out_path)
    assert st2["input_triangle_count"] == st["input_triangle_count"] and abs(st2["assembled_triangle_count"] - st["assembled_triangle_count"]) <= 2
    assert np.mean(col2 == col) > 0.99


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sup_320x200", "toon_320x200", "emily_320x200", "loco_320x200"])
def test_c_host_renders_like_the_reference(name, tmp_path):
    subprocess.run(["make", "-C", HOST_DIR], check=True, capture_output=True)
    sc = cases.SMALL[name]()
    scene_path, out_path = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(scene_path, sc)
    r = subprocess.run([HOST_BIN, scene_path, out_path, "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    col, dep, st = read_frame(out_path)
    frames = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))
    parity.assert_frames_match(col, dep, frames[name + "/colors"], frames[name + "/depths"], name)
    import json
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))[name]
    assert st == golden["stats"]
