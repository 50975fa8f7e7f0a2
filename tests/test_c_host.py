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
