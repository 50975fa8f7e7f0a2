"""Device groups (mlv_device_desc.num_gpus > 1): the multi-GPU fan-out behind the single-device calls ("multi-GPU
fan-out hidden behind the same calls"). On a single-GPU box every rank of the group runs on cuda:0 (MLV_DEVICE_GROUP_SAME_GPU):
the host-side fan-out, the peer-pointer exchange and the Stats sums are the code a real N-GPU group runs."""
import json
import os
import subprocess

import numpy as np
import pytest

import cases
import parity
from conftest import ROOT
from malevich_b200 import Device, scenes

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
FRAMES = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))


@pytest.mark.parametrize("name,num_gpus,stripe,peer", [("toon_320x200", 2, 0, False), ("toon_320x200", 3, 1, False), ("emily_320x200", 2, 0, True), ("loco_320x200", 4, 2, False),
                                                       ("sup_320x200", 2, 0, False), ("toon_320x200", 3, 1, True), ("loco_320x200", 4, 3, True)])
def test_group_frame_equals_the_golden_frame(name, num_gpus, stripe, peer):
    """peer=False: the frame is composed in host memory (every rank copies the rows it owns); peer=True: composed on the device
    by the peer-memory exchange, rank 0's image read back."""
    sc = cases.SMALL[name]()
    with Device(sc.width, sc.height, cuda_device=0, num_gpus=num_gpus, group_same_gpu=True, stripe_height_tiles=stripe, group_peer_exchange=peer) as dev:
        for _ in range(2):  # the second frame draws into the other framebuffer of every rank's pair
            scenes.render(dev, sc)
            col, dep = dev.present()
            st = dev.stats()
            dev.reset_stats()
            parity.assert_frames_match(col, dep, FRAMES[name + "/colors"], FRAMES[name + "/depths"], f"{name} on a group of {num_gpus}")
            assert st == GOLDEN[name]["stats"]
        assert dev.kernel_launch_count > 0


def test_group_command_list_and_constants():
    """A frame recorded once per GPU, replayed with one graph launch per GPU, its PerFrameCB replaced between replays."""
    from malevich_b200 import camera
    sc = cases.SMALL["toon_320x200"]()
    with Device(sc.width, sc.height, cuda_device=0, num_gpus=2, group_same_gpu=True, stripe_height_tiles=0) as dev, Device(sc.width, sc.height, cuda_device=0) as one:
        scenes.upload(dev, sc)
        cl = dev.record(lambda: scenes.render(dev, sc))
        cl.execute()
        col, dep = dev.present()
        parity.assert_frames_match(col, dep, FRAMES["toon_320x200/colors"], FRAMES["toon_320x200/depths"], "recorded group frame")
        cb = camera.per_frame_cb(sc.width, sc.height, (3.4, 1.1, 1.0), 0.1, -0.05)
        cl.set_constants(cb)
        cl.execute()
        col2, dep2 = dev.present()
        cl.release()
        sc2 = scenes.toon(sc.width, sc.height, cb=cb)
        scenes.render(one, sc2)
        ref_col, ref_dep = one.present()
        assert np.array_equal(dep2, ref_dep) and np.array_equal(col2, ref_col)  # the same kernels: bit-identical to one GPU


def test_entry_points_outside_the_fan_out_refuse_a_group():
    from malevich_b200 import _lib as L
    with Device(320, 200, cuda_device=0, num_gpus=2, group_same_gpu=True, stripe_height_tiles=0) as dev:
        with pytest.raises(Exception) as e:
            dev.composite_pack()
        assert "device group" in str(e.value)
    # a resource of a single device is not accepted by a group, and the other way round
    import ctypes as C
    import numpy as np
    with Device(320, 200, cuda_device=0, num_gpus=2, group_same_gpu=True, stripe_height_tiles=0) as grp, Device(320, 200, cuda_device=0) as one:
        vb = np.zeros((3, 8), dtype=np.float32)
        h_one, h_grp = C.c_void_p(), C.c_void_p()
        L.check(one._lib.mlv_create_buffer(one._h, vb.ctypes.data_as(C.c_void_p), vb.nbytes, L.BUFFER_VERTEX, C.byref(h_one)))
        L.check(grp._lib.mlv_create_buffer(grp._h, vb.ctypes.data_as(C.c_void_p), vb.nbytes, L.BUFFER_VERTEX, C.byref(h_grp)))
        assert grp._lib.mlv_ia_set_vertex_buffer(grp._h, h_one) == L.MLV_ERR_INVALID_ARGUMENT
        assert one._lib.mlv_ia_set_vertex_buffer(one._h, h_grp) == L.MLV_ERR_INVALID_ARGUMENT
        assert grp._lib.mlv_ia_set_vertex_buffer(grp._h, h_grp) == L.MLV_OK
        grp._lib.mlv_ia_set_vertex_buffer(grp._h, None)
        one._lib.mlv_release_buffer(one._h, h_one)
        grp._lib.mlv_release_buffer(grp._h, h_grp)


def test_c_host_drives_a_group(tmp_path):
    """host/render_host: the reference's render() over the compat shim with MLV_NUM_GPUS=2 (both ranks on cuda:0)."""
    import test_c_host as T
    subprocess.run(["make", "-C", T.HOST_DIR], check=True, capture_output=True)
    sc = cases.SMALL["toon_320x200"]()
    scene_path, out_path = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    T.write_scene(scene_path, sc)
    env = dict(os.environ, MLV_NUM_GPUS="2", MLV_GROUP_SAME_GPU="1")
    r = subprocess.run([T.HOST_BIN, scene_path, out_path, "2"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stderr
    col, dep, st = T.read_frame(out_path)
    parity.assert_frames_match(col, dep, FRAMES["toon_320x200/colors"], FRAMES["toon_320x200/depths"], "C host on a group of 2")
    assert st == GOLDEN["toon_320x200"]["stats"]


@pytest.mark.parametrize("num_ranks,stripe", [(2, 13), (3, 1), (5, 2)])
def test_owned_rows_of_separate_rank_devices_compose_in_host_memory(num_ranks, stripe):
    """mlv_present_owned_rows_async from one device per rank (what one process per GPU does with a shared pinned frame): every
    rank delivers exactly the rows it owns, several frames in a row (the two packed chunks alternate), into one host frame."""
    sc = cases.SMALL["toon_320x200"]()
    devs = [Device(sc.width, sc.height, cuda_device=0, num_ranks=num_ranks, rank=r, stripe_height_tiles=stripe) for r in range(num_ranks)]
    try:
        frame = np.zeros((sc.height, sc.width), dtype=np.uint32)
        if num_ranks == 3:  # page-locked through the library (a host that does not link CUDA): the copies are then really asynchronous
            devs[0].register_host_memory(frame.ctypes.data, frame.nbytes)
        for it in range(3):
            frame[...] = 0xDEADBEEF
            for d in devs:
                scenes.render(d, sc)
                d.present_owned_rows_async(frame)
            for d in devs:
                d.present_wait()
            want = FRAMES["toon_320x200/colors"]
            assert parity.compare_frames(frame, FRAMES["toon_320x200/depths"], want, FRAMES["toon_320x200/depths"])["color_max_diff"] <= 1, it
            assert not np.any(frame == 0xDEADBEEF)
    finally:
        if num_ranks == 3:
            devs[0].finish()
            devs[0].unregister_host_memory(frame.ctypes.data)
        for d in devs:
            d.close()
