"""GPU tests (`-m gpu`) of command lists (mlv_begin/finish/execute_command_list): a recorded frame replayed with one CUDA-graph
launch must give exactly the frame the immediate calls give -- image, depth, Stats -- also across constant-buffer updates,
buffer uploads between executions, arenas that grow while recording, and sort-first ranks."""
import os

import numpy as np
import pytest

import cases
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _device(*a, **k):
    from malevich_b200 import Device
    return Device(*a, **k)


def _immediate(sc):
    from malevich_b200 import scenes
    with _device(sc.width, sc.height) as dev:
        dev.reset_stats()
        scenes.render(dev, sc)
        col, dep = dev.present()
        return col, dep, dev.stats()


@pytest.mark.parametrize("name", ["toon_320x200", "ftm_320x200", "emily_320x200", "synth_320x200"])
def test_recorded_frame_replays_bit_identically(name):
    from malevich_b200 import scenes
    sc = cases.SMALL[name]()
    ref_col, ref_dep, ref_stats = _immediate(sc)
    with _device(sc.width, sc.height) as dev:  # fresh device: every arena grows WHILE recording
        def frame():
            dev.reset_stats()
            scenes.render(dev, sc)
            dev.resolve()
        cl = dev.record(frame)
        info = cl.info
        assert info["draws"] == len(sc.objects) and info["kernel_launches"] >= 4 * len(sc.objects) + 2
        n0 = dev.kernel_launch_count
        for _ in range(3):
            cl.execute()
            col, dep = dev.present()
            assert np.array_equal(col, ref_col) and np.array_equal(dep.view(np.uint32), ref_dep.view(np.uint32))
            assert dev.stats() == ref_stats  # reset_stats is part of the recording, like memset(&stats, 0) in render() (main.c:1268)
        assert dev.kernel_launch_count - n0 == 3 * (info["kernel_launches"] + 1)  # + the resolve of each present()
        # immediate mode still works after executions (vertex cache side stream re-joins) ...
        dev.reset_stats()
        scenes.render(dev, sc)
        col, dep = dev.present()
        assert np.array_equal(col, ref_col) and np.array_equal(dep.view(np.uint32), ref_dep.view(np.uint32)) and dev.stats() == ref_stats
        # ... and so does the list after immediate frames
        cl.execute()
        col, dep = dev.present()
        assert np.array_equal(col, ref_col) and dev.stats() == ref_stats
        cl.release()


def test_constants_of_a_recorded_list_follow_the_camera():
    """update() rewrites the PerFrameCB every frame (main.c:1595-1597): mlv_command_list_set_constants replaces it in the
    recorded geometry kernels; the replay equals an immediate frame with that camera."""
    from malevich_b200 import camera, scenes
    sc = cases.SMALL["ftm_320x200"]()
    poses = [((-8.0, 5.0, 1.2), -2.8, 0.1), ((-2.0, 1.5, 1.0), -2.0, 0.3), scenes.FTM_SCREENSHOT_POSE]
    with _device(sc.width, sc.height) as dev:
        cl = dev.record(lambda: scenes.render(dev, sc))
        for pose in poses:
            cb = camera.per_frame_cb(sc.width, sc.height, *pose)
            want_col, want_dep, _ = _immediate(scenes.ftm(sc.width, sc.height, cb=cb))
            cl.set_constants(cb)
            cl.execute()
            col, dep = dev.present()
            assert np.array_equal(col, want_col) and np.array_equal(dep.view(np.uint32), want_dep.view(np.uint32)), str(pose)
        # one draw only: the sky (last draw) rendered with another camera than the rest
        cb_a, cb_b = camera.per_frame_cb(sc.width, sc.height, *poses[0]), camera.per_frame_cb(sc.width, sc.height, *poses[1])
        cl.set_constants(cb_a)
        cl.set_constants(cb_b, draw_index=len(sc.objects) - 1)
        cl.execute()
        col, dep = dev.present()
        with _device(sc.width, sc.height) as imm:
            sa = scenes.ftm(sc.width, sc.height, cb=cb_a)
            scenes.render(imm, scenes.Scene(sa.name, sa.width, sa.height, sa.objects[:-1], sa.per_frame_cb))
            sb = scenes.ftm(sc.width, sc.height, cb=cb_b)
            scenes.render(imm, scenes.Scene(sb.name, sb.width, sb.height, sb.objects[-1:], sb.per_frame_cb), clear=False)
            want_col, want_dep = imm.present()
        assert np.array_equal(col, want_col) and np.array_equal(dep.view(np.uint32), want_dep.view(np.uint32))
        with pytest.raises(Exception):
            cl.set_constants(cb_a, draw_index=len(sc.objects))
        cl.release()


def test_buffer_uploads_between_executions_and_recording_rules():
    import ctypes as C
    from malevich_b200 import _lib as L, scenes
    sc = cases.SMALL["toon_320x200"]()
    ref_col, ref_dep, _ = _immediate(sc)
    with _device(sc.width, sc.height) as dev:
        scenes.upload(dev, sc)
        dev.begin_command_list()
        scenes.render(dev, sc)
        for bad in (dev.present, dev.stats, dev.finish, dev.begin_command_list, lambda: dev.present_async(ref_col.copy())):
            with pytest.raises(Exception):
                bad()  # nothing that synchronises or reads back can be recorded
        cl = dev.finish_command_list()
        with pytest.raises(Exception):
            dev.finish_command_list()  # nothing is being recorded
        cl.execute()
        col, dep = dev.present()
        assert np.array_equal(col, ref_col) and np.array_equal(dep.view(np.uint32), ref_dep.view(np.uint32))
        # new contents for the first object's vertex buffer (shifted a little): the recording binds the object, not the bytes
        o = sc.objects[0]
        moved = o.vertex_buffer.copy()
        moved[:, 2] += 0.05
        h = dev._buffer(o.vertex_buffer, L.BUFFER_VERTEX)
        L.check(dev._lib.mlv_update_buffer(dev._h, h, moved.ctypes.data_as(C.c_void_p), moved.nbytes))
        cl.execute()
        col2, dep2 = dev.present()
        sc2 = scenes.Scene(sc.name, sc.width, sc.height, [scenes.SceneObject(moved, o.index_buffer, o.vertex_shader, o.pixel_shader, o.texture)] + sc.objects[1:], sc.per_frame_cb)
        want_col, want_dep, _ = _immediate(sc2)
        assert np.array_equal(col2, want_col) and np.array_equal(dep2.view(np.uint32), want_dep.view(np.uint32))
        assert not np.array_equal(col2, ref_col)
        cl.release()


def test_command_lists_of_sort_first_ranks_compose_to_the_single_device_image():
    """Every rank of a sort-first split replays its own recorded frame (the exchange stays immediate: pack, gather, unpack)."""
    import torch
    from malevich_b200 import scenes
    from test_gpu_parity import _as_tensor
    sc = cases.SMALL["synth_320x200"]()
    ref_col, _, ref_stats = _immediate(sc)
    for world, stripe in ((2, 13), (4, 1)):
        devs = [_device(sc.width, sc.height, num_ranks=world, rank=r, stripe_height_tiles=stripe) for r in range(world)]
        try:
            lists = []
            for d in devs:
                def frame(d=d):
                    d.reset_stats()
                    scenes.render(d, sc)
                    d.composite_pack()
                lists.append(d.record(frame))
            for rep in range(2):
                sums = {"assembled_triangle_count": 0, "active_bin_count": 0, "total_triangle_count_in_bins": 0}
                for d, cl in zip(devs, lists):
                    cl.execute()
                    d.finish()
                    st = d.stats()
                    for k in sums:
                        sums[k] += st[k]
                assert sums == {k: ref_stats[k] for k in sums}
                _, chunk = devs[0].composite_layout()
                for r, d in enumerate(devs):
                    src_ptr, _ = d.composite_layout()
                    src = _as_tensor(src_ptr + r * chunk, chunk)
                    for d2 in devs:
                        dst_ptr, _ = d2.composite_layout()
                        _as_tensor(dst_ptr + r * chunk, chunk).copy_(src)
                torch.cuda.synchronize()
                for d in devs:
                    d.composite_unpack()
                    d.finish()
                    out = _as_tensor(d.resolved_color_ptr(), sc.width * sc.height * 4).cpu().numpy().view(np.uint32).reshape(sc.height, sc.width)
                    assert np.array_equal(out, ref_col), f"world {world} stripe {stripe} rep {rep}"
            for cl in lists:
                cl.release()
        finally:
            for d in devs:
                d.close()


@pytest.mark.parametrize("num_ranks", [1, 2])
def test_two_recordings_alternate_framebuffers_under_asynchronous_presents(num_ranks):
    """A frame recorded twice (each recording opens with a full clear, so each takes the other tiled framebuffer of the pair),
    the two lists replayed in turn with a different camera each time, every frame handed to an asynchronous present whose
    resolve / pack and copy run on the read-back stream while the next frame is already being drawn: every delivered frame
    is the frame of ITS camera, bit for bit what immediate mode gives."""
    from malevich_b200 import camera, scenes
    sc = cases.SMALL["ftm_320x200"]()
    poses = [((-8.0, 5.0, 1.2), -2.8, 0.1), ((-2.0, 1.5, 1.0), -2.0, 0.3), scenes.FTM_SCREENSHOT_POSE, ((-6.0, 3.0, 2.0), -2.5, 0.2)]
    cbs = [camera.per_frame_cb(sc.width, sc.height, *p) for p in poses]
    want = []
    for cb in cbs:
        col, dep, _ = _immediate(scenes.ftm(sc.width, sc.height, cb=cb))
        want.append(col)
    with _device(sc.width, sc.height, num_ranks=num_ranks, rank=0, stripe_height_tiles=3) as dev:
        scenes.upload(dev, sc)
        lists = [dev.record(lambda: scenes.render(dev, sc)) for _ in range(2)]
        frames = [np.zeros((sc.height, sc.width), np.uint32) for _ in range(len(cbs) * 3)]
        owned = np.array([((y // 8) // 3) % num_ranks == 0 for y in range(sc.height)])
        for it in range(len(frames)):
            cl = lists[it % 2]
            cl.set_constants(cbs[it % len(cbs)])
            cl.execute()
            if it >= 2:
                dev.present_wait()  # bounds the frames in flight like a double-buffered host does
            dev.present_owned_rows_async(frames[it])  # (one rank: the whole frame; rank 0 of 2: the rows it owns)
        dev.present_wait()
        dev.finish()
        for it, f in enumerate(frames):
            assert np.array_equal(f[owned], want[it % len(cbs)][owned]), it
            assert not f[~owned].any()
        for cl in lists:
            cl.release()


def test_a_list_without_clears_draws_over_the_contents_of_its_own_framebuffer():
    """A recording that does not open with a full clear is bound to the tiled framebuffer it was recorded on. Clears that
    are still pending from immediate mode are issued THERE (seen through a clear colour that changes every frame); when the
    contents it would draw over have moved to the other framebuffer of the pair (an asynchronous present made the next
    cleared frame move over) the execution is refused."""
    from malevich_b200 import scenes
    sc = cases.SMALL["loco_320x200"]()  # one small mesh: 98 % of the frame shows the clear colour
    colors = [(0.1, 0.2, 0.3, 1.0), (0.9, 0.1, 0.4, 1.0), (0.2, 0.8, 0.6, 1.0)]

    def frame(dev, rgba, draw):
        dev.clear_render_target_view(rgba)
        dev.clear_depth_stencil_view(scenes.CLEAR_DEPTH)
        draw()
    want = []
    with _device(sc.width, sc.height) as one:
        for rgba in colors:
            frame(one, rgba, lambda: scenes.render(one, sc, clear=False))
            want.append(one.present())
    assert not np.array_equal(want[0][0], want[1][0])  # the background shows
    with _device(sc.width, sc.height) as dev:
        scenes.upload(dev, sc)
        scenes.render(dev, sc)  # sizes the arenas
        dev.finish()
        draws = dev.record(lambda: scenes.render(dev, sc, clear=False))
        host = np.zeros((sc.height, sc.width), np.uint32)
        for it, rgba in enumerate(colors):
            # the asynchronous present of the previous frame keeps the framebuffer busy: the pending full clear must still be
            # issued on the LIST's framebuffer (a clear that moved to the other one would leave the old background behind)
            frame(dev, rgba, draws.execute)
            dev.present_async(host)
            if it == len(colors) - 1:
                dev.present_wait()
                assert np.array_equal(host, want[it][0]), it
        col, dep = dev.present()
        assert np.array_equal(col, want[-1][0]) and np.array_equal(dep.view(np.uint32), want[-1][1].view(np.uint32))
        # an immediate frame that opens with a full clear while a present still reads the framebuffer moves to the other one ...
        dev.present_async(host)
        frame(dev, colors[0], lambda: scenes.render(dev, sc, clear=False))
        # ... and the list, which draws over "what is there", must not silently draw over the old image in its own framebuffer
        with pytest.raises(Exception) as e:
            draws.execute()
        assert "other tiled framebuffer" in str(e.value)
        dev.present_wait()
        col, dep = dev.present()  # the immediate frame is intact
        assert np.array_equal(col, want[0][0]) and np.array_equal(dep.view(np.uint32), want[0][1].view(np.uint32))
        draws.release()
