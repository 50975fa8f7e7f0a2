"""GPU parity tests (`-m gpu`): the CUDA path, called through the C-ABI, against
(1) the committed golden fixtures generated from the reference, and (2) the reference itself
(oracle/_ref) run live on the same inputs, stage by stage.

Bars: assembled triangles, bin lists (per-tile order), coverage masks, tile minima, depth, Stats are
BIT-EXACT; colour within 1/255 per channel on >= 99.9 % of pixels, none off by more than 2/255
(BASELINE.json north_star).
"""
import ctypes
import json
import os
import sys

import numpy as np
import pytest

import cases
import parity
from conftest import ROOT, have_ref

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _fnv(a):
    from oracle.ref_oracle import fnv64_words
    return fnv64_words(a)


def _device(*a, **k):
    from malevich_b200 import Device
    return Device(*a, **k)


def _oracle(w, h):
    from oracle.ref_oracle import RefOracle
    return RefOracle(w, h, threads=1)


@pytest.mark.parametrize("name", sorted(cases.SMALL))
def test_small_cases_against_committed_golden_frames(name):
    from malevich_b200 import scenes
    frames = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))
    sc = cases.SMALL[name]()
    assert sc.per_frame_cb.tobytes().hex() == GOLDEN[name]["per_frame_cb_hex"], "constant buffer differs from the one the golden frames were made with"
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        col, dep = dev.present()
        assert dev.stats() == GOLDEN[name]["stats"]
    parity.assert_frames_match(col, dep, frames[name + "/colors"], frames[name + "/depths"], name)
    assert _fnv(dep) == GOLDEN[name]["depth_fnv"]


@pytest.mark.parametrize("name", sorted(cases.SMALL) + sorted(cases.FULL))
def test_every_stage_against_live_reference(name):
    """Draw by draw: VS output, assembled triangles + attributes, bin lists in order, coverage masks, tile minima."""
    sc = {**cases.SMALL, **cases.FULL}[name]()
    if not have_ref(sc.width, sc.height):
        pytest.skip("oracle/_ref not available for this resolution")
    orc = _oracle(sc.width, sc.height)
    if not parity.host_vrsqrtps_matches_table(orc):
        pytest.skip("host vrsqrtps differs from the committed Intel table: live oracle not canonical here (golden-frame tests still apply)")
    with _device(sc.width, sc.height, debug_capture=True) as dev:
        results = parity.render_both_staged(dev, orc, sc)
        for draw_name, res in results:
            assert parity.staged_ok(res), f"{name}/{draw_name}: {res}"
        col, dep = dev.present()
        assert dev.stats() == orc.stats() == GOLDEN[name]["stats"]
    parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), name)
    assert _fnv(dep) == GOLDEN[name]["depth_fnv"]


def test_config5_full_size_against_live_reference():
    """BASELINE.json config 5 at full size: 10 M triangles, 3840x2160, 8 draws."""
    from malevich_b200 import scenes
    name = "config5_synthetic_3840x2160"
    sc = cases.CONFIG5[name]()
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        col, dep = dev.present()
        st = dev.stats()
    assert st == GOLDEN[name]["stats"]
    assert _fnv(dep) == GOLDEN[name]["depth_fnv"]  # depth bit-exact against the reference, via the committed hash
    if have_ref(sc.width, sc.height):
        orc = _oracle(sc.width, sc.height)
        orc.render(sc)  # (basic_ps does not read the normal, so this comparison does not depend on the host's vrsqrtps)
        parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), name)


def test_full_size_properties_without_oracle():
    """Size-independent properties at full size: determinism (two runs bit-identical), idempotence of re-drawing
    the same geometry (depth unchanged, Hi-Z rejects nothing it should not), clear restores the cleared state."""
    from malevich_b200 import scenes
    sc = cases.FULL["config2_ftm_1920x1080"]()
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        c1, d1 = dev.present()
        scenes.render(dev, sc)
        c2, d2 = dev.present()
        assert np.array_equal(c1, c2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
        scenes.render(dev, sc, clear=False)  # same geometry again on top: z >= depth passes on ties with identical colour
        c3, d3 = dev.present()
        assert np.array_equal(d1.view(np.uint32), d3.view(np.uint32)) and np.array_equal(c1, c3)
        dev.clear_render_target_view(scenes.CLEAR_COLOR)
        dev.clear_depth_stencil_view(0.0)
        c4, d4 = dev.present()
        assert np.all(d4 == 0.0) and np.all(c4 == 0x00D8DFE3)  # encode_color_as_u32: R in the low byte (math.h:322-324)
        assert np.all(dev.debug_tile_min_depths() == 0.0)


def test_draw_without_indices_equals_draw_indexed_with_identity_indices():
    from malevich_b200 import scenes
    sc = cases.SMALL["toon_320x200"]()
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        c1, d1 = dev.present()
        # same frame with draw(): bind each object, no index buffer
        dev.clear_render_target_view(scenes.CLEAR_COLOR)
        dev.clear_depth_stencil_view(0.0)
        gp = dev.graphics_pipeline
        for o in sc.objects:
            assert np.array_equal(o.index_buffer, np.arange(o.index_count, dtype=np.uint32))
            gp.ia.p_vertex_buffer, gp.ia.p_index_buffer = o.vertex_buffer, None
            gp.vs.p_shader_resource_views[0] = gp.ps.p_shader_resource_views[0] = o.texture
            dev.draw(o.index_count)
        c2, d2 = dev.present()
    assert np.array_equal(c1, c2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))


def test_draw_indexed_extensions_16bit_indices_start_index_base_vertex():
    """The reference's TODOs (main.c:72, 1219) implemented: u16 index buffers, StartIndexLocation, BaseVertexLocation.
    Oracle: the same triangles drawn the reference's way (u32 indices from 0) must give the same frame and Stats."""
    from malevich_b200 import scenes
    sc = cases.SMALL["emily_320x200"]()  # indexed sphere (vertex reuse) + fullscreen quad
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        ref_col, ref_dep = dev.present()
        ref_stats = dev.stats()
        # (a) u16 indices
        o = sc.objects[0]
        assert o.vertex_buffer.shape[0] < 65536
        sc16 = scenes.Scene(sc.name, sc.width, sc.height, [scenes.SceneObject(o.vertex_buffer, o.index_buffer.astype(np.uint16), o.vertex_shader, o.pixel_shader, o.texture)] + sc.objects[1:],
                            sc.per_frame_cb)
        dev.reset_stats()
        scenes.render(dev, sc16)
        c, d = dev.present()
        assert np.array_equal(c, ref_col) and np.array_equal(d.view(np.uint32), ref_dep.view(np.uint32)) and dev.stats() == ref_stats
        # (b) one shared vertex/index buffer pair, sub-ranges addressed with start index + base vertex
        pad_v = np.zeros((5, 8), np.float32)          # junk in front so that base_vertex matters
        pad_i = np.full(24, 3, np.uint32)              # junk in front so that start_index matters
        vb_all = np.concatenate([pad_v, o.vertex_buffer])
        ib_all = np.concatenate([pad_i, o.index_buffer])
        dev.reset_stats()
        dev.clear_render_target_view(scenes.CLEAR_COLOR)
        dev.clear_depth_stencil_view(0.0)
        gp = dev.graphics_pipeline
        gp.ia.p_vertex_buffer, gp.ia.p_index_buffer = vb_all, ib_all
        gp.vs.shader, gp.ps.shader = o.vertex_shader, o.pixel_shader
        gp.vs.p_shader_resource_views[0] = gp.ps.p_shader_resource_views[0] = o.texture
        dev.draw_indexed(o.index_count, start_index_location=24, base_vertex_location=5)
        q = sc.objects[1]
        gp.ia.p_vertex_buffer, gp.ia.p_index_buffer = q.vertex_buffer, q.index_buffer
        gp.vs.shader, gp.ps.shader = q.vertex_shader, q.pixel_shader
        gp.vs.p_shader_resource_views[0] = gp.ps.p_shader_resource_views[0] = q.texture
        dev.draw_indexed(q.index_count)
        c, d = dev.present()
        assert np.array_equal(c, ref_col) and np.array_equal(d.view(np.uint32), ref_dep.view(np.uint32)) and dev.stats() == ref_stats


def test_edge_cases_and_error_behaviour():
    from malevich_b200 import MalevichError, scenes
    from malevich_b200 import _lib as L
    sc = cases.SMALL["sup_320x200"]()
    with _device(sc.width, sc.height, debug_capture=True) as dev:
        # the reference's asserts (main.c:666,670,1230) surface as errors
        gp = dev.graphics_pipeline
        o = sc.objects[0]
        with pytest.raises(MalevichError):
            dev.draw_indexed(24)  # nothing bound
        scenes.render(dev, sc)
        with pytest.raises(MalevichError) as e:
            dev.draw_indexed(20)  # not divisible by 8
        assert e.value.code == L.MLV_ERR_INVALID_ARGUMENT
        with pytest.raises(MalevichError):
            dev.draw_indexed(16)  # divisible by 8, not by 3
        with pytest.raises(MalevichError):
            dev.draw_indexed(48)  # beyond the bound index buffer
        gp.rs.viewport.width = 100.0
        with pytest.raises(MalevichError):
            dev.draw_indexed(24)  # viewport must equal the render target
        gp.rs.viewport.width = float(sc.width)
        # empty draw: nothing happens, nothing breaks
        dev.draw_indexed(0)
        # all-degenerate draw (the reference's own (0,0,0) padding triangles): zero-area triangles are KEPT and binned (main.c:856)
        dev.reset_stats()
        gp.ia.p_index_buffer = np.zeros(24, dtype=np.uint32)
        dev.draw_indexed(24)
        st = dev.stats()
        assert st["input_triangle_count"] == 8 and st["assembled_triangle_count"] == 8
        infos = dev.debug_masks()
        assert len(infos) == st["total_triangle_count_in_bins"] and np.all(infos["fragment_mask"] == 0)
        # w == 0 vertices are dropped (main.c:759); NaN positions neither crash nor assemble
        vb = o.vertex_buffer.copy()
        vb[:, 3] = 0.0
        gp.ia.p_vertex_buffer, gp.ia.p_index_buffer = vb, o.index_buffer
        dev.reset_stats()
        dev.draw_indexed(24)
        assert dev.stats()["assembled_triangle_count"] == 0
        vb2 = o.vertex_buffer.copy()
        vb2[:, 0] = np.nan
        gp.ia.p_vertex_buffer = vb2
        dev.draw_indexed(24)
        dev.present()


def test_clipping_heavy_camera_against_live_reference():
    """Camera inside the FTM geometry: near-plane and side-plane clipping (main.c:609-660) on many triangles."""
    from malevich_b200 import camera, scenes
    if not have_ref(320, 200):
        pytest.skip("oracle/_ref not available")
    for pose in (((-2.0, 1.5, 1.0), -2.0, 0.3), ((0.5, 0.2, 0.6), 0.7, -0.4), ((3.5, 1.0, 1.0), 3.0, 1.2)):
        cb = camera.per_frame_cb(320, 200, *pose)
        sc = scenes.ftm(320, 200, cb=cb)
        orc = _oracle(320, 200)
        if not parity.host_vrsqrtps_matches_table(orc):
            pytest.skip("host vrsqrtps differs from the committed Intel table")
        with _device(320, 200, debug_capture=True) as dev:
            for draw_name, res in parity.render_both_staged(dev, orc, sc):
                assert parity.staged_ok(res), f"{pose}/{draw_name}: {res}"
            col, dep = dev.present()
            assert dev.stats() == orc.stats()
        parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), str(pose))


def test_sort_first_partition_composes_to_single_gpu_image():
    """SURVEY.md 8e on ONE GPU: render each rank's stripes with its own device object, exchange the packed chunks by
    hand (what ncclAllGather does), unpack -- the composed image must be bit-identical to the single-device image and
    the per-rank counters (assembled triangles, bins, pairs) must sum to the single-device Stats. Covers interleaved
    stripes and contiguous bands (where chunk culling and the band-restricted scan/clear are active)."""
    for case, combos in (("ftm_320x200", ((2, 1), (4, 2), (3, 1), (8, 1), (2, 13), (4, 7), (5, 5))), ("synth_320x200", ((2, 13), (8, 4), (3, 2))),
                         ("emily_320x200", ((2, 13), (4, 7))), ("loco_320x200", ((2, 13), (3, 9)))):
        _check_partition(cases.SMALL[case](), combos)


def test_sort_first_partition_with_several_chunks_per_persistent_cta():
    """Same check on a mesh of 2 x 400 000 triangles (1 563 chunks of 256 per draw): with sort-first culling k_geom runs as
    a persistent grid of at most 592 CTAs, so every CTA walks several chunks and carries its Stats across them."""
    from malevich_b200 import scenes
    _check_partition(scenes.synthetic(1280, 720, layers=2, nx=500, ny=400), ((2, 45), (4, 23), (8, 2)))


def _check_partition(sc, combos):
    import torch
    from malevich_b200 import scenes
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        ref_col, ref_dep = dev.present()
        ref_stats = dev.stats()
    for world, stripe in combos:
        devs = [_device(sc.width, sc.height, num_ranks=world, rank=r, stripe_height_tiles=stripe) for r in range(world)]
        try:
            chunks, sums = [], {"assembled_triangle_count": 0, "active_bin_count": 0, "total_triangle_count_in_bins": 0}
            for d in devs:
                scenes.render(d, sc)
                d.composite_pack()
                d.finish()
                st = d.stats()
                assert st["vertex_count"] == ref_stats["vertex_count"] and st["input_triangle_count"] == ref_stats["input_triangle_count"]
                for k in sums:
                    sums[k] += st[k]
            assert sums == {k: ref_stats[k] for k in sums}
            ptr0, chunk = devs[0].composite_layout()
            # the "all-gather": copy rank r's chunk into every device's gather buffer at offset r*chunk
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
                assert np.array_equal(out, ref_col), f"world {world} stripe {stripe}"
        finally:
            for d in devs:
                d.close()


def test_peer_memory_composite_in_one_process():
    """mlv_composite_broadcast / mlv_composite_wait with every rank as a device object of this process on one GPU (plain
    pointers instead of cudaIpc mappings): after two frames (both image parities, increasing sequence numbers) every rank
    holds the complete image, bit-identical to the single-device one."""
    from malevich_b200 import scenes
    sc = cases.SMALL["ftm_320x200"]()
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        ref_col, _ = dev.present()
    for world, stripe in ((2, 13), (4, 2), (8, 1), (3, 5)):
        devs = [_device(sc.width, sc.height, num_ranks=world, rank=r, stripe_height_tiles=stripe) for r in range(world)]
        try:
            infos = [d.composite_peer_export() for d in devs]
            for d in devs:
                d.composite_peer_attach(infos, same_process=True)
            for frame in range(2):
                for d in devs:
                    scenes.render(d, sc)
                    d.composite_broadcast()
                for d in devs:  # all stripes are on their way before anybody spins (the streams of one GPU may share a hardware queue)
                    d.finish()
                for d in devs:
                    d.composite_wait()
                    d.finish()
                    out = _as_tensor(d.resolved_color_ptr(), sc.width * sc.height * 4).cpu().numpy().view(np.uint32).reshape(sc.height, sc.width)
                    assert np.array_equal(out, ref_col), f"world {world} stripe {stripe} frame {frame}"
                    d.stats()  # surfaces MLV_FLAG_COMPOSITE_TIMEOUT, if any
            with pytest.raises(Exception):
                devs[0].composite_wait()  # nothing to wait for
        finally:
            for d in devs:
                d.close()


def test_async_peer_memory_composite_in_one_process():
    """mlv_composite_broadcast_async / mlv_composite_join, pipelined the way bench.py runs them: frame f+1 is drawn (into the
    other tiled framebuffer of the pair) while frame f is still being exchanged; the joined image is always the complete
    frame f. Two different scenes alternate so that a stale or half-exchanged image cannot pass."""
    from malevich_b200 import scenes
    scs = [cases.SMALL["ftm_320x200"](), cases.SMALL["toon_320x200"]()]
    refs = []
    for sc in scs:
        with _device(sc.width, sc.height) as dev:
            scenes.render(dev, sc)
            refs.append(dev.present()[0])
    W, H = scs[0].width, scs[0].height
    for world, stripe in ((2, 13), (3, 2), (2, 1)):
        devs = [_device(W, H, num_ranks=world, rank=r, stripe_height_tiles=stripe) for r in range(world)]
        try:
            infos = [d.composite_peer_export() for d in devs]
            for d in devs:
                d.composite_peer_attach(infos, same_process=True)
            with pytest.raises(Exception):
                devs[0].composite_join()  # nothing to join
            # One host thread drives every rank here, so nothing that synchronises the whole GPU (cudaFree when an arena
            # grows, a caching-allocator miss) may happen while a rank's wait kernel spins for a peer this thread has not
            # issued yet: size the arenas and the result buffers first. (One process per GPU has no such coupling.)
            import torch
            for d in devs:
                for sc in scs:
                    scenes.render(d, sc)
                d.finish()
                d._got = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            frames = 5
            for f in range(frames + 1):
                for d in devs:
                    if f < frames:
                        scenes.render(d, scs[f % 2])
                    if f > 0:
                        d.composite_join()  # exchange of frame f-1
                        out = _as_tensor(d.resolved_color_ptr(), W * H * 4)
                        with torch.cuda.stream(torch.cuda.ExternalStream(d.stream)):
                            d._got.copy_(out)  # consumer on the device stream, issued before the next broadcast
                    if f < frames:
                        d.composite_broadcast_async()
                if f > 0:
                    for r, d in enumerate(devs):
                        d.finish()
                        got = d._got.cpu().numpy().view(np.uint32).reshape(H, W)
                        assert np.array_equal(got, refs[(f - 1) % 2]), f"world {world} stripe {stripe} frame {f - 1} rank {r}"
            for d in devs:
                d.stats()  # surfaces MLV_FLAG_COMPOSITE_TIMEOUT, if any
            # the synchronous form still works afterwards, and a frame without a full clear waits for the exchange
            for d in devs:
                scenes.render(d, scs[0])
                d.composite_broadcast_async()
                scenes.render(d, scs[1], clear=False)  # draws on top of frame 0 in the SAME framebuffer
                d.composite_join()
            for d in devs:
                d.finish()
                out = _as_tensor(d.resolved_color_ptr(), W * H * 4).cpu().numpy().view(np.uint32).reshape(H, W)
                assert np.array_equal(out, refs[0])
        finally:
            for d in devs:
                d.close()


def _ipc_rank(rank, world, conns, result_q):
    """One process = one rank; all ranks share cuda:0 here (on the real machine each has its own GPU)."""
    try:
        import numpy as np
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cases as cases_
        from malevich_b200 import Device, scenes
        sc = cases_.SMALL["toon_320x200"]()
        with Device(sc.width, sc.height) as single:
            scenes.render(single, sc)
            ref_col, _ = single.present()
        with Device(sc.width, sc.height, num_ranks=world, rank=rank, stripe_height_tiles=3) as dev:
            mine = dev.composite_peer_export()
            for c in conns:
                c.send((rank, mine))
            infos = {rank: mine}
            for c in conns:
                r, b = c.recv()
                infos[r] = b
            dev.composite_peer_attach([infos[r] for r in range(world)], same_process=False)
            ok = True
            for frame in range(3):
                scenes.render(dev, sc)
                dev.composite_broadcast()
                dev.composite_wait()
                dev.finish()
                dev.stats()
                import torch
                out = _as_tensor(dev.resolved_color_ptr(), sc.width * sc.height * 4).cpu().numpy().view(np.uint32).reshape(sc.height, sc.width)
                ok = ok and bool(np.array_equal(out, ref_col))
            # the pipelined form: draw frame f+1, join + consume frame f, start the exchange of frame f+1
            sc2 = cases_.SMALL["ftm_320x200"]()
            with Device(sc2.width, sc2.height) as single:
                scenes.render(single, sc2)
                ref2, _ = single.present()
            pair, refs2, got = [sc, sc2], [ref_col, ref2], []
            for frame in range(5):
                if frame < 4:
                    scenes.render(dev, pair[frame % 2])
                if frame > 0:
                    dev.composite_join()
                    with torch.cuda.stream(torch.cuda.ExternalStream(dev.stream)):
                        got.append(_as_tensor(dev.resolved_color_ptr(), sc.width * sc.height * 4).clone())
                if frame < 4:
                    dev.composite_broadcast_async()
            dev.finish()
            dev.stats()
            for frame, g in enumerate(got):
                ok = ok and bool(np.array_equal(g.cpu().numpy().view(np.uint32).reshape(sc.height, sc.width), refs2[frame % 2]))
            for c in conns:  # nobody unmaps while a peer may still be writing
                c.send("done")
            for c in conns:
                c.recv()
        result_q.put((rank, ok, ""))
    except Exception as e:  # noqa: BLE001
        result_q.put((rank, False, repr(e)))


def test_peer_memory_composite_across_processes_with_cuda_ipc():
    """Two processes, one rank each, exchanging mlv_peer_info over a pipe and writing into each other's images through
    cudaIpc mappings -- the multi-GPU data path of bench.py --gpus N, exercised on a single GPU."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    a, b = ctx.Pipe()
    q = ctx.Queue()
    procs = [ctx.Process(target=_ipc_rank, args=(0, 2, [a], q)), ctx.Process(target=_ipc_rank, args=(1, 2, [b], q))]
    for p in procs:
        p.start()
    try:
        results = [q.get(timeout=240) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for rank, ok, err in results:
        assert ok, f"rank {rank}: {err or 'image differs from the single-device image'}"


def _as_tensor(ptr, nbytes):
    import torch

    class _Raw:
        def __init__(self):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(_Raw(), device="cuda")


def test_work_counters_equal_stats_without_hiz_and_shrink_with_it():
    """mlv_work_counters: with Hi-Z at binning time disabled (debug capture keeps every pair, like the reference) every assembled
    triangle writes a record and every counted pair is listed, so the counters equal Stats; with it, hidden layers drop out."""
    from malevich_b200 import scenes
    sc = scenes.synthetic(1280, 720, layers=3, nx=400, ny=200)
    with _device(sc.width, sc.height, debug_capture=True) as dev:
        scenes.render(dev, sc)
        st, w = dev.stats(), dev.work_counters()
        assert w["records_written"] == st["assembled_triangle_count"]
        assert w["pairs_listed"] == st["total_triangle_count_in_bins"]
        assert w["tiles_visited"] == st["active_bin_count"]
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        st2, w2 = dev.stats(), dev.work_counters()
        assert st2 == st
        assert 0 < w2["records_written"] < st["assembled_triangle_count"] * 0.6   # layers 2 and 3 lie behind layer 1
        assert 0 < w2["pairs_listed"] < st["total_triangle_count_in_bins"] * 0.6
        assert (sc.width // 8) * (sc.height // 8) <= w2["tiles_visited"] <= st["active_bin_count"]
        dev.reset_stats()
        assert dev.work_counters() == {"records_written": 0, "pairs_listed": 0, "tiles_visited": 0}


def test_range_uploads_marked_on_a_caller_stream_and_composite_readback():
    """The sharded-upload entry points of the sort-first path, on one GPU: a buffer created with a padded capacity and
    filled by mlv_update_buffer_range in three pieces, declared complete on a caller's stream with
    mlv_buffer_mark_updated, draws the same frame; mlv_composite_readback_async delivers the composited image."""
    import ctypes as C
    import torch
    from malevich_b200 import _lib as L, scenes
    lib = L.load()
    sc = cases.SMALL["toon_320x200"]()
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        ref_col, ref_dep = dev.present()
    with _device(sc.width, sc.height) as dev:
        side = torch.cuda.Stream()
        keep = []
        for o in sc.objects:
            for arr, kind in ((o.vertex_buffer, L.BUFFER_VERTEX), (o.index_buffer, L.BUFFER_INDEX)):
                h = C.c_void_p()
                L.check(lib.mlv_create_buffer(dev._h, None, arr.nbytes + 4096, kind, C.byref(h)))  # capacity only, no contents
                dev._buffers[(id(arr), kind)] = (arr, h)
                raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
                keep.append(raw)
                cuts = [0, raw.nbytes // 3 // 16 * 16, raw.nbytes // 2 // 16 * 16, raw.nbytes]
                for lo, hi in zip(cuts[:-1], cuts[1:]):
                    L.check(lib.mlv_update_buffer_range(dev._h, h, lo, C.c_void_p(raw.ctypes.data + lo), hi - lo))
                assert int(lib.mlv_buffer_device_ptr(h)) != 0
                ev = torch.cuda.Event()
                ev.record(torch.cuda.ExternalStream(dev.copy_stream))
                side.wait_event(ev)
                L.check(lib.mlv_buffer_mark_updated(dev._h, h, C.c_void_p(side.cuda_stream)))
                assert lib.mlv_update_buffer_range(dev._h, h, arr.nbytes + 4096 - 8, C.c_void_p(raw.ctypes.data), 16) == L.MLV_ERR_INVALID_ARGUMENT
        scenes.render(dev, sc)
        col, dep = dev.present()
        assert np.array_equal(col, ref_col) and np.array_equal(dep, ref_dep)
        with pytest.raises(Exception):
            dev.composite_readback_async(col)  # single-rank device
    devs = [_device(sc.width, sc.height, num_ranks=2, rank=r, stripe_height_tiles=13) for r in range(2)]
    try:
        infos = [d.composite_peer_export() for d in devs]
        for d in devs:
            d.composite_peer_attach(infos, same_process=True)
        for d in devs:
            scenes.render(d, sc)
            d.finish()
        out = [torch.empty(sc.width * sc.height, dtype=torch.int32, pin_memory=True) for _ in devs]
        for d in devs:
            scenes.render(d, sc)
            d.composite_broadcast()
        for d in devs:
            d.finish()
        for d, o in zip(devs, out):
            with pytest.raises(Exception):
                d.composite_readback_async(o.numpy())  # exchange not complete yet
            d.composite_wait()
            d.composite_readback_async(o.numpy())
            d.present_wait()
            assert np.array_equal(o.numpy().view(np.uint32).reshape(sc.height, sc.width), ref_col)
    finally:
        for d in devs:
            d.close()


def test_async_present_and_streamed_uploads_match_blocking_present():
    """mlv_present_readback_async + mlv_present_wait deliver the image mlv_present_readback does, also when every buffer
    is re-uploaded before every frame (uploads on the copy stream, read-back on its own stream, two frames in flight)."""
    import ctypes as C
    from malevich_b200 import _lib as L
    from malevich_b200 import scenes
    sc = cases.SMALL["ftm_320x200"]()
    with _device(sc.width, sc.height) as dev:
        scenes.render(dev, sc)
        ref_col, ref_dep = dev.present()
        bufs = [(np.zeros_like(ref_col), np.zeros_like(ref_dep)) for _ in range(2)]
        for frame in range(4):
            for obj in sc.objects:
                for arr, kind in ((obj.vertex_buffer, L.BUFFER_VERTEX), (obj.index_buffer, L.BUFFER_INDEX)):
                    L.check(dev._lib.mlv_update_buffer(dev._h, dev._buffer(arr, kind), arr.ctypes.data_as(C.c_void_p), arr.nbytes))
            scenes.render(dev, sc)
            dev.present_wait()
            if frame:
                col, dep = bufs[(frame - 1) % 2]
                assert np.array_equal(col, ref_col) and np.array_equal(dep.view(np.uint32), ref_dep.view(np.uint32))
                col[...] = 0
            dev.present_async(*bufs[frame % 2])
        dev.finish()
        assert np.array_equal(bufs[1][0], ref_col)


def test_texture_srgb_to_linear_on_device_matches_reference_load_path():
    """load_texture's is_in_srgb branch (main.c:546-558) run on the GPU: exhaustive over all 256 byte values in every
    channel, a ragged size (texel count not a multiple of 4), and end to end -- TOON drawn with sRGB textures that the
    device re-quantises itself must reproduce the golden frame made with host-converted textures."""
    from malevich_b200 import Texture2D, assets, scenes
    b = np.arange(256, dtype=np.uint32)
    every_byte = (b | (b << 8) | (b << 16) | (b << 24)).reshape(16, 16)
    rng = np.random.default_rng(7)
    ragged = rng.integers(0, 2**32, size=(61, 127), dtype=np.uint32)
    sc = cases.SMALL["toon_320x200"]()
    with _device(sc.width, sc.height) as dev:
        for arr in (every_byte, ragged, assets.standin_texture_srgb(3, 128)):
            got = dev.read_texture(Texture2D(arr, is_in_srgb=True))
            assert np.array_equal(got, assets.srgb_texture_to_linear(arr))
            if have_ref(320, 200):
                assert np.array_equal(got, _oracle(320, 200).texture_srgb_to_linear(arr))
        for i, obj in enumerate(sc.objects):
            obj.texture = Texture2D(assets.standin_texture_srgb(i), is_in_srgb=True)
        scenes.render(dev, sc)
        col, dep = dev.present()
    frames = np.load(os.path.join(ROOT, "tests", "golden", "small_frames.npz"))
    parity.assert_frames_match(col, dep, frames["toon_320x200/colors"], frames["toon_320x200/depths"], "toon with device-converted sRGB textures")


def test_no_cpu_fallback_and_kernels_launch():
    from malevich_b200 import scenes
    sc = cases.SMALL["sup_320x200"]()
    with _device(sc.width, sc.height) as dev:
        n0 = dev.kernel_launch_count
        scenes.render(dev, sc)
        dev.present()
        assert dev.kernel_launch_count - n0 == 1 + 7 + 1  # fused clear + (vertex cache, front, front clip, back, scan, fill, tile) + resolve


# ---- production path (no debug capture): the kernels bench.py times --------------------------------------------------

def _render_production_draw_by_draw(dev, orc, sc, per_draw):
    """Same frame on the production device (Hi-Z at binning time, unsorted lists, vertex cache, no record for hidden
    triangles) and on the live reference, draw by draw; `per_draw(i, obj, tile_min_before)` runs after each pair of draws."""
    from malevich_b200 import scenes as S
    import malevich_b200._lib as L
    dev.clear_render_target_view(S.CLEAR_COLOR)
    dev.clear_depth_stencil_view(S.CLEAR_DEPTH)
    orc.begin_frame(sc.per_frame_cb, S.CLEAR_COLOR, S.CLEAR_DEPTH)
    dev.reset_stats()
    gp = dev.graphics_pipeline
    gp.ia.primitive_topology = L.PRIMITIVE_TOPOLOGY_TRIANGLELIST
    gp.rs.viewport.width, gp.rs.viewport.height = float(sc.width), float(sc.height)
    gp.rs.viewport.top_left_x = gp.rs.viewport.top_left_y = 0.0
    gp.rs.viewport.min_depth, gp.rs.viewport.max_depth = 0.0, 1.0
    gp.vs.p_constant_buffers[0] = sc.per_frame_cb
    for i, o in enumerate(sc.objects):
        tile_min_before = orc.tile_min_depths()
        gp.ia.input_layout = o.vertex_shader.in_vertex_size // 8
        gp.vs.output_register_count = o.vertex_shader.out_vertex_size // 128
        gp.vs.shader, gp.ps.shader = o.vertex_shader, o.pixel_shader
        gp.ia.p_index_buffer, gp.ia.p_vertex_buffer = o.index_buffer, o.vertex_buffer
        gp.vs.p_shader_resource_views[0] = gp.ps.p_shader_resource_views[0] = o.texture
        dev.draw_indexed(o.index_count)
        orc.draw(o.vertex_buffer, o.index_buffer, o.vertex_shader.vs_main, o.pixel_shader.ps_main, o.texture, staged=True)
        per_draw(i, o, tile_min_before)


@pytest.mark.parametrize("name", sorted(cases.FULL))
def test_production_path_full_size_against_live_reference(name):
    """BASELINE configs 1-4 (+ SUPREMATISM, the screenshot scene) at full resolution through the PRODUCTION kernels --
    debug_capture off, i.e. exactly what bench.py times -- compared with the live reference after EVERY draw: depth
    bit-exact, tile minima equal, colour within tolerance, and at the end Stats + the committed depth hash."""
    sc = cases.FULL[name]()
    if not have_ref(sc.width, sc.height):
        pytest.skip("oracle/_ref not available for this resolution")
    orc = _oracle(sc.width, sc.height)
    if not parity.host_vrsqrtps_matches_table(orc):
        pytest.skip("host vrsqrtps differs from the committed Intel table")
    with _device(sc.width, sc.height) as dev:
        def check(i, o, _):
            col, dep = dev.present()
            parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), f"{name} after draw {i} ({o.name})")
            assert np.array_equal(dev.debug_tile_min_depths(), orc.tile_min_depths()), f"{name}: tile minima differ after draw {i}"
        _render_production_draw_by_draw(dev, orc, sc, check)
        col, dep = dev.present()
        assert dev.stats() == orc.stats() == GOLDEN[name]["stats"]
    assert _fnv(dep) == GOLDEN[name]["depth_fnv"]
    from malevich_b200._lib import fnv64_words
    assert fnv64_words(dep) == GOLDEN[name]["depth_fnv"]  # the library's own hash helper (what bench.py prints) agrees


@pytest.mark.parametrize("name", sorted(cases.SMALL) + ["config1_toon_1280x720", "config2_ftm_1920x1080"])
def test_production_bin_lists_equal_reference_lists_minus_hiz_rejected_pairs(name):
    """The per-tile lists k_tile consumes in production (fused counting in k_geom / k_geom_clip / k_bin_big, scan, unordered
    fill, Hi-Z at binning time) are, per bin and as sorted sets, the reference's triangle_ids (main.c:950-962) minus the
    pairs its rasterizer rejects by Hi-Z (tri.max_depth < a_tile_min_depths[bin], main.c:1003-1010); a bin is on the
    work list iff it has a surviving pair (or is visited for write_tile's tile-minimum refresh only). Keys are mapped to the
    reference's ids through a debug-capture device rendering the same frame."""
    sc = {**cases.SMALL, **cases.FULL}[name]()
    if not have_ref(sc.width, sc.height):
        pytest.skip("oracle/_ref not available for this resolution")
    orc = _oracle(sc.width, sc.height)
    if not parity.host_vrsqrtps_matches_table(orc):
        pytest.skip("host vrsqrtps differs from the committed Intel table")
    # pass 1: key of every reference id, per draw (debug capture)
    keys_by_id = []
    with _device(sc.width, sc.height, debug_capture=True) as dbg:
        _render_production_draw_by_draw(dbg, orc, sc, lambda i, o, tm: keys_by_id.append(dbg.debug_keys()))
    # pass 2: production lists
    with _device(sc.width, sc.height) as dev:
        def check(i, o, tile_min_before):
            tris, _ = orc.staged_triangles()
            ids, rbins = orc.staged_bins()
            keys, gbins = dev.bin_lists()
            kmap = keys_by_id[i]
            assert len(kmap) == len(tris)
            got = {int(b["bin_index"]): np.sort(keys[int(b["num_triangles_upto"]):int(b["num_triangles_upto"]) + int(b["num_triangles_self"])]) for b in gbins}
            assert len(got) == len(gbins), "a bin is listed twice"
            assert np.all(np.diff(gbins["bin_index"].astype(np.int64)) > 0), "work list not in ascending bin order"
            ref_bins = set()
            for b in rbins:
                bi = int(b["bin_index"])
                ref_bins.add(bi)
                lst = ids[int(b["num_triangles_upto"]):int(b["num_triangles_upto"]) + int(b["num_triangles_self"])]
                keep = ~(tris["max_depth"][lst] < tile_min_before[bi])
                want = kmap[lst[keep]]
                if want.size:
                    assert bi in got and np.array_equal(got[bi], want), f"{name} draw {i} bin {bi}: list differs"
                else:
                    assert bi not in got or got[bi].size == 0, f"{name} draw {i} bin {bi}: pairs that Hi-Z rejects are listed"
            assert set(got) <= ref_bins, f"{name} draw {i}: bins on the work list that the reference leaves empty"
            assert int(sum(g.size for g in got.values())) == len(keys)
        _render_production_draw_by_draw(dev, orc, sc, check)
        assert dev.stats() == orc.stats()


def test_every_triangle_clipped_into_a_fan_of_three():
    """a draw whose triangles ALL cross two frustum planes (fan of 3 each) fills 3T overflow slots. The
    reference renders it (3T surviving triangles <= T + max(2T, 512), main.c:739-740); so must the device, bit-exact."""
    from malevich_b200 import scenes
    from malevich_b200.device import passthrough_ps, passthrough_vs
    W, H = 320, 200
    if not have_ref(W, H):
        pytest.skip("oracle/_ref not available")
    n = 512  # triangles (index count 1536: divisible by 8 and 3)
    rng = np.random.default_rng(11)
    vb = np.zeros((3 * n, 8), np.float32)
    for t in range(n):
        # passthrough_vs maps x,y in [0,1] to clip [-1,1] (w = 1): one vertex inside, two beyond the right AND the top plane,
        # clockwise/counter-clockwise chosen so that the triangle survives back-face culling (signed_area <= 0)
        cx, cy = rng.uniform(0.55, 0.9, 2)
        p = np.array([[cx, cy], [cx + 0.5, cy + 0.05], [cx + 0.05, cy + 0.5]], np.float32)
        z = np.float32(rng.uniform(0.2, 0.8))
        vb[3 * t:3 * t + 3, 0:2] = p
        vb[3 * t:3 * t + 3, 2] = z
        vb[3 * t:3 * t + 3, 3] = 1.0
        vb[3 * t:3 * t + 3, 4:7] = rng.uniform(0, 1, (3, 3)).astype(np.float32)
    ib = np.arange(3 * n, dtype=np.uint32)
    sc = scenes.Scene("fan3", W, H, [scenes.SceneObject(vb, ib, passthrough_vs, passthrough_ps, None, "fan3")], np.zeros((3, 4, 4), np.float32))
    orc = _oracle(W, H)
    orc.render(sc)
    st_ref = orc.stats()
    assert st_ref["assembled_triangle_count"] > 2 * n, f"test geometry no longer clips into fans of three: {st_ref}"
    for debug in (False, True):
        with _device(W, H, debug_capture=debug) as dev:
            scenes.render(dev, sc)
            col, dep = dev.present()
            assert dev.stats() == st_ref
        parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), f"fan3 debug={debug}")


def test_out_of_range_and_infinite_vertices_snap_like_x86():
    """the reference's (i32) casts of the snapped coordinates are x86 cvttsd2si -- 0x80000000 for NaN and for
    anything outside the i32 range (huge x/w, w = Inf) -- and such triangles are assembled, bounded and binned from those
    values. Staged comparison against the live reference on vertices that provoke every case."""
    from malevich_b200 import scenes
    from malevich_b200.device import passthrough_ps, passthrough_vs
    W, H = 320, 200
    if not have_ref(W, H):
        pytest.skip("oracle/_ref not available")
    big, inf = np.float32(3.0e7), np.float32(np.inf)
    tris = [
        [(0.2, 0.2, 0.5, 1.0), (0.8, 0.3, 0.5, 1.0), (0.4, 0.9, 0.5, 1.0)],          # ordinary (control)
        [(0.5, 0.5, 0.5, inf), (0.6, 0.5, 0.5, 1.0), (0.5, 0.6, 0.5, 1.0)],           # w = Inf: rw = 0, x*rw = 0
        [(0.5 * big, 0.5, 0.5, big), (0.6, 0.5, 0.5, 1.0), (0.5, 0.7, 0.5, 1.0)],      # huge but consistent
        [(0.3, 0.3, 0.5, 1e-30), (0.3, 0.3, 0.5, 1e-30), (0.3, 0.3, 0.5, 1e-30)],      # x/w far outside the i32 range
        [(inf, 0.5, 0.5, inf), (0.6, 0.5, 0.5, 1.0), (0.5, 0.7, 0.5, 1.0)],            # Inf * 0 = NaN
        [(0.1, 0.1, 0.5, 1.0), (0.1, 0.1, 0.5, 1.0), (0.1, 0.1, 0.5, 1.0)],            # degenerate point
        [(0.7, 0.2, 0.25, 1.0), (0.9, 0.2, 0.25, 1.0), (0.8, 0.4, 0.25, 1.0)],
        [(0.2, 0.7, 0.75, 1.0), (0.4, 0.7, 0.75, 1.0), (0.3, 0.9, 0.75, 1.0)],
    ]
    vb = np.zeros((24, 8), np.float32)
    for t, tri in enumerate(tris):
        for k, v in enumerate(tri):
            vb[3 * t + k, 0:4] = v
            vb[3 * t + k, 4:7] = (0.25 * k + 0.1, 0.5, 0.125 * t)
    sc = scenes.Scene("snap", W, H, [scenes.SceneObject(vb, np.arange(24, dtype=np.uint32), passthrough_vs, passthrough_ps, None, "snap")], np.zeros((3, 4, 4), np.float32))
    orc = _oracle(W, H)
    with np.errstate(all="ignore"):
        with _device(W, H, debug_capture=True) as dev:
            for draw_name, res in parity.render_both_staged(dev, orc, sc):
                assert parity.staged_ok(res), f"{draw_name}: {res}"
            col, dep = dev.present()
            assert dev.stats() == orc.stats()
        parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), "snap")
        with _device(W, H) as dev:
            scenes.render(dev, sc)
            col, dep = dev.present()
            assert dev.stats() == orc.stats()
        parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), "snap production")


def test_committed_gpu_frame_hashes_are_the_production_frames():
    """tests/golden/gpu_frames.json (what bench.py's `matches_golden.color` compares with, at every GPU count) holds the hashes
    of the frames the production path renders for the five BASELINE configs; tools/make_gpu_golden.py wrote them only after the
    frames passed the parity bars against the live reference. Here: the hashes still are those frames', and (where the reference
    is available) the frames still pass."""
    from malevich_b200 import scenes
    from malevich_b200._lib import fnv64_words
    committed = json.load(open(os.path.join(ROOT, "tests", "golden", "gpu_frames.json")))
    for cfg, key in ((1, "config1_toon_1280x720"), (2, "config2_ftm_1920x1080"), (3, "config3_emily_1920x1080"), (4, "config4_locomotive_3840x2160"), (5, "config5_synthetic_3840x2160")):
        sc = scenes.CONFIGS[cfg]()
        with _device(sc.width, sc.height) as dev:
            cl = dev.record(lambda: (dev.reset_stats(), scenes.render(dev, sc)))  # the way bench.py issues the frame
            cl.execute()
            col, dep = dev.present()
            assert dev.stats() == GOLDEN[key]["stats"]
            cl.release()
        assert fnv64_words(dep) == committed[key]["depth_fnv"] == GOLDEN[key]["depth_fnv"], key
        assert fnv64_words(col) == committed[key]["color_fnv"], key
        if cfg != 5 and have_ref(sc.width, sc.height):  # (config 5 against the live reference: test_config5_full_size_against_live_reference)
            orc = _oracle(sc.width, sc.height)
            if parity.host_vrsqrtps_matches_table(orc):
                orc.render(sc)
                parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), key)
