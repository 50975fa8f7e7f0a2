#!/usr/bin/env python3
"""bench.py -- frames/s of the draw pipeline (clear -> draw_indexed x N -> framebuffer) on 1..8 B200.

A "step" is ONE FRAME of the workload: colour+depth clear, every draw_indexed of the scene, resolve to
the row-major framebuffer (+ sort-first composite over NCCL when --gpus > 1).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 5]        # this repo's CUDA path
  python bench.py --impl reference ...                                    # the reference's own CPU pipeline (oracle/_ref)

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for how every field is derived.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "frames_per_s"
UNIT = "frames/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(scene, stats, num_draws):
    """SURVEY.md 8d per-unit table -> bytes per frame, split per stage."""
    px = scene.width * scene.height
    bins = (scene.width // 8) * (scene.height // 8)
    idx = sum(o.index_count for o in scene.objects)
    vtx = sum(o.vertex_buffer.shape[0] for o in scene.objects)
    tex = sum((o.texture.p_data.nbytes if o.texture is not None else 0) + 192 for o in scene.objects)
    tri, pairs, tdraws = stats["assembled_triangle_count"], stats["total_triangle_count_in_bins"], stats["active_bin_count"]
    per_stage = {
        "clear": 8 * px,
        "geometry": 4 * idx + 32 * vtx + tex + 216 * tri,          # inputs once + setup-record write
        "bin": 8 * pairs + 16 * bins * num_draws,                  # list write+read, headers write+read
        "tile": 216 * tri + 1032 * tdraws,                         # setup-record read + framebuffer tile read/write + tile-min
        "present": 4 * px,
    }
    return sum(per_stage.values()), per_stage


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_scene(config: int):
    from malevich_b200 import scenes
    return scenes.CONFIGS[config]()


def workload_name(config: int, scene) -> str:
    return {1: "config1 TOON", 2: "config2 FTM", 3: "config3 EMILY(stand-in sphere)+fullscreen radiance", 4: "config4 LOCOMOTIVE(stand-in torus knot)",
            5: "config5 synthetic 10M-triangle grid, 8 draws"}[config] + f" {scene.width}x{scene.height}"


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank: int):
    """The reference's own CPU pipeline (compiled from its sources into oracle/_ref), all host threads."""
    if rank != 0:
        return
    from oracle.ref_oracle import RefOracle, have
    scene = build_scene(args.config)
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, scene), "input_triangles": scene.input_triangles, "draws": len(scene.objects)}}
    if not have(scene.width, scene.height):
        base["unavailable"] = f"oracle/_ref/libmalevich_ref_{scene.width}x{scene.height}.so not built (needs /root/reference at build time)"
        print(json.dumps(base), flush=True)
        return
    orc = RefOracle(scene.width, scene.height, threads=os.cpu_count() or 1)
    cores = orc.max_threads()
    for _ in range(args.warmup):
        orc.render(scene)
    orc.prof_reset()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.render(scene)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    prof = {k: round(v / max(args.steps, 1), 3) for k, v in orc.prof().items()}
    fps = 1.0 / dt
    base.update({"value": fps, "ms_per_step": dt * 1e3, "mtri_per_s": scene.input_triangles * fps / 1e6, "gpix_per_s": scene.width * scene.height * fps / 1e9,
                 "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"{args.steps} full frames of the same workload, render() = clears + all draws (main.c:1265-1299)",
                                  "stage_ms": prof},
                 "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
    print(json.dumps(base), flush=True)


def cpu_baseline_sample(scene, frames: int = 2):
    from oracle.ref_oracle import RefOracle, have
    if not have(scene.width, scene.height):
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    orc = RefOracle(scene.width, scene.height, threads=os.cpu_count() or 1)
    orc.render(scene)  # warm-up frame
    best = 1e30
    for _ in range(frames):
        t0 = time.perf_counter()
        orc.render(scene)
        best = min(best, time.perf_counter() - t0)
    return {"value": 1.0 / best, "unit": UNIT, "cores": orc.max_threads(), "kind": "reference",
            "sample": f"best of {frames} full frames (after 1 warm-up) of the same workload on the host cores, reference render() compiled from its own sources"}


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from malevich_b200 import Device, scenes

    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene = build_scene(args.config)
    # sort-first split: by default one contiguous band of tile rows per rank (lets a rank skip geometry chunks outside its band)
    stripe = args.stripe if args.stripe > 0 else max(1, -(-(scene.height // 8) // world))
    dev = Device(scene.width, scene.height, cuda_device=local_rank, num_ranks=world, rank=rank, stripe_height_tiles=stripe)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))
    from malevich_b200 import partition

    def shard_bytes(nbytes: int) -> int:  # equal shards of a buffer, 16-byte granules, for the in-place all-gather of sharded uploads
        return partition.shard_bytes(nbytes, world)

    if multi:  # buffers padded to world x shard so that every rank's shard has the same size
        from malevich_b200 import _lib as L0
        for o in scene.objects:
            for arr, kind in ((o.vertex_buffer, L0.BUFFER_VERTEX), (o.index_buffer, L0.BUFFER_INDEX)):
                dev.adopt_buffer(arr, kind, shard_bytes(arr.nbytes) * world)
    scenes.upload(dev, scene)

    gather = None
    if multi:
        class _Raw:  # CUDA array interface over the library-owned gather buffer (no copy)
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        ptr, chunk = dev.composite_layout()
        gather = torch.as_tensor(_Raw(ptr, chunk * world), device=torch.device("cuda", local_rank))
        my_chunk = gather[rank * chunk:(rank + 1) * chunk]

    # Exchange step: by default the fused peer-memory composite (each rank resolves its tiles straight into every rank's
    # image over NVLink, include/malevich_b200.h); --composite nccl keeps pack -> ncclAllGather -> unpack. NCCL is always
    # the transport for the one-off handle exchange, the barriers and the max/sum reductions of the measurements.
    composite = "none"
    if multi:
        composite = args.composite
        if composite == "p2p":
            try:
                mine = torch.frombuffer(bytearray(dev.composite_peer_export()), dtype=torch.uint8).cuda()
                every = torch.empty(world * mine.numel(), dtype=torch.uint8, device="cuda")
                dist.all_gather_into_tensor(every, mine)
                blob = every.cpu().numpy().tobytes()
                n = mine.numel()
                dev.composite_peer_attach([blob[r * n:(r + 1) * n] for r in range(world)], same_process=False)
                ok = torch.ones(1, device="cuda")
            except Exception as e:  # noqa: BLE001 -- e.g. no peer access between the GPUs of this box
                print(f"[bench] rank {rank}: peer-memory composite unavailable ({e}); using ncclAllGather", file=sys.stderr)
                ok = torch.zeros(1, device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                composite = "nccl"

    # p2p: the exchange of frame f runs on the library's exchange stream underneath the drawing of frame f+1 (see
    # mlv_composite_broadcast_async): frame() draws frame f+1, joins the exchange of frame f and starts that of frame
    # f+1; drain() joins the last one. Every timed region ends with drain(), so K steps contain K complete exchanges.
    pending = [False]

    def drain():
        if pending[0]:
            dev.composite_join()
            pending[0] = False

    def frame():
        scenes.render(dev, scene)
        if composite == "p2p":
            drain()
            dev.composite_broadcast_async()
            pending[0] = True
        elif composite == "nccl":
            dev.composite_pack()
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gather, my_chunk)
            dev.composite_unpack()
        else:
            dev.resolve()

    def barrier():
        drain()
        dev.finish()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum_dict(d: dict) -> dict:
        if not multi:
            return d
        keys = sorted(d)
        t = torch.tensor([float(d[k]) for k in keys], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return {k: int(v) for k, v in zip(keys, t.tolist())}

    # ---- warm-up, stats of one frame (work counts for the algorithmic-bytes model)
    for _ in range(max(args.warmup, 3)):
        frame()
    dev.finish()
    dev.reset_stats()
    frame()
    stats_local = dev.stats()
    work = reduce_sum_dict(dev.work_counters())  # what the kernels really processed in that frame (Hi-Z at binning time removes the rest)
    # per-rank shares (a triangle is counted by the rank that owns its first tile row); the sum is the reference's Stats
    stats = reduce_sum_dict({k: stats_local[k] for k in ("assembled_triangle_count", "active_bin_count", "total_triangle_count_in_bins")})

    # ---- timed region: device-resident inputs, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = dev.kernel_launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        frame()
    drain()
    e1.record(stream)
    barrier()
    launches = dev.kernel_launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = reduce_max(e0.elapsed_time(e1) / args.steps)

    # ---- per-stage kernel times (CUDA events around every launch, same stream)
    dev.profile_begin()
    for _ in range(args.steps):
        frame()
    drain()
    prof = dev.profile_end()
    stage_ms = {k: v[0] / args.steps for k, v in prof.items()}
    stage_launches = {k: v[1] // args.steps for k, v in prof.items()}

    # ---- end to end through the public API with HOST buffers: every frame uploads every vertex/index buffer,
    # texture and the constant buffer from pinned host memory, renders, and reads the framebuffer back to the host
    def pinned_like(a: np.ndarray) -> np.ndarray:
        t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
        out = t.numpy().view(a.dtype).reshape(a.shape)
        out[...] = a
        out_keep.append(t)
        return out
    out_keep = []
    import ctypes as C
    from malevich_b200 import _lib as L
    lib = L.load()
    host_inputs = []
    h2d = 0
    for o in scene.objects:
        for arr, kind in ((o.vertex_buffer, L.BUFFER_VERTEX), (o.index_buffer, L.BUFFER_INDEX)):
            host_inputs.append((dev._buffer(arr, kind), pinned_like(arr)))
            h2d += arr.nbytes
    tex_seen = {}
    for o in scene.objects:
        if o.texture is not None and id(o.texture) not in tex_seen:
            tex_seen[id(o.texture)] = (dev._texture(o.texture), pinned_like(o.texture.p_data))
    h2d += sum(p.nbytes for _, p in tex_seen.values()) + 192
    colors_host = pinned_like(np.zeros((scene.height, scene.width), np.uint32))
    d2h = colors_host.nbytes

    colors_host2 = pinned_like(colors_host)
    e2e_step = [0]

    # N > 1: the geometry is replicated, so every byte crosses PCIe ONCE per frame: rank r uploads shard r of every
    # buffer over its own link, an in-place ncclAllGather on a side stream replicates the shards over NVLink, and the
    # buffer is marked complete on that stream (draws that bind it wait for exactly that). The composited frame is read
    # back by rank 0 on the read-back stream, one frame behind (the exchange of frame f overlaps frame f+1).
    if multi and composite == "p2p":
        class _RawBuf:
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        cuda_dev = torch.device("cuda", local_rank)
        copy_stream = torch.cuda.ExternalStream(dev.copy_stream, device=cuda_dev)
        ag_stream = torch.cuda.Stream(device=cuda_dev)
        sharded = []
        for h, p in host_inputs:
            sb = shard_bytes(p.nbytes)
            full = torch.as_tensor(_RawBuf(int(lib.mlv_buffer_device_ptr(h)), sb * world), device=cuda_dev)
            lo, count = partition.shard_range(p.nbytes, world, rank)
            sharded.append((h, p.ctypes.data + lo, lo, count, full, full[rank * sb:(rank + 1) * sb], torch.cuda.Event()))

    def frame_e2e_sharded():
        for h, host_ptr, lo, n, full, mine, ev in sharded:
            L.check(lib.mlv_update_buffer_range(dev._h, h, lo, C.c_void_p(host_ptr), n))
            ev.record(copy_stream)
            ag_stream.wait_event(ev)
            with torch.cuda.stream(ag_stream):
                dist.all_gather_into_tensor(full, mine)
            L.check(lib.mlv_buffer_mark_updated(dev._h, h, C.c_void_p(ag_stream.cuda_stream)))
        scenes.render(dev, scene)
        had = pending[0]
        drain()  # join the exchange of the previous frame
        if had and rank == 0:
            dev.present_wait()
            dev.composite_readback_async(colors_host if e2e_step[0] % 2 == 0 else colors_host2)
            e2e_step[0] += 1
        dev.composite_broadcast_async()
        pending[0] = True

    def frame_e2e():
        if multi and composite == "p2p":
            return frame_e2e_sharded()
        for h, p in host_inputs:
            L.check(lib.mlv_update_buffer(dev._h, h, p.ctypes.data_as(C.c_void_p), p.nbytes))
        if multi:
            frame()
            drain()
            dev.finish()  # composite result is in the resolved image; read it back below
            _readback_multi()
        else:
            # double-buffered present: the host takes delivery of frame f-1 while frame f is already queued, and the
            # device-to-host copy of frame f overlaps the uploads of frame f+1 (PCIe is full duplex)
            scenes.render(dev, scene)
            dev.present_wait()
            dev.present_async(colors_host if e2e_step[0] % 2 == 0 else colors_host2)
            e2e_step[0] += 1

    def _readback_multi():
        torch.cuda.synchronize()
        n = scene.width * scene.height * 4
        class _Raw2:
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        src = torch.as_tensor(_Raw2(dev.resolved_color_ptr(), n), device=torch.device("cuda", local_rank))
        torch.from_numpy(colors_host.view(np.uint8).reshape(-1)).copy_(src, non_blocking=False)

    for _ in range(2):
        frame_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame_e2e()
    if multi and composite == "p2p":  # deliver the last frame too: K steps = K frames uploaded, drawn, exchanged and read back
        drain()
        if rank == 0:
            dev.present_wait()
            dev.composite_readback_async(colors_host if e2e_step[0] % 2 == 0 else colors_host2)
            last_e2e_image = colors_host if e2e_step[0] % 2 == 0 else colors_host2
    barrier()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3 / args.steps)
    e2e_check = None
    if multi and composite == "p2p":
        frame()   # every rank takes part in the exchange
        barrier()
    if multi and composite == "p2p" and rank == 0:
        # the frame delivered end to end (sharded uploads, asynchronous exchange) is the frame the resident path renders
        ref_img = torch.as_tensor(_RawBuf(dev.resolved_color_ptr(), scene.width * scene.height * 4), device=cuda_dev).cpu().numpy().view(np.uint32).reshape(scene.height, scene.width)
        e2e_check = bool(np.array_equal(ref_img, last_e2e_image))
    tex_bytes = sum(p.nbytes for _, p in tex_seen.values())

    if rank == 0:
        peak, peak_src = load_peaks()
        b_alg, b_stage = algorithmic_bytes(scene, stats, len(scene.objects))
        fps = 1e3 / ms
        # dominant kernel = the kernel with the largest summed time over the step; its algorithmic bytes (SURVEY.md 8d per-unit
        # figures x the units it processed) over its own CUDA-event time
        kernel_ms = {"k_front": stage_ms["geometry"], "k_back": stage_ms["geometry_back"], "k_tile": stage_ms["tile"], "k_front_clip": stage_ms["clip"], "k_vertex": stage_ms["vertex_cache"],
                     "k_bin_scan": stage_ms["bin_scan"], "k_bin_fill": stage_ms["bin_fill"]}
        kernel_launches = {"k_front": stage_launches["geometry"], "k_back": stage_launches["geometry_back"], "k_tile": stage_launches["tile"], "k_front_clip": stage_launches["clip"], "k_vertex": stage_launches["vertex_cache"],
                           "k_bin_scan": stage_launches["bin_scan"], "k_bin_fill": stage_launches["bin_fill"]}
        # Algorithmic bytes of a kernel = SURVEY.md 8d's per-unit figures x the units the kernel REALLY processed (device work
        # counters): Stats count every assembled triangle and pair like the reference, but a triangle that Hi-Z rejects at
        # binning time writes no record and a tile without surviving pairs is not visited, so those units move no bytes.
        idx_bytes = 4 * sum(o.index_count for o in scene.objects)
        vtx_bytes = 32 * sum(o.vertex_buffer.shape[0] for o in scene.objects)
        tri_in = sum(o.index_count // 3 for o in scene.objects)
        kernel_bytes = {"k_front": (idx_bytes + vtx_bytes + 16 * tri_in) / world,                           # inputs once + 16 B of bounds per input triangle
                        "k_back": (16 * tri_in + 216 * work["records_written"]) / world,                    # bounds read + record write
                        "k_tile": (216 * work["records_written"] + 1032 * work["tiles_visited"] + 4 * work["pairs_listed"]) / world,  # record read + tile read/write + list read
                        "k_front_clip": 0, "k_vertex": vtx_bytes / world, "k_bin_scan": 16 * (scene.width // 8) * (scene.height // 8) * len(scene.objects) / world,
                        "k_bin_fill": (16 * stats["assembled_triangle_count"] + 4 * work["pairs_listed"]) / world}
        dom = max(("k_front", "k_back", "k_tile"), key=kernel_ms.get)  # the kernels that carry the geometry and record streams; the others move < 5 % of the bytes

        def kernel_roofline(k):
            n = max(kernel_launches[k], 1)
            ms_l, b_l = kernel_ms[k] / n, kernel_bytes[k] / n
            a = b_l / (ms_l * 1e-3) / 1e9 if ms_l > 0 else 0.0
            return {"achieved": a, "frac": a / peak, "algorithmic_bytes_per_launch": b_l, "ms_per_launch": ms_l, "launches_per_step": kernel_launches[k]}
        dom_r = kernel_roofline(dom)
        achieved, dom_bytes_per_launch, dom_ms_per_launch, dom_launches = dom_r["achieved"], dom_r["algorithmic_bytes_per_launch"], dom_r["ms_per_launch"], dom_r["launches_per_step"]
        # measured DRAM bytes per launch of that kernel from the committed ncu capture (same workload, 1 GPU), else null
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic_config5_n1.json")
        if args.config == 5 and world == 1 and os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)["kernels"].get(dom, {}).get("dram_bytes_per_launch")
        out = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "mtri_per_s": scene.input_triangles * fps / 1e6, "gpix_per_s": scene.width * scene.height * fps / 1e9,
            "config": {"workload": workload_name(args.config, scene), "input_triangles": scene.input_triangles, "draws": len(scene.objects),
                       "assembled_triangles": stats["assembled_triangle_count"], "tri_tile_pairs": stats["total_triangle_count_in_bins"],
                       "tile_draws": stats["active_bin_count"], "parallelism": f"sort-first x{world}, stripe {stripe} tile rows, composite {composite}" if multi else "single GPU",
                       "l2": "no flush: per-frame working set (inputs + per-draw setup records) >> 126 MB L2" if args.config == 5 else "no flush"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes_per_launch,
                         "ms_per_launch": dom_ms_per_launch, "launches_per_step": dom_launches,
                         "units_per_step": {"records_written": work["records_written"], "tiles_visited": work["tiles_visited"], "pairs_listed": work["pairs_listed"]},
                         "note": "algorithmic bytes = SURVEY 8d per-unit figures x the units the kernel really processed (mlv_work_counters), averaged over its launches of a frame; "
                                 "the launch time is its CUDA-event bracket (no overlap with neighbours)"},
            "roofline_kernels": {k: kernel_roofline(k) for k in ("k_front", "k_back", "k_tile", "k_vertex", "k_bin_fill")},
            "roofline_frame": {"algorithmic_bytes_per_frame": b_alg, "achieved": b_alg / (ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                               "frac": b_alg / (ms * 1e-3) / 1e9 / world / peak, "bytes_by_stage": b_stage,
                               "note": "B_alg of the REFERENCE algorithm for this frame (SURVEY 8d: every assembled triangle's record written and read, every touched tile-draw loaded and "
                                       "stored) over the measured frame time: delivered work per second, not DRAM throughput -- it exceeds the peak when Hi-Z at binning time removes work"},
            "stage_ms_per_step": {k: round(v, 4) for k, v in stage_ms.items()},
            "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d - tex_bytes), "d2h_bytes_per_step": int(d2h),
                    "note": ("every vertex/index buffer + constant buffer re-uploaded from pinned host memory each frame; framebuffer read back to pinned host memory each frame (double-buffered: the host collects frame f-1 while frame f is queued); textures stay resident"
                             if not (multi and composite == "p2p") else
                             "whole job: every vertex/index buffer crosses PCIe once per frame -- rank r uploads shard r of each buffer from pinned host memory, an in-place ncclAllGather replicates it over NVLink -- and rank 0 reads the composited frame back to pinned host memory (one frame behind: the exchange of frame f overlaps frame f+1); textures stay resident"),
                    **({"image_matches_resident_path": e2e_check} if e2e_check is not None else {})},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if not multi and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample(scene)
        print(json.dumps(out), flush=True)
    dev.close()
    if multi:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--stripe", type=int, default=0, help="stripe height in tile rows for the sort-first split (0 = one contiguous band per rank)")
    ap.add_argument("--composite", default="p2p", choices=["p2p", "nccl"], help="multi-GPU exchange step: fused peer-memory broadcast (default) or pack + ncclAllGather + unpack")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 10:
            args.steps = 10  # bounded: a config-5 frame is seconds of CPU work
        if args.warmup > 2:
            args.warmup = 2
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun when called plainly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29541",
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its version
    # banner there) are pointed at stderr; the line itself goes to a private duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = json_out
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
