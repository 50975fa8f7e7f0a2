#!/usr/bin/env python3
"""bench.py -- frames/s of the draw pipeline (clear -> draw_indexed x N -> framebuffer) on 1..8 B200.

A "step" is ONE FRAME of the workload: colour+depth clear, every draw_indexed of the scene, resolve to
the row-major framebuffer (+ sort-first composite over NVLink when --gpus > 1).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 5]        # this repo's CUDA path
  python bench.py --impl reference ...                                    # the reference's own CPU pipeline (oracle/_ref)

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for how every field is derived.
"""
from __future__ import annotations

import argparse
import csv
import hashlib
import json
import os
import shutil
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
    """SURVEY.md 8d per-unit table -> bytes per frame of the REFERENCE algorithm, split per stage."""
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


WORKLOADS = {1: "config1 TOON", 2: "config2 FTM", 3: "config3 EMILY(stand-in sphere)+fullscreen radiance", 4: "config4 LOCOMOTIVE(stand-in torus knot)",
             5: "config5 synthetic 10M-triangle grid, 8 draws"}
GOLDEN_KEYS = {1: "config1_toon_1280x720", 2: "config2_ftm_1920x1080", 3: "config3_emily_1920x1080", 4: "config4_locomotive_3840x2160", 5: "config5_synthetic_3840x2160"}


def workload_config(config: int, scene) -> dict:
    """The `config` object: identical for this repo's arm and the reference arm (what was rendered, nothing derived)."""
    return {"workload": f"{WORKLOADS[config]} {scene.width}x{scene.height}", "input_triangles": scene.input_triangles, "draws": len(scene.objects),
            "width": scene.width, "height": scene.height}


def golden_hashes(config: int) -> dict:
    """Committed frame hashes of the workload: the reference's depth image (tests/golden/golden.json, made from oracle/_ref) and
    this repo's own colour image (tests/golden/gpu_frames.json: written on a GPU box by tools/make_gpu_golden.py only after
    the frame passed the colour tolerance against the live reference; colour is <= 1/255, not bit-exact, so the reference's
    own colour hash cannot be matched)."""
    out = {"depth_fnv": None, "color_fnv": None}
    try:
        out["depth_fnv"] = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))[GOLDEN_KEYS[config]]["depth_fnv"]
    except Exception:
        pass
    try:
        out["color_fnv"] = json.load(open(os.path.join(ROOT, "tests", "golden", "gpu_frames.json")))[GOLDEN_KEYS[config]]["color_fnv"]
    except Exception:
        pass
    return out


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank: int):
    """The reference's own CPU pipeline (compiled from its sources into oracle/_ref), all host threads."""
    if rank != 0:
        return
    from oracle.ref_oracle import RefOracle, have
    scene = build_scene(args.config)
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic", "config": workload_config(args.config, scene)}
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


def measure_traffic(config: int, kernel: str, timeout_s: int = 150):
    """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the launches of ONE frame
    of the workload on one GPU), measured now by a child process under ncu. Never a timing: only byte counters are read.
    Returns (bytes_per_launch or None, how)."""
    ncu = shutil.which("ncu")
    if ncu is None:
        return None, "ncu not on PATH"
    with tempfile.TemporaryDirectory() as tmp:
        log = os.path.join(tmp, "traffic.csv")
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:^{kernel}", "--csv", "--log-file", log,
               sys.executable, os.path.join(ROOT, "tools", "profile_rank.py"), "1", "0", "2", str(config)]
        try:
            r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout_s)
            if r.returncode != 0 or not os.path.exists(log):
                return None, f"ncu exit code {r.returncode}"
            lines = open(log).read().splitlines()
            start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
            rows = list(csv.reader(lines[start:]))
            h = rows[0]
            idi, vi = h.index("ID"), h.index("Metric Value")
            per = {}
            for row in rows[1:]:
                if len(row) > vi:
                    per[row[idi]] = per.get(row[idi], 0.0) + float(row[vi].replace(",", "") or 0)
            ids = sorted(per, key=int)
            if not ids:
                return None, "no launch of the kernel captured"
            frame = ids[len(ids) // 2:]  # the child renders two frames: the second one (warm arenas)
            return sum(per[i] for i in frame) / len(frame), f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the {len(frame)} launches of one frame (child process, measured in this run)"
        except Exception as e:  # noqa: BLE001
            return None, f"ncu capture failed: {type(e).__name__}"


def kernels_source_hash() -> str:
    h = hashlib.sha256()
    for name in ("kernels.cuh", "shaders.cuh", "mlv_internal.cuh", "malevich_b200.cu"):
        h.update(open(os.path.join(ROOT, "malevich_b200", "csrc", name), "rb").read())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    import ctypes as C

    import torch
    import torch.distributed as dist
    from malevich_b200 import Device, partition, scenes
    from malevich_b200 import _lib as L

    torch.cuda.set_device(local_rank)
    cuda_dev = torch.device("cuda", local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=cuda_dev)
    scene = build_scene(args.config)
    lib = L.load()
    # sort-first split: by default one contiguous band of tile rows per rank (lets a rank skip geometry chunks outside its band)
    stripe = args.stripe if args.stripe > 0 else max(1, -(-(scene.height // 8) // world))
    dev = Device(scene.width, scene.height, cuda_device=local_rank, num_ranks=world, rank=rank, stripe_height_tiles=stripe)
    stream = torch.cuda.ExternalStream(dev.stream, device=cuda_dev)

    class _Raw:  # CUDA array interface over library-owned device memory (no copy)
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}

    def as_tensor(ptr, nbytes):
        return torch.as_tensor(_Raw(ptr, nbytes), device=cuda_dev)

    scenes.upload(dev, scene)

    # ---- exchange step: by default the asynchronous peer-memory composite (each rank's band travels into every rank's image
    # over NVLink underneath the next frame, include/malevich_b200.h); --composite nccl keeps pack -> ncclAllGather -> unpack.
    # NCCL is always the transport for the one-off handle exchange, the barriers and the reductions of the measurements.
    composite = "none"
    gather = my_chunk = None
    if multi:
        composite = args.composite
        ptr, chunk = dev.composite_layout()
        gather = as_tensor(ptr, chunk * world)
        my_chunk = gather[rank * chunk:(rank + 1) * chunk]
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

    # ---- the frame, recorded once (D3D11 deferred-context pattern: mlv_begin/finish/execute_command_list). Replaying it is one
    # CUDA-graph launch; `--immediate` issues the calls one by one instead (what round 1 measured).
    def record_frame():
        dev.reset_stats()
        scenes.render(dev, scene)
        if not multi:
            dev.resolve()
    scenes.render(dev, scene)  # sizes every arena before the recording
    dev.finish()
    # The frame is recorded TWICE: a recording that opens with a full clear alternates between the two tiled framebuffers of the
    # device (include/malevich_b200.h, command lists), so replaying the two lists in turn lets frame f+1 start while the exchange
    # (N > 1) or the asynchronous present (resolve / pack + copy on the read-back stream) of frame f still reads the other one.
    n_lists = 2
    frame_lists = None if args.immediate else [dev.record(record_frame) for _ in range(n_lists)]
    # the same frame without the device-resident resolve: what the end-to-end loops replay (the present resolves and copies
    # on its own, and a recording that writes the resolved image has to wait for a read-back of that image still in flight)
    def record_draws():
        dev.reset_stats()
        scenes.render(dev, scene)
    draws_lists = None if args.immediate else (frame_lists if multi else [dev.record(record_draws) for _ in range(n_lists)])
    frame_no = [0]

    pending = [False]  # p2p: an exchange has been started and not yet joined

    def drain():
        if pending[0]:
            dev.composite_join()
            pending[0] = False

    def render_frame():
        if frame_lists is not None:
            frame_lists[frame_no[0] % n_lists].execute()
            frame_no[0] += 1
        else:
            record_frame()

    def frame():
        """clear + all draws + resolve (N = 1) / + exchange (N > 1). p2p: the exchange of frame f runs on the library's exchange
        stream underneath frame f+1; every timed region ends with drain(), so K steps contain K complete exchanges."""
        render_frame()
        if composite == "p2p":
            drain()
            dev.composite_broadcast_async()
            pending[0] = True
        elif composite == "nccl":
            dev.composite_pack()
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gather, my_chunk)
            dev.composite_unpack()

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

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        frame()
    barrier()
    # work counts of one frame (the recording resets Stats at its start, like memset(&stats, 0) in render(), main.c:1268)
    frame()
    drain()
    stats_local = dev.stats()
    work = reduce_sum_dict(dev.work_counters())  # what the kernels really processed in that frame (Hi-Z at binning time removes the rest)
    # per-rank shares (a triangle is counted by the rank that owns its first tile row); the sum is the reference's Stats
    stats = reduce_sum_dict({k: stats_local[k] for k in ("assembled_triangle_count", "active_bin_count", "total_triangle_count_in_bins")})

    # ---- the frame the timed path renders, hashed (outside every timed region) and compared with the committed frames
    hashes = None
    colors = np.empty((scene.height, scene.width), np.uint32)
    if rank == 0:
        if multi:
            colors[...] = as_tensor(dev.resolved_color_ptr(), colors.nbytes).cpu().numpy().view(np.uint32).reshape(colors.shape)
            depth_fnv = None  # depth is not exchanged between the ranks
        else:
            depths = np.empty((scene.height, scene.width), np.float32)
            dev.present_into(colors, depths)
            depth_fnv = L.fnv64_words(depths)
        want = golden_hashes(args.config)
        hashes = {"color": L.fnv64_words(colors), "depth": depth_fnv, "golden_color": want["color_fnv"], "golden_depth": want["depth_fnv"]}
    barrier()

    # ---- timed region: device-resident inputs, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)
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

    # ---- per-stage kernel times (immediate mode: CUDA events around every launch, all on one stream, nothing overlaps)
    dev.profile_begin()
    for _ in range(args.steps):
        record_frame()
        if multi:
            dev.composite_pack()
    prof = dev.profile_end()
    stage_ms = {k: v[0] / args.steps for k, v in prof.items()}
    stage_launches = {k: v[1] // args.steps for k, v in prof.items()}

    # ---- end to end through the public API with HOST buffers, two ways:
    # (1) `e2e`: the reference's own frame loop (main.c:1587-1602): geometry and textures are loaded once (init(), main.c:1317-1420),
    #     every frame update() rewrites the 192-byte PerFrameCB (main.c:1595-1597), render() draws, the frame is presented. Here:
    #     the constant buffer of every draw is replaced from host memory (mlv_command_list_set_constants), the recorded frame is
    #     replayed, the framebuffer is read back to pinned host memory.
    # (2) `e2e_streaming`: a host that streams its geometry -- every vertex and index byte re-uploaded from pinned host memory each
    #     frame (one vertex slab + one index slab addressed with start index / base vertex; at N > 1 each rank uploads its shard
    #     over its own PCIe link and an in-place ncclAllGather replicates it over NVLink), plus the same read-back.
    keep = []

    def pinned_like(a: np.ndarray) -> np.ndarray:
        t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
        out = t.numpy().view(a.dtype).reshape(a.shape)
        out[...] = a
        keep.append(t)
        return out
    host_frames = [pinned_like(np.zeros((scene.height, scene.width), np.uint32)) for _ in range(2)]
    d2h = host_frames[0].nbytes
    e2e_n = [0]
    cb_host = np.ascontiguousarray(scene.per_frame_cb, dtype=np.float32)

    # N > 1: the frame is composed IN HOST MEMORY (args.e2e_composite == "host", default). Its destination is the host, so the
    # ranks exchange nothing: every rank copies the rows it owns into ONE shared pinned frame (a /dev/shm mapping page-locked by
    # each rank, malevich_b200/hostframe.py) over its own PCIe link -- N links carry the 33 MB instead of rank 0's alone. Rank 0
    # is the consumer (the process that would blit it, main.c:1301-1306): it takes frame f once every rank has published its rows
    # of f, one frame behind the frame being queued. "device": compose on the device with the peer-memory exchange and let rank
    # 0 read the whole image back (what this bench measured before; kept for comparison).
    shared = None
    e2e_composite = "single GPU"
    if multi:
        e2e_composite = args.e2e_composite if composite == "p2p" else "device"
        if e2e_composite == "host":
            from malevich_b200.hostframe import SharedHostFrames
            ok = torch.ones(1, device="cuda")
            name = f"mlv_frames_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}"

            def attempt(step):  # every rank reaches every collective below whatever fails on it
                try:
                    step()
                except Exception as e:  # noqa: BLE001 -- e.g. /dev/shm too small, or the mapping cannot be page-locked
                    print(f"[bench] rank {rank}: shared host frame unavailable ({e}); rank 0 reads the composited frame back", file=sys.stderr)
                    ok.zero_()

            def create():
                nonlocal shared
                shared = SharedHostFrames(name, scene.height, scene.width, world, rank, slots=3, create=True)

            def attach():
                nonlocal shared
                if rank != 0:
                    shared = SharedHostFrames(name, scene.height, scene.width, world, rank, slots=3)
                dev.register_host_memory(shared.address, shared.nbytes)
                shared.reset()
            if rank == 0:
                attempt(create)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # (also the barrier: the mapping exists, or nobody goes on)
            if ok.item() != 0:
                attempt(attach)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                e2e_composite = "device"
    taken = [None]

    def deliver():  # hand the finished frame to the host: double-buffered, the copy of frame f overlaps frame f+1
        f = e2e_n[0]
        if multi and e2e_composite == "host":
            if f > 0:
                dev.present_wait()       # this rank's rows of frame f-1 are in host memory
                shared.publish(f)
                if rank == 0:
                    taken[0] = shared.take(f - 1)  # ... and so are everybody else's: frame f-1 is whole
            dev.present_owned_rows_async(shared.slot_for_next())
        elif multi:
            # exchange and read-back of THIS frame right behind its draws: the next frame's draws (the other tiled framebuffer of
            # the pair) are queued while rank 0's 33 MB copy is in flight, so the period is max(render, read-back) + exchange.
            drain()
            dev.composite_broadcast_async()
            dev.composite_join()
            if rank == 0:
                dev.present_wait()  # the previous frame has arrived in the other host buffer
                dev.composite_readback_async(host_frames[f % 2])
        else:
            dev.present_wait()
            dev.present_async(host_frames[f % 2])
        e2e_n[0] += 1

    def flush_e2e():  # deliver the last frame too: K steps = K frames drawn, composed and in host memory
        f = e2e_n[0]
        if multi and e2e_composite == "host":
            if f > 0:
                dev.present_wait()
                shared.publish(f)
                if rank == 0:
                    taken[0] = shared.take(f - 1)
        elif multi:
            drain()
            if rank == 0:
                dev.present_wait()
        dev.finish()

    def last_delivered():
        if multi and e2e_composite == "host":
            return taken[0]
        return host_frames[(e2e_n[0] - 1) % 2]

    def frame_e2e_static():
        if draws_lists is not None:
            dl = draws_lists[frame_no[0] % n_lists]
            frame_no[0] += 1
            dl.set_constants(cb_host)  # update(): the camera of this frame
            dl.execute()
        else:
            record_draws()  # (the Python mirror sends the constant buffer with every draw)
        deliver()

    def time_e2e(step, frames):
        """Wall clock over `frames` frames, the last one included: the clock stops when this rank holds (rank 0: when the host
        frame holds every rank's rows of) the last frame -- the pipeline's drain (one render + one read-back) is inside the timed
        region, which is why at least `frames` >= --steps frames are timed: over 20 frames the drain alone is 5-10 % of the figure."""
        for _ in range(2):
            step()
        flush_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(frames):
            step()
        flush_e2e()
        t1 = time.perf_counter()
        barrier()
        return reduce_max((t1 - t0) * 1e3 / frames)

    e2e_streaming = None
    if not multi or composite == "p2p":
        e2e_frames = max(args.steps, 200)
        e2e_static_ms = time_e2e(frame_e2e_static, e2e_frames)
        static_ok = bool(np.array_equal(last_delivered(), colors)) if rank == 0 else None
        e2e = {"value": 1e3 / e2e_static_ms, "unit": UNIT, "ms_per_step": e2e_static_ms, "h2d_bytes_per_step": 192 * len(scene.objects), "d2h_bytes_per_step": int(d2h),
               "image_matches_resident_path": static_ok, "frames_timed": e2e_frames,
               "note": "the reference's frame loop (main.c:1587-1602): geometry and textures resident (loaded once by init()), per frame the PerFrameCB of every draw replaced from host memory "
                       "(update(), main.c:1595-1597), the recorded frame replayed, the framebuffer read back to pinned host memory (double-buffered: the host collects frame f-1 while frame f is queued"
                       + ("; composed in host memory: every rank copies the rows it owns into one shared pinned frame over its own PCIe link, rank 0 takes the frame when all rows have arrived)" if e2e_composite == "host"
                          else "; composed on the device, rank 0 reads the composited frame)" if multi else ")"),
               "composite": e2e_composite}

        # ---- streaming geometry: slabs
        vtx_counts = [o.vertex_buffer.shape[0] for o in scene.objects]
        base_vertex = np.concatenate([[0], np.cumsum(vtx_counts)[:-1]]).astype(np.int64)
        start_index = np.concatenate([[0], np.cumsum([o.index_count for o in scene.objects])[:-1]]).astype(np.int64)
        vb_slab = pinned_like(np.concatenate([np.ascontiguousarray(o.vertex_buffer, dtype=np.float32) for o in scene.objects]))
        ib_slab = pinned_like(np.concatenate([np.ascontiguousarray(o.index_buffer, dtype=np.uint32) for o in scene.objects]))
        slabs = []
        for arr, kind in ((vb_slab, L.BUFFER_VERTEX), (ib_slab, L.BUFFER_INDEX)):
            sb = partition.shard_bytes(arr.nbytes, world)
            h = dev.adopt_buffer(arr, kind, sb * world)
            lo, count = partition.shard_range(arr.nbytes, world, rank)
            full = as_tensor(int(lib.mlv_buffer_device_ptr(h)), sb * world)
            slabs.append((h, arr, lo, count, full, full[rank * sb:(rank + 1) * sb], torch.cuda.Event()))
        copy_stream = torch.cuda.ExternalStream(dev.copy_stream, device=cuda_dev)
        ag_stream = torch.cuda.Stream(device=cuda_dev)

        def record_slab_frame():
            dev.reset_stats()
            dev.clear_render_target_view(scenes.CLEAR_COLOR)
            dev.clear_depth_stencil_view(scenes.CLEAR_DEPTH)
            gp = dev.graphics_pipeline
            gp.vs.p_constant_buffers[0] = scene.per_frame_cb
            gp.ia.p_vertex_buffer, gp.ia.p_index_buffer = vb_slab, ib_slab
            for o, s0, b0 in zip(scene.objects, start_index, base_vertex):
                gp.vs.shader, gp.ps.shader = o.vertex_shader, o.pixel_shader
                gp.vs.p_shader_resource_views[0] = gp.ps.p_shader_resource_views[0] = o.texture
                dev.draw_indexed(o.index_count, start_index_location=int(s0), base_vertex_location=int(b0))
        record_slab_frame()  # sizes the arenas for the slab draws
        dev.finish()
        slab_list = None if args.immediate else dev.record(record_slab_frame)

        def frame_e2e_streaming():
            for h, arr, lo, count, full, mine, ev in slabs:
                if multi:
                    L.check(lib.mlv_update_buffer_range(dev._h, h, lo, C.c_void_p(arr.ctypes.data + lo), count))
                    ev.record(copy_stream)
                    ag_stream.wait_event(ev)
                    with torch.cuda.stream(ag_stream):
                        dist.all_gather_into_tensor(full, mine)
                    L.check(lib.mlv_buffer_mark_updated(dev._h, h, C.c_void_p(ag_stream.cuda_stream)))
                else:
                    L.check(lib.mlv_update_buffer(dev._h, h, arr.ctypes.data_as(C.c_void_p), arr.nbytes))
            if slab_list is not None:
                slab_list.execute()
            else:
                record_slab_frame()
            deliver()
        e2e_streaming_ms = time_e2e(frame_e2e_streaming, max(args.steps, 40))
        streaming_ok = bool(np.array_equal(last_delivered(), colors)) if rank == 0 else None
        e2e_streaming = {"value": 1e3 / e2e_streaming_ms, "unit": UNIT, "ms_per_step": e2e_streaming_ms, "h2d_bytes_per_step": int(vb_slab.nbytes + ib_slab.nbytes + 192 * len(scene.objects)),
                         "d2h_bytes_per_step": int(d2h), "image_matches_resident_path": streaming_ok, "frames_timed": max(args.steps, 40),
                         "note": "a host that streams its geometry: one vertex slab + one index slab (draws address them with start index / base vertex) re-uploaded from pinned host memory every frame"
                                 + (", each rank uploading its shard over its own PCIe link and an in-place ncclAllGather replicating it over NVLink (2 collectives per frame)" if multi else "")
                                 + "; same read-back as e2e; textures stay resident"}
    else:
        # --composite nccl: e2e = static geometry with a blocking read-back by rank 0
        def frame_e2e_nccl():
            if frame_lists is not None:
                frame_lists[frame_no[0] % n_lists].set_constants(cb_host)
            frame()
            dev.finish()
            if rank == 0:
                torch.from_numpy(host_frames[0].view(np.uint8).reshape(-1)).copy_(as_tensor(dev.resolved_color_ptr(), d2h))
        for _ in range(2):
            frame_e2e_nccl()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            frame_e2e_nccl()
        barrier()
        e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3 / args.steps)
        e2e = {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 192 * len(scene.objects), "d2h_bytes_per_step": int(d2h),
               "note": "static geometry, per-frame constant buffer, pack + ncclAllGather + unpack, blocking read-back by rank 0"}

    if rank == 0:
        peak, peak_src = load_peaks()
        b_alg, b_stage = algorithmic_bytes(scene, stats, len(scene.objects))
        fps = 1e3 / ms
        # Per-kernel roofline: algorithmic bytes = SURVEY.md 8d's per-unit figures x the units the kernel REALLY processed (device
        # work counters: Stats count every assembled triangle and pair like the reference, but a triangle that Hi-Z rejects at
        # binning time writes no record and a tile without surviving pairs is not visited), over the kernel's own CUDA-event time.
        kstage = {"k_front": "geometry", "k_back": "geometry_back", "k_tile": "tile", "k_front_clip": "clip", "k_vertex": "vertex_cache", "k_bin_scan": "bin_scan", "k_fill": "bin_fill"}
        kernel_ms = {k: stage_ms[s] for k, s in kstage.items()}
        kernel_launches = {k: stage_launches[s] for k, s in kstage.items()}
        idx_bytes = 4 * sum(o.index_count for o in scene.objects)
        vtx_bytes = 32 * sum(o.vertex_buffer.shape[0] for o in scene.objects)
        tri_in = sum(o.index_count // 3 for o in scene.objects)
        kernel_bytes = {"k_front": (idx_bytes + vtx_bytes + 16 * tri_in) / world,                           # inputs once + 16 B of bounds per input triangle
                        "k_back": (16 * tri_in + 216 * work["records_written"]) / world,                    # bounds read + record write
                        "k_tile": (216 * work["records_written"] + 1032 * work["tiles_visited"] + 4 * work["pairs_listed"]) / world,  # record read + tile read/write + list read
                        "k_front_clip": 0, "k_vertex": vtx_bytes / world, "k_bin_scan": 16 * (scene.width // 8) * (scene.height // 8) * len(scene.objects) / world,
                        "k_fill": (16 * stats["assembled_triangle_count"] + 4 * work["pairs_listed"]) / world}
        dom = max(("k_front", "k_back", "k_tile"), key=kernel_ms.get)  # the kernels that carry the geometry and record streams; the others move < 5 % of the bytes

        def kernel_roofline(k):
            n = max(kernel_launches[k], 1)
            ms_l, b_l = kernel_ms[k] / n, kernel_bytes[k] / n
            a = b_l / (ms_l * 1e-3) / 1e9 if ms_l > 0 else 0.0
            return {"achieved": a, "frac": a / peak, "algorithmic_bytes_per_launch": b_l, "ms_per_launch": ms_l, "launches_per_step": kernel_launches[k]}
        dom_r = kernel_roofline(dom)
        traffic, traffic_how = (None, "not measured (--no-traffic, or more than one GPU)")
        if not multi and not args.no_traffic:
            traffic, traffic_how = measure_traffic(args.config, dom)
        out = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "mtri_per_s": scene.input_triangles * fps / 1e6, "gpix_per_s": scene.width * scene.height * fps / 1e9,
            "config": workload_config(args.config, scene),
            "run": {"parallelism": f"sort-first x{world}, stripe {stripe} tile rows, composite {composite}" if multi else "single GPU",
                    "issue": "immediate calls" if args.immediate else "recorded command list (one CUDA-graph launch per frame)",
                    "assembled_triangles": stats["assembled_triangle_count"], "tri_tile_pairs": stats["total_triangle_count_in_bins"], "tile_draws": stats["active_bin_count"],
                    "l2": "no flush: per-frame working set (inputs + per-draw setup records) >> 126 MB L2" if args.config == 5 else "no flush", "kernels_source_sha256_16": kernels_source_hash()},
            "frame_fnv": {"color": hashes["color"], "depth": hashes["depth"]},
            "matches_golden": {"color": (hashes["color"] == hashes["golden_color"]) if hashes["golden_color"] else None,
                               "depth": (hashes["depth"] == hashes["golden_depth"]) if (hashes["depth"] and hashes["golden_depth"]) else None,
                               "note": "depth: the reference's own depth image (tests/golden/golden.json, bit-exact); colour: this repo's frame as committed after passing the <= 1/255 "
                                       "tolerance against the live reference (tests/golden/gpu_frames.json); depth is not exchanged between ranks, so it is hashed at N = 1 only"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_r["achieved"], "peak": peak, "unit": "GB/s",
                         "frac": dom_r["achieved"] / peak, "traffic": traffic, "traffic_source": traffic_how, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_r["algorithmic_bytes_per_launch"], "ms_per_launch": dom_r["ms_per_launch"], "launches_per_step": dom_r["launches_per_step"],
                         "units_per_step": {"records_written": work["records_written"], "tiles_visited": work["tiles_visited"], "pairs_listed": work["pairs_listed"]},
                         "note": "algorithmic bytes = SURVEY 8d per-unit figures x the units the kernel really processed (mlv_work_counters), averaged over its launches of a frame; "
                                 "the launch time is its CUDA-event bracket in immediate mode (no overlap with neighbours)"},
            "roofline_kernels": {k: kernel_roofline(k) for k in ("k_front", "k_back", "k_tile", "k_vertex", "k_fill")},
            "roofline_frame": {"algorithmic_bytes_per_frame": b_alg, "achieved": b_alg / (ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                               "frac": b_alg / (ms * 1e-3) / 1e9 / world / peak, "bytes_by_stage": b_stage,
                               "note": "B_alg of the REFERENCE algorithm for this frame (SURVEY 8d: every assembled triangle's record written and read, every touched tile-draw loaded and "
                                       "stored) over the measured frame time: delivered reference work per second, NOT a DRAM-throughput fraction -- it exceeds the peak when Hi-Z at binning time removes work"},
            "stage_ms_per_step": {k: round(v, 4) for k, v in stage_ms.items()},
            "e2e": e2e, "e2e_streaming": e2e_streaming,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if not multi and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample(scene)
        print(json.dumps(out), flush=True)
    if shared is not None:
        dev.finish()
        try:
            dev.unregister_host_memory(shared.address)
        except Exception:  # noqa: BLE001
            pass
        taken[0] = None
        shared.close(unlink=False)
    dev.close()
    if multi:
        dist.barrier()
        if shared is not None and rank == 0:
            try:
                os.unlink(shared.path)
            except OSError:
                pass
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--stripe", type=int, default=0, help="stripe height in tile rows for the sort-first split (0 = one contiguous band per rank)")
    ap.add_argument("--composite", default="p2p", choices=["p2p", "nccl"], help="multi-GPU exchange step: asynchronous peer-memory broadcast (default) or pack + ncclAllGather + unpack")
    ap.add_argument("--e2e-composite", default="host", choices=["host", "device"], help="N > 1, end-to-end loops: compose the frame in host memory (every rank delivers the rows it owns over its own PCIe link; default) "
                    "or on the device (peer-memory exchange, rank 0 reads the whole image back)")
    ap.add_argument("--immediate", action="store_true", help="issue every call every frame instead of replaying a recorded command list")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child process that measures roofline.traffic")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
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
