#!/usr/bin/env python3
"""A device GROUP on real GPUs: ONE process, one host thread, N CUDA devices behind the single-device calls
(mlv_device_desc.num_gpus, include/malevich_b200.h "DEVICE GROUPS"). Renders a BASELINE config as a recorded command list
per GPU, composes the frame over peer memory (default) or ncclAllGather (--nccl), checks colour AND depth hashes of the
composed frame against the committed golden frames and prints one JSON line.
Usage: python tools/group_bench.py <num_gpus> [--config K] [--nccl] [--stripe S] [--frames F] [--same-gpu]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from malevich_b200 import Device, scenes, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("num_gpus", type=int)
ap.add_argument("--config", type=int, default=5)
ap.add_argument("--nccl", action="store_true", help="compose on the device with pack + ncclAllGather + unpack, read rank 0's image back")
ap.add_argument("--peer", action="store_true", help="compose on the device with the asynchronous peer-memory exchange, read rank 0's image back (default: compose in host memory, every GPU delivers its band)")
ap.add_argument("--stripe", type=int, default=0)
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--same-gpu", action="store_true")
a = ap.parse_args()
KEYS = {1: "config1_toon_1280x720", 2: "config2_ftm_1920x1080", 3: "config3_emily_1920x1080", 4: "config4_locomotive_3840x2160", 5: "config5_synthetic_3840x2160"}
golden = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gpu_frames.json")))[KEYS[a.config]]
sc = scenes.CONFIGS[a.config]()
with Device(sc.width, sc.height, cuda_device=0, num_gpus=a.num_gpus, group_same_gpu=a.same_gpu, group_nccl=a.nccl, group_peer_exchange=a.peer, stripe_height_tiles=a.stripe) as dev:
    scenes.upload(dev, sc)
    scenes.render(dev, sc)
    col, dep = dev.present()
    stats = dev.stats()
    cls = [dev.record(lambda: scenes.render(dev, sc)) for _ in range(2)]  # two recordings: they alternate between the two tiled framebuffers of every GPU
    # pinned host memory (through torch, when present) and a read-back one frame behind: the host collects frame f-1 while the
    # GPUs render frame f -- what bench.py's e2e does with one process per GPU
    try:
        import torch
        bufs = [torch.empty((sc.height, sc.width), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32) for _ in range(2)]
        pinned = True
    except Exception:
        bufs = [np.empty((sc.height, sc.width), dtype=np.uint32) for _ in range(2)]
        pinned = False
    colors = bufs[0]
    def frame(f):
        cls[f & 1].execute()
        dev.present_wait()            # frame f-1 has arrived in bufs[(f-1) & 1]
        dev.present_async(bufs[f & 1])
    for f in range(6):
        frame(f)
    dev.present_wait()
    dev.finish()
    t0 = time.perf_counter()
    for f in range(a.frames):
        frame(f)
    dev.present_wait()
    dev.finish()
    ms = 1e3 * (time.perf_counter() - t0) / a.frames
    colors = bufs[(a.frames - 1) & 1]
    for cl in cls:
        cl.release()
print(json.dumps({"tool": "group_bench", "num_gpus": a.num_gpus, "config": KEYS[a.config], "exchange": "ncclAllGather" if a.nccl else ("peer memory" if a.peer else "none: every GPU copies its band to the host frame over its own PCIe link"),
                  "stripe_height_tiles": a.stripe or "one band per GPU", "ms_per_frame_incl_readback": round(ms, 4), "frames_per_s": round(1e3 / ms, 1),
                  "color_fnv": L.fnv64_words(col), "depth_fnv": L.fnv64_words(dep), "matches_golden": {"color": L.fnv64_words(col) == golden["color_fnv"], "depth": L.fnv64_words(dep) == golden["depth_fnv"]},
                  "replayed_frame_matches": L.fnv64_words(colors) == golden["color_fnv"], "stats": stats,
                  "pinned_host_memory": pinned, "note": "one process, one host thread; each frame = one graph launch per GPU + exchange + read-back of the composed 4-byte-per-pixel image to host memory, one frame behind"}))
