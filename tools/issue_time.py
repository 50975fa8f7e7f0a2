#!/usr/bin/env python3
"""Host issue time vs device time per frame, as ONE rank of a sort-first split (no NCCL), in immediate mode and as a
recorded command list (one CUDA-graph launch per frame).
Usage: python tools/issue_time.py <num_ranks> <rank> [frames] [config] [stripe]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from malevich_b200 import Device, scenes
world, rank = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sc = scenes.CONFIGS[int(sys.argv[4]) if len(sys.argv) > 4 else 5]()
stripe = int(sys.argv[5]) if len(sys.argv) > 5 and int(sys.argv[5]) > 0 else max(1, -(-(sc.height // 8) // world))
with Device(sc.width, sc.height, num_ranks=world, rank=rank, stripe_height_tiles=stripe) as dev:
    scenes.upload(dev, sc)
    def frame():
        scenes.render(dev, sc)
        dev.composite_pack() if world > 1 else dev.resolve()
    def timed(step, label):
        for _ in range(3):
            step()
        dev.finish()
        t0 = time.perf_counter()
        for _ in range(frames):
            step()
        t1 = time.perf_counter()
        dev.finish()
        t2 = time.perf_counter()
        print(f"world {world} rank {rank} stripe {stripe} {label}: issue {1e3*(t1-t0)/frames:.3f} ms/frame, issue+drain {1e3*(t2-t0)/frames:.3f} ms/frame", flush=True)
    timed(frame, "immediate")
    cl = dev.record(frame)
    timed(cl.execute, "command list")
    cl.release()
