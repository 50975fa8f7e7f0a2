#!/usr/bin/env python3
"""Device time per frame of EVERY rank of a sort-first split, one after the other on one GPU (recorded command list,
no exchange): shows the load balance of the split. Usage: python tools/rank_times.py <num_ranks> [frames] [config] [stripe]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from malevich_b200 import Device, scenes
world = int(sys.argv[1])
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 30
config = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sc = scenes.CONFIGS[config]()
stripe = int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) > 0 else max(1, -(-(sc.height // 8) // world))
out = []
for rank in range(world):
    with Device(sc.width, sc.height, num_ranks=world, rank=rank, stripe_height_tiles=stripe) as dev:
        scenes.upload(dev, sc)
        def frame():
            scenes.render(dev, sc)
            dev.composite_pack() if world > 1 else dev.resolve()
        frame(); dev.finish()
        cl = dev.record(frame)
        for _ in range(5):
            cl.execute()
        dev.finish()
        t0 = time.perf_counter()
        for _ in range(frames):
            cl.execute()
        dev.finish()
        ms = 1e3 * (time.perf_counter() - t0) / frames
        st = dev.stats(); wk = dev.work_counters()
        cl.release()
    out.append({"rank": rank, "ms_per_frame": round(ms, 4), "assembled": st["assembled_triangle_count"], "pairs": st["total_triangle_count_in_bins"], "records_written": wk["records_written"]})
    print(out[-1], flush=True)
print(json.dumps({"tool": "rank_times", "num_ranks": world, "config": config, "stripe_height_tiles": stripe, "ranks": out, "max_ms": max(o["ms_per_frame"] for o in out), "min_ms": min(o["ms_per_frame"] for o in out)}))
