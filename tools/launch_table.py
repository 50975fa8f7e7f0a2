#!/usr/bin/env python3
"""Per-launch durations (us) of one frame, immediate mode with every launch bracketed by CUDA events, as ONE rank of a
sort-first split. One line per draw. Usage: python tools/launch_table.py [num_ranks] [rank] [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from malevich_b200 import Device, scenes
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = scenes.CONFIGS[int(sys.argv[3]) if len(sys.argv) > 3 else 5]()
stripe = max(1, -(-(sc.height // 8) // world))
with Device(sc.width, sc.height, num_ranks=world, rank=rank, stripe_height_tiles=stripe) as dev:
    scenes.upload(dev, sc)
    def frame():
        scenes.render(dev, sc)
        dev.composite_pack() if world > 1 else dev.resolve()
    for _ in range(3):
        frame()
    dev.finish()
    dev.profile_begin()
    frame()
    dev.profile_end()
    ev = dev.profile_events()
line, total = [], {}
print(f"world {world} rank {rank} config {sc.name}: {len(ev)} launches")
for stage, start, dur in ev:
    total[stage] = total.get(stage, 0.0) + dur
    line.append(f"{stage}={dur*1e3:.1f}")
    if stage in ("tile", "resolve", "composite", "clear"):
        print("  " + " ".join(line)); line = []
if line: print("  " + " ".join(line))
print("  totals(us): " + " ".join(f"{k}={v*1e3:.1f}" for k, v in total.items()))
