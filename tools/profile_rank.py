#!/usr/bin/env python3
"""Render config-N frames as ONE rank of a sort-first split, without NCCL, so the per-rank kernels can be put under ncu
(ncu must never wrap a multi-rank command). Usage: python tools/profile_rank.py <num_ranks> <rank> [frames] [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from malevich_b200 import Device, scenes
world, rank = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc = scenes.CONFIGS[int(sys.argv[4]) if len(sys.argv) > 4 else 5]()
stripe = max(1, -(-(sc.height // 8) // world))
with Device(sc.width, sc.height, num_ranks=world, rank=rank, stripe_height_tiles=stripe) as dev:
    scenes.upload(dev, sc)
    for _ in range(frames):
        scenes.render(dev, sc)
        if world > 1:
            dev.composite_pack()
        else:
            dev.resolve()
    dev.finish()
    print(dev.stats())
