#!/usr/bin/env python3
"""Frame time of the config-5 workload as a function of the number of layers (draws): the slope is the cost of one
hidden draw, the intercept the cost of the visible one. Usage: python tools/marginal_draw.py [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from malevich_b200 import Device, scenes
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
full = scenes.synthetic()
with Device(full.width, full.height) as dev:
    scenes.upload(dev, full)
    for layers in (1, 2, 4, 8):
        sc = scenes.Scene(full.name, full.width, full.height, full.objects[:layers], full.per_frame_cb, full.camera_pose)
        for _ in range(3):
            scenes.render(dev, sc); dev.resolve()
        dev.finish()
        t0 = time.perf_counter()
        for _ in range(frames):
            scenes.render(dev, sc); dev.resolve()
        dev.finish()
        print(f"layers {layers}: {1e3 * (time.perf_counter() - t0) / frames:.3f} ms/frame")
