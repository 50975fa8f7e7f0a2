#!/usr/bin/env python3
"""Render one of the BASELINE configs on the GPU and write it as a binary PPM (replaces the reference's GDI blit,
main.c:286-357, for a headless box).  Usage: python tools/render_ppm.py <config 1-5|sup> <out.ppm> [width height]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from malevich_b200 import Device, scenes

which, out = sys.argv[1], sys.argv[2]
builder = scenes.suprematism if which == "sup" else scenes.CONFIGS[int(which)]
sc = builder(int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else builder()
with Device(sc.width, sc.height) as dev:
    scenes.render(dev, sc)
    col, _ = dev.present(want_depth=False)
    print(dev.stats())
rgb = np.stack([(col >> 16) & 0xFF, (col >> 8) & 0xFF, col & 0xFF], axis=-1).astype(np.uint8)  # 0x00RRGGBB (main.c:305-308)
with open(out, "wb") as f:
    f.write(b"P6\n%d %d\n255\n" % (sc.width, sc.height))
    f.write(rgb.tobytes())
print("wrote", out)
