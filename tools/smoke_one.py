import sys, os
sys.path.insert(0, '/root/repo')
from malevich_b200 import Device, scenes
sc = scenes.toon(320, 200)
with Device(sc.width, sc.height) as dev:
    print("created", flush=True)
    scenes.upload(dev, sc); dev.finish(); print("uploaded", flush=True)
    scenes.render(dev, sc); print("issued", flush=True)
    dev.finish(); print("finished", flush=True)
    print(dev.stats(), flush=True)
