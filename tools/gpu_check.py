#!/usr/bin/env python3
"""Ad-hoc GPU-vs-oracle report (run under gpurun). Writes gpurun_out/gpu_check.json.
Usage: python tools/gpu_check.py [scene ...]   scenes: sup320 sup1200 toon ftm emily loco synth_small synth
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from malevich_b200 import Device, scenes
from oracle.ref_oracle import RefOracle
import parity


def build(name):
    if name == "sup320":
        return scenes.suprematism(320, 200)
    if name == "sup1200":
        return scenes.suprematism(1200, 720)
    if name == "toon":
        return scenes.toon()
    if name == "toon320":
        return scenes.toon(320, 200)
    if name == "ftm":
        return scenes.ftm()
    if name == "ftm1200":
        return scenes.ftm(1200, 720)
    if name == "ftm4k":
        return scenes.ftm(3840, 2160)
    if name == "emily":
        return scenes.emily(n_lat=128, n_lon=256)
    if name == "loco":
        return scenes.locomotive(n_u=1024, n_v=64)
    if name == "synth_small":
        return scenes.synthetic(1280, 720, layers=3, nx=400, ny=200)
    if name == "synth":
        return scenes.synthetic()
    raise SystemExit("unknown scene " + name)


def main():
    names = sys.argv[1:] or ["sup320", "sup1200", "toon320", "toon", "ftm1200", "ftm", "emily", "loco", "synth_small"]
    report = {}
    ok_all = True
    for name in names:
        sc = build(name)
        staged = name != "synth"
        t0 = time.time()
        orc = RefOracle(sc.width, sc.height, threads=1)
        entry = {"width": sc.width, "height": sc.height, "input_triangles": sc.input_triangles}
        with Device(sc.width, sc.height, debug_capture=staged) as dev:
            if staged:
                res = parity.render_both_staged(dev, orc, sc)
                entry["draws"] = [{"name": n, "ok": parity.staged_ok(r), **{k: v for k, v in r.items()}} for n, r in res]
                stage_ok = all(parity.staged_ok(r) for _, r in res)
            else:
                scenes.render(dev, sc)
                orc.render(sc)
                stage_ok = True
            gc, gd = dev.present()
            entry["stats_gpu"] = dev.stats()
            entry["stats_ref"] = orc.stats()
            fr = parity.compare_frames(gc, gd, orc.colors(), orc.depths())
            entry["frame"] = fr
            frame_ok = fr["depth_bit_exact"] and fr["color_within1_fraction"] >= parity.COLOR_TOL_FRACTION and fr["color_max_diff"] <= parity.COLOR_TOL_MAX
            entry["ok"] = bool(stage_ok and frame_ok and entry["stats_gpu"] == entry["stats_ref"])
        entry["seconds"] = round(time.time() - t0, 2)
        ok_all &= entry["ok"]
        report[name] = entry
        print(name, "OK" if entry["ok"] else "MISMATCH", json.dumps(entry, default=str)[:3000], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gpu_check.json", "w") as f:
        json.dump(report, f, indent=1, default=str)
    print("ALL OK" if ok_all else "SOME MISMATCH")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
