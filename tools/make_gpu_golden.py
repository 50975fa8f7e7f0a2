#!/usr/bin/env python3
"""Writes gpurun_out/gpu_frames.json (to be committed as tests/golden/gpu_frames.json): the colour and depth hashes of the frames
the PRODUCTION path renders for the five BASELINE configs -- each written only after the frame passed the parity bars against
the live reference (depth bit-exact, colour within 1/255 on >= 99.9 % of the pixels, none off by more than 2/255, Stats equal).
bench.py prints the same hashes for the frame it times (`frame_fnv` / `matches_golden`), at every GPU count.
Run on a GPU box: python tools/make_gpu_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
from malevich_b200 import Device, scenes
from malevich_b200._lib import fnv64_words
from oracle.ref_oracle import RefOracle

KEYS = {1: "config1_toon_1280x720", 2: "config2_ftm_1920x1080", 3: "config3_emily_1920x1080", 4: "config4_locomotive_3840x2160", 5: "config5_synthetic_3840x2160"}
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
out = {}
for cfg, key in KEYS.items():
    sc = scenes.CONFIGS[cfg]()
    orc = RefOracle(sc.width, sc.height, threads=os.cpu_count() or 1)
    orc.render(sc)
    with Device(sc.width, sc.height) as dev:
        dev.reset_stats()
        scenes.render(dev, sc)
        col, dep = dev.present()
        st = dev.stats()
    r = parity.assert_frames_match(col, dep, orc.colors(), orc.depths(), key)
    assert st == orc.stats() == golden[key]["stats"], (st, orc.stats())
    assert fnv64_words(dep) == golden[key]["depth_fnv"]
    out[key] = {"color_fnv": fnv64_words(col), "depth_fnv": fnv64_words(dep), "color_exact_fraction_vs_reference": r["color_exact_fraction"], "color_max_diff_vs_reference": r["color_max_diff"]}
    print(key, out[key], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_frames.json"), "w"), indent=1, sort_keys=True)
