"""Named parity cases shared by tests/golden/make_golden.py and the tests.

Small cases (320x200) have their full oracle colour+depth images committed under tests/golden/; the
full-size BASELINE.json configs are pinned by hashes + Stats and compared live against oracle/_ref.
"""
from malevich_b200 import scenes

SMALL = {
    "sup_320x200": lambda: scenes.suprematism(320, 200),
    "toon_320x200": lambda: scenes.toon(320, 200),
    "ftm_320x200": lambda: scenes.ftm(320, 200),
    "emily_320x200": lambda: scenes.emily(320, 200, n_lat=64, n_lon=128),
    "loco_320x200": lambda: scenes.locomotive(320, 200, n_u=512, n_v=32),
    "synth_320x200": lambda: scenes.synthetic(320, 200, layers=3, nx=120, ny=60),
}

FULL = {
    "sup_1200x720": lambda: scenes.suprematism(1200, 720),
    "ftm_screenshot_1200x720": lambda: scenes.ftm(1200, 720),
    "config1_toon_1280x720": lambda: scenes.toon(),
    "config2_ftm_1920x1080": lambda: scenes.ftm(),
    "config3_emily_1920x1080": lambda: scenes.emily(),
    "config4_locomotive_3840x2160": lambda: scenes.locomotive(),
}

# config 5 (10 M triangles, 3840x2160) takes the 1-thread oracle ~20 s; it has its own test
CONFIG5 = {"config5_synthetic_3840x2160": lambda: scenes.synthetic()}
