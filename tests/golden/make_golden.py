#!/usr/bin/env python3
"""Generates tests/golden/golden.json and tests/golden/small_frames.npz from the REFERENCE itself
(oracle/_ref, compiled from /root/reference by oracle/build_ref.sh), single-threaded = canonical order
(SURVEY.md 8a N1). Run in the build container:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import cases
from oracle.ref_oracle import RefOracle, fnv64_words


def run(name, build, frames_out):
    sc = build()
    orc = RefOracle(sc.width, sc.height, threads=1)
    orc.render(sc, staged_last=True)
    col, dep = orc.colors(), orc.depths()
    tris, attrs = orc.staged_triangles()
    ids, bins = orc.staged_bins()
    infos = orc.staged_tile_infos()
    entry = {
        "width": sc.width, "height": sc.height, "draws": len(sc.objects), "input_triangles": sc.input_triangles,
        "per_frame_cb_hex": sc.per_frame_cb.tobytes().hex(),
        "stats": orc.stats(),
        "color_fnv": fnv64_words(col), "depth_fnv": fnv64_words(dep),
        "tile_min_fnv": fnv64_words(orc.tile_min_depths() + np.float32(0.0)),
        "last_draw": {"triangles": int(len(tris)), "edges_fnv": fnv64_words(np.ascontiguousarray(tris["edges"])) if len(tris) else None,
                      "triangle_ids_fnv": fnv64_words(ids) if len(ids) else None,
                      "masks_fnv": fnv64_words(np.ascontiguousarray(infos["fragment_mask"])) if len(infos) else None,
                      "compacted_bins": int(len(bins)), "pairs": int(len(ids))},
    }
    if frames_out is not None:
        frames_out[name + "/colors"] = col
        frames_out[name + "/depths"] = dep
    return entry


def main():
    golden, frames = {}, {}
    for name, build in cases.SMALL.items():
        golden[name] = run(name, build, frames)
        print(name, golden[name]["stats"], flush=True)
    for name, build in {**cases.FULL, **cases.CONFIG5}.items():
        golden[name] = run(name, build, None)
        print(name, golden[name]["stats"], flush=True)
    # the screenshot pose with the constant buffer produced by the REFERENCE's own camera code (main.c:1422-1562)
    from malevich_b200 import scenes
    ref_cb = RefOracle(1200, 720).camera(*scenes.FTM_SCREENSHOT_POSE)
    golden["ftm_screenshot_refcam_1200x720"] = run("ftm_screenshot_refcam_1200x720", lambda: scenes.ftm(1200, 720, cb=ref_cb), None)
    # known answers that come from the reference repo itself, not from this build
    golden["_reference_known_answers"] = {
        "suprematism_1200x720_pixel_histogram": {"0x131510": 353587, "0x244a9b": 125591, "0xf5f5ed": 384822},
        "screenshot_png_overlay_ftm_1200x720": {"vertex_count": 143808, "input_triangle_count": 47936, "assembled_triangle_count": 23606,
                                                 "active_bin_count": 27481, "avg_triangles_per_bin": 5.49216},
    }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "small_frames.npz"), **frames)
    print("wrote golden.json and small_frames.npz", os.path.getsize(os.path.join(HERE, "small_frames.npz")), "bytes")


if __name__ == "__main__":
    main()
