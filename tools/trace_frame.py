#!/usr/bin/env python3
"""One frame as a timeline -- the replacement for the reference's Remotery scopes (rmt_BeginCPUSample in main.c:663, 699, 737,
916, 984, 1047, 1192, 1205; viewer external/Remotery/vis). Every kernel launch of the frame is bracketed with CUDA events
(mlv_profile_begin / mlv_profile_end / mlv_profile_read_events) and written as a Chrome / Perfetto trace (chrome://tracing,
ui.perfetto.dev) whose scope names are the reference's.

Usage: python tools/trace_frame.py [config 1-5] [out.json]   (run under gpurun; default gpurun_out/trace_config<k>.json)
Bracketing serialises the launches (no programmatic overlap), so the timeline shows each kernel's own duration, like ncu."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from malevich_b200 import Device, scenes

# device stage -> the reference scope(s) it replaces
REF_SCOPE = {
    "clear": "clear_render_target_view+clear_depth_stencil_view",
    "vertex_cache": "vertex_shader_stage (post-transform cache)",
    "geometry": "input_assembler_stage+vertex_shader_stage+primitive_assembly_stage",
    "clip": "primitive_assembly_stage (clipper)",
    "bin_count": "binner (count, large triangles)",
    "bin_scan": "binner (scan+compact)",
    "bin_fill": "binner (fill)",
    "tile": "rasterizer+pixel_shader_stage",
    "resolve": "present",
    "composite": "present (composite)",
}


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join("gpurun_out", f"trace_config{cfg}.json")
    sc = scenes.CONFIGS[cfg]()
    with Device(sc.width, sc.height) as dev:
        scenes.upload(dev, sc)
        for _ in range(3):
            scenes.render(dev, sc)
            dev.resolve()
        dev.finish()
        dev.profile_begin()
        scenes.render(dev, sc)
        dev.resolve()
        totals = dev.profile_end()
        events = dev.profile_events()
    trace = [{"name": "process_name", "ph": "M", "pid": 0, "args": {"name": f"malevich_b200 config {cfg}: {sc.name} {sc.width}x{sc.height}"}},
             {"name": "thread_name", "ph": "M", "pid": 0, "tid": 0, "args": {"name": "device stream (render)"}}]
    for stage, start, dur in events:
        trace.append({"name": REF_SCOPE.get(stage, stage), "cat": stage, "ph": "X", "pid": 0, "tid": 0, "ts": round(start * 1e3, 3), "dur": round(dur * 1e3, 3)})
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        json.dump({"traceEvents": trace, "displayTimeUnit": "ns",
                   "otherData": {"stage_ms": {k: round(v[0], 4) for k, v in totals.items()}, "launches": len(events)}}, f, indent=0)
    span = max((s + d for _, s, d in events), default=0.0)
    print(f"{len(events)} launches, {span:.3f} ms serialised -> {out}")
    for k, (ms, n) in totals.items():
        if n:
            print(f"  {k:13s} {n:3d} launches {ms:8.4f} ms  [{REF_SCOPE.get(k, k)}]")


if __name__ == "__main__":
    main()
