#!/usr/bin/env python3
"""One RECORDED frame (command list = CUDA graph, front halves on their own streams) as it really overlaps on the device:
every geometry / binning / tile launch stamps %globaltimer (mlv_timeline_*). Prints one line per launch sorted by start
and writes a Chrome/Perfetto trace, one track per draw, scope names = the reference's Remotery scopes (main.c:663-1047).
Usage: python tools/timeline_frame.py [num_ranks] [rank] [config] [out.json]   (one rank of a sort-first split, no NCCL)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from malevich_b200 import Device, scenes
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 5
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join("gpurun_out", f"timeline_config{cfg}_w{world}_r{rank}.json")
REF_SCOPE = {"vertex_cache": "vertex_shader_stage (post-transform cache)", "geometry": "input_assembler+vertex_shader+primitive_assembly (front)",
             "clip": "primitive_assembly_stage (clipper)", "geometry_back": "Hi-Z + binner pass 1 + setup (back)", "bin_scan": "binner (scan)",
             "bin_fill": "binner (fill)", "tile": "rasterizer+pixel_shader_stage"}
sc = scenes.CONFIGS[cfg]()
stripe = max(1, -(-(sc.height // 8) // world))
with Device(sc.width, sc.height, num_ranks=world, rank=rank, stripe_height_tiles=stripe) as dev:
    scenes.upload(dev, sc)
    def frame():
        scenes.render(dev, sc)
        dev.composite_pack() if world > 1 else dev.resolve()
    frame(); dev.finish()
    dev.timeline_begin()
    cl = dev.record(frame)
    dev.timeline_end()
    for _ in range(3):
        cl.execute()
    dev.finish()
    dev.timeline_reset()
    cl.execute()
    dev.finish()
    ev = dev.timeline_read()
    cl.release()
ev = [e for e in ev if e[2] >= 0]
t_end = max(e[4] for e in ev)
print(f"world {world} rank {rank} config {cfg} ({sc.name}): {len(ev)} stamped launches, first stamp -> last end {t_end:.1f} us")
print("  draw stage          resident    start      end   (us)   busy")
for st, d, r, s, e in sorted(ev, key=lambda x: x[2]):
    print(f"  {d:4d} {st:14s} {r:8.1f} {s:8.1f} {e:8.1f}        {e - s:6.1f}")
trace = [{"name": "process_name", "ph": "M", "pid": 0, "args": {"name": f"malevich_b200 recorded frame, config {cfg} {sc.width}x{sc.height}, rank {rank} of {world}"}}]
for d in sorted({e[1] for e in ev}):
    trace.append({"name": "thread_name", "ph": "M", "pid": 0, "tid": d, "args": {"name": f"draw {d}"}})
for st, d, r, s, e in ev:
    trace.append({"name": REF_SCOPE.get(st, st), "cat": st, "ph": "X", "pid": 0, "tid": d, "ts": round(s, 3), "dur": round(max(e - s, 0.001), 3), "args": {"resident_us": round(r, 3)}})
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump({"traceEvents": trace, "displayTimeUnit": "ns"}, open(out, "w"), indent=0)
print("->", out)
