#!/usr/bin/env python3
"""Turns the raw captures a gpurun call left under gpurun_out/ into the tracked summaries under profiles/.
Usage: python tools/make_profiles.py <launches.csv> <full.ncu-rep> <bench_n1.json> [round-tag]"""
import csv, json, os, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, bench_json = sys.argv[1:4]
tag = sys.argv[4] if len(sys.argv) > 4 else "r02"
out = lambda name: os.path.join(ROOT, "profiles", f"{tag}_{name}")

# ---- launch list: verbatim copy (minus ncu's banner lines) + per-kernel summary + per-launch DRAM traffic
lines = open(launch_csv).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
open(out("launches_config5_n1.csv"), "w").write("\n".join(lines[start:]) + "\n")
rows = list(csv.reader(lines[start:]))
h = rows[0]
ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = OrderedDict()
for r in rows[1:]:
    if len(r) > vi:
        per.setdefault((r[idi], r[ki].split("(")[0].replace("void ", "")), {})[r[mi]] = float(r[vi].replace(",", "") or 0)
agg = OrderedDict()
for (_, k), m in per.items():
    base = k.split("<")[0]
    a = agg.setdefault(base, {"name": k, "n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "each": []})
    a["n"] += 1; a["ns"] += m.get("gpu__time_duration.sum", 0); a["rd"] += m.get("dram__bytes_read.sum", 0); a["wr"] += m.get("dram__bytes_write.sum", 0)
    a["each"].append(round(m.get("gpu__time_duration.sum", 0) / 1e3, 1))
tot = sum(a["ns"] for a in agg.values())
with open(out("launch_summary_config5_n1.txt"), "w") as f:
    f.write(f"one frame of config 5 under ncu (cold-cache, serialised launches): {tot / 1e3:.1f} us in {sum(a['n'] for a in agg.values())} launches\n")
    for b, a in sorted(agg.items(), key=lambda x: -x[1]["ns"]):
        f.write("%-26s n=%2d  sum %7.1f us (%4.1f%%)  dram rd %7.1f MB wr %6.1f MB  per-launch us %s\n" % (a["name"], a["n"], a["ns"] / 1e3, 100 * a["ns"] / tot, a["rd"] / 1e6, a["wr"] / 1e6, a["each"]))
json.dump({"source": f"profiles/{tag}_launches_config5_n1.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one frame of config 5 on 1 B200)",
           "kernels": {b: {"launches_per_frame": a["n"], "dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"], "ncu_ns_per_launch": a["ns"] / a["n"]} for b, a in agg.items()}},
          open(out("traffic_config5_n1.json"), "w"), indent=1)

# ---- --set full capture: per-kernel summary, selected raw metrics, per-source-line hot spots
tools = os.path.join(ROOT, "tools")
open(out("ncu_full_summary.txt"), "w").write(subprocess.run([sys.executable, os.path.join(tools, "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
want = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
cols = [raw[0].index(w) for w in want if w in raw[0]]
with open(out("ncu_full_raw_selected.csv"), "w", newline="") as f:
    csv.writer(f).writerows([[r[c] for c in cols] for r in raw])
with open(out("ncu_source_hotspots.txt"), "w") as f:
    for title, kern, extra in (("k_tile, visible draw", "k_tile", ["--launch-count", "1"]), ("k_back, visible draw (launch 1)", "k_back", ["--launch-count", "1"]),
                               ("k_back, hidden draw (launch 2)", "k_back", ["--launch-skip", "1", "--launch-count", "1"]),
                               ("k_front (the same work for every draw)", "k_front<", ["--launch-count", "1"])):
        f.write(f"==== {title}: CUDA source lines ranked by warp instructions executed (ncu --page source)\n")
        f.write(subprocess.run([sys.executable, os.path.join(tools, "ncu_lines.py"), rep, kern, "30"] + extra, capture_output=True, text=True).stdout + "\n")

# ---- the bench line(s): config 5 = the headline; every other bench_*.json / timeline / table found next to it travels too
json.dump(json.loads(open(bench_json).read().strip().splitlines()[-1]), open(out("bench_config5_n1.json"), "w"), indent=1)
src_dir = os.path.dirname(os.path.abspath(bench_json))
import shutil
for name, dst in [(f"bench_c{c}.json", f"bench_config{c}_n1.json") for c in (1, 2, 3, 4)] + [("bench_ref.json", "bench_reference_arm_config5.json")] + \
                 [(f"bench_n{n}.json", f"bench_config5_n{n}.json") for n in (2, 4, 8)] + [(f"bench_n{n}_nccl.json", f"bench_config5_n{n}_nccl.json") for n in (2, 4, 8)] + \
                 [(f"group_n{n}_{m}.json", f"group_config5_n{n}_{m}.json") for n in (2, 4, 8) for m in ("peer", "nccl")] + [(f"group_n{n}.json", f"group_config5_n{n}_host.json") for n in (2, 4, 8)] + [("group_n8_ftm.json", "group_config2_n8_peer.json"), ("group_n2_ftm.json", "group_config2_n2_peer.json")]:
    path = os.path.join(src_dir, name)
    if os.path.exists(path) and open(path).read().strip():
        json.dump(json.loads(open(path).read().strip().splitlines()[-1]), open(out(dst), "w"), indent=1)
for name in ("timeline_config5_n1.json", "timeline_config5_n1.txt", "timeline_config5_rank3of8.json", "timeline_config5_rank3of8.txt", "timeline_config2_n1.json", "timeline_config2_n1.txt",
             "launch_table_config5_n1.txt", "trace_config5.json", "sanitizer_memcheck.log", "sanitizer_racecheck.log", "sanitizer_synccheck.log"):
    path = os.path.join(src_dir, name)
    if os.path.exists(path):
        shutil.copy(path, out(name))
w8 = os.path.join(src_dir, "launches_w8_r3.csv")
if os.path.exists(w8):
    l8 = open(w8).read().splitlines()
    s8 = next(i for i, l in enumerate(l8) if l.startswith('"ID"'))
    open(out("launches_config5_rank3of8.csv"), "w").write("\n".join(l8[s8:]) + "\n")
print("wrote", sorted(x for x in os.listdir(os.path.join(ROOT, "profiles")) if x.startswith(tag)))
