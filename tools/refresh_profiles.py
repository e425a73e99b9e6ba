"""Copies the artefacts of tools/r2_measure.sh from gpurun_out/ into profiles/ (bench lines, per-layer and training tables, launch
lists) and regenerates the derived files: profiles/r02_traffic.json from the DRAM-traffic capture, the tables inside
profiles/r02_ncu_conv_f16f8.md and r02_ncu_warp.md from the .ncu-rep captures (their prose sections are kept)."""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))

for f in ("r02_bench_n1.json", "r02_bench_n1_f16x3.json", "r02_bench_reference_arm.json", "r02_layers_4x544x992_f16f8.txt",
          "r02_layers_4x544x992_f16x3.txt", "r02_layers_8x192x192_f16f8.txt", "r02_layers_8x192x192_f16x3.txt",
          "r02_train_cfg3_B16_192.txt", "r02_pwc_launches.txt"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))

# ---- DRAM traffic of the 138 conv launches of one forward
src = os.path.join(G, "r02_traffic_f16f8.csv")
if os.path.exists(src):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    mi, vi, ui, ii = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1}
    tot, ids = {}, set()
    for r in rows[1:]:
        tot[r[mi]] = tot.get(r[mi], 0.0) + float(r[vi].replace(",", "")) * mult[r[ui]]
        ids.add(r[ii])
    old = json.load(open(os.path.join(P, "r02_traffic.json")))
    old["f16f8"] = {"launches": len(ids), "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
                    "gpu_time_ms_under_ncu": tot["gpu__time_duration.sum"]}
    json.dump(old, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)

# ---- launch list of the bench command
src = os.path.join(G, "r02_launches_f16f8.txt")
if os.path.exists(src):
    dst = os.path.join(P, "r02_launches_4x544x992.md")
    head = open(dst).read().split("```")[0]
    open(dst, "w").write(head + "```\n" + open(src).read().rstrip() + "\n```\n")


def summary(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    return out.split("\n", 3)[3].rstrip()          # drop the "### <file>" heading


# ---- tables of the conv captures: one "### " section per capture, in this order
reps = [os.path.join(G, f"prof_f16f8_{n}.ncu-rep") for n in ("conv64", "pool64", "conv128", "head")]
dst = os.path.join(P, "r02_ncu_conv_f16f8.md")
if all(os.path.exists(r) for r in reps):
    parts = re.split(r"(?m)^(### .*|## Reading.*)$", open(dst).read())
    out, k = parts[0], 0
    for i in range(1, len(parts), 2):
        if parts[i].startswith("### "):
            out += parts[i] + "\n\n" + summary(reps[k]) + "\n\n"
            k += 1
        else:
            out += parts[i] + parts[i + 1]
    open(dst, "w").write(out)

rep = os.path.join(G, "prof_pwc_level2.ncu-rep")
dst = os.path.join(P, "r02_ncu_pwc.md")
if os.path.exists(rep):
    old = open(dst).read()
    i0, i1 = old.index("| metric | unit |"), old.index("\n## Reading")
    open(dst, "w").write(old[:i0] + summary(rep) + "\n" + old[i1:])

rep = os.path.join(G, "prof_warp.ncu-rep")
dst = os.path.join(P, "r02_ncu_warp.md")
if os.path.exists(rep):
    new = {m.group(1): m.group(0) for m in re.finditer(r"\| `([^`]+)` \|[^\n]*", summary(rep))}
    open(dst, "w").write(re.sub(r"\| `([^`]+)` \|[^\n]*", lambda m: new.get(m.group(1), m.group(0)), open(dst).read()))
print("profiles refreshed")
