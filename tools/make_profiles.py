#!/usr/bin/env python
"""Turn the artefacts of one profiling gpurun call into the tracked files under profiles/.

The GPU-side commands (run from the repo root on the box, see DESIGN.md "Measurement"):

    python bench.py > gpurun_out/bench_TAG.json
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_TAG.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --bf-size 4096 --no-rig8 --no-configs --warm-seconds 0
    ncu --set full --clock-control none --import-source on --launch-skip 75 -c 25 -f -o gpurun_out/prof_TAG \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --bf-size 4096 --no-rig8 --no-configs --warm-seconds 0

usage: tools/make_profiles.py TAG   (reads gpurun_out/*_TAG.*, writes profiles/TAG_*)"""
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")


def short(name):
    name = name.replace("orbk::", "").replace("<unnamed>::", "").replace("void ", "")
    m = re.match(r"([\w:]+(?:<[^>]*>)?)", name)
    return m.group(1).replace("(int)", "") if m else name


# ---- bench line ---------------------------------------------------------------------------------
bench = json.loads(open(os.path.join(go, f"bench_{tag}.json")).read().strip().splitlines()[-1])
json.dump(bench, open(os.path.join(pr, f"{tag}_bench_n1.json"), "w"), indent=1)

# ---- launch list --------------------------------------------------------------------------------
src = os.path.join(go, f"launches_{tag}.csv")
lines = [l for l in open(src).read().splitlines() if l.startswith('"')]
rows = list(csv.DictReader(lines))
open(os.path.join(pr, f"{tag}_launches.csv"), "w").write("\n".join(lines) + "\n")
# one timed step of the device-resident leg = the launches between the warm-up and the e2e leg;
# shares are taken over all extractor/matcher launches of the run (same mix every step)
agg, other = {}, {}
STEP_KERNELS = ("k_pyr_resize", "k_fast_cells", "k_octree", "k_blur", "k_orient_describe", "k_build_grid", "k_init_")
for r in rows:
    k = short(r["Kernel Name"])
    if "at::" in k or "elementwise" in k or "reduce" in k.lower():
        continue
    a = (agg if k.startswith(STEP_KERNELS) else other).setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(os.path.join(pr, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag} — ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --bf-size 4096 --no-rig8 --no-configs --warm-seconds 0`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and "
            "serialised: compare SHARES with bench.py's `stages`, not absolutes.  Extractor and "
            "SearchForInitialization kernels of all steps (device-resident leg and streaming leg).  In bench.py the "
            "matcher runs on a side stream underneath camera 2's pyramid / FAST, so its own share and the pyramid's "
            "read a little higher there than in this serialised list.\n\n"
            "| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
    for k, (n, us) in agg.items():
        f.write(f"| {k} | {n} | {us:.1f} | {us / n:.1f} | {100 * us / tot:.1f} % |\n")
    f.write("\nKernels of the other legs of the same run (brute-force leg, widened rows of SURVEY.md 8(f)), not part of a step:\n\n"
            "| kernel | launches | total us | avg us |\n|---|---|---|---|\n")
    for k, (n, us) in other.items():
        f.write(f"| {k} | {n} | {us:.1f} | {us / n:.1f} |\n")
    st = bench.get("stages", {})
    if st:
        ssum = sum(v["ms"] for v in st.values())
        f.write("\nbench.py stage shares of the same build (CUDA events, warm): " +
                ", ".join(f"{k} {100 * v['ms'] / ssum:.1f} %" for k, v in st.items()) + "\n")

# ---- full capture -------------------------------------------------------------------------------
rep = os.path.join(go, f"prof_{tag}.ncu-rep")
subprocess.run([sys.executable, os.path.join(root, "tools", "profile_summary.py"), rep,
                os.path.join(pr, f"{tag}_ncu_all_kernels.md"),
                f"{tag}: every extractor/matcher kernel of one bench step (256 rig-frames per handle)"],
               check=True, stdout=subprocess.DEVNULL)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
ki, ri, wi, gi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("launch__grid_size")
ii = h.index("smsp__inst_executed.sum")
units = dict(zip(h, rr[1]))
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
stage_of = {"k_pyr_resize": "pyramid", "k_fast_cells": "fast", "k_octree": "octree", "k_blur": "blur",
            "k_orient_describe": "orient_describe"}
F = bench["config"]["rig_frames_per_gpu"]
per = {}
instr = {}
seen_handle = {}
for r in rr[2:]:
    k = short(r[ki]).split("<")[0]
    if k not in stage_of:
        continue
    st = stage_of[k]
    # first handle only (nFeatures 1000): the first occurrence of each kernel (7 for the pyramid)
    n_max = 7 if st == "pyramid" else 1
    if seen_handle.get(st, 0) >= n_max:
        continue
    seen_handle[st] = seen_handle.get(st, 0) + 1
    b = float(r[ri]) * scale.get(units["dram__bytes_read.sum"], 1.0) + float(r[wi]) * scale.get(units["dram__bytes_write.sum"], 1.0)
    per[st] = per.get(st, 0.0) + b / F
    instr[st] = instr.get(st, 0.0) + float(r[ii]) / F
json.dump({"source": f"profiles/{tag}_ncu_all_kernels.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of the "
                     f"nFeatures-1000 handle's launches, one launch = {F} camera-frames)",
           "dram_bytes_per_camera_frame": per, "warp_instr_per_camera_frame": instr}, open(os.path.join(pr, f"{tag}_traffic.json"), "w"), indent=1)
print(open(os.path.join(pr, f"{tag}_launches.md")).read())
print(json.dumps(per, indent=1))
