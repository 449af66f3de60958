"""Turns the files profiles/measure_round.sh left in gpurun_out/ into the tracked summaries under profiles/:
bench JSON lines, the ncu launch-list summary (per-kernel share of the listed time) and the per-kernel raw metrics
of the --set full captures (+ traffic.json, which bench.py reads for roofline.traffic).
    python profiles/summarize.py r1"""
import csv
import json
import os
import shutil
import sys
from collections import OrderedDict

R = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(os.path.join(P, "bench"), exist_ok=True)

for name in os.listdir(O):
    if name.startswith(f"{R}_bench_") and name.endswith(".json") and os.path.getsize(os.path.join(O, name)) > 0:
        shutil.copy(os.path.join(O, name), os.path.join(P, "bench", name))

# ---- launch list ---------------------------------------------------------------------------------------------------
ll = os.path.join(O, f"{R}_launches_cfg2.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi or "gpu__time_duration.sum" not in r:
            continue
        k = r[ki]
        ns = float(r[vi].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        ns = ns * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    mine = {k: v for k, v in agg.items() if "dsn::" in k or "nms_t" in k}       # ncu prints the inline-namespace NMS kernels without the outer dsn::
    total = sum(v[1] for v in mine.values()) or 1.0
    shutil.copy(ll, os.path.join(P, f"{R}_launches_cfg2.csv"))
    with open(os.path.join(P, f"{R}_launches_cfg2_summary.csv"), "w") as f:
        f.write(f"# ncu launch list of `bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary` (cfg 2, bs 64, CUDA-graph replay, 3 batches in flight), {R}\n")
        f.write("# per-launch gpu__time_duration.sum (cold-cache, serialised): compare SHARES, not absolutes\n")
        f.write("kernel,launches,mean_ns,share_of_listed_time\n")
        for k, (n, t) in sorted(mine.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k[:90]}\",{n},{t / n:.0f},{t / total:.3f}\n")
        other = sum(v[1] for k, v in agg.items() if k not in mine)
        f.write(f"# launches of other kernels (torch fills / copies in setup, reductions): {sum(v[0] for k, v in agg.items() if k not in mine)}, "
                f"{other / 1e3:.0f} us in total\n")

# ---- --set full captures ------------------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]
summary, traffic = OrderedDict(), OrderedDict()
PROF_BS = {2: 64, 3: 512, 4: 32, 5: 128}       # profiles/measure_round.sh: the bench's own batch sizes
for c in (2, 3, 4, 5):
    raw = os.path.join(O, f"{R}_ncu_full_cfg{c}_raw.csv")
    if not os.path.exists(raw) or os.path.getsize(raw) == 0:
        continue
    rows = list(csv.reader(open(raw, errors="replace")))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    shutil.copy(raw, os.path.join(P, f"{R}_ncu_full_cfg{c}_raw.csv"))
    seen = OrderedDict()
    for r in rows[2:]:
        if len(r) > ki:
            seen[r[ki]] = r                      # last capture of each kernel (the first is the cold one)
    # a later capture of single kernels of the same config (e.g. after a kernel change) overrides those rows
    for extra in sorted(f for f in os.listdir(O) if f.startswith(f"{R}_ncu_full_cfg{c}_") and f.endswith("_raw.csv") and f != os.path.basename(raw)):
        erows = list(csv.reader(open(os.path.join(O, extra), errors="replace")))
        if len(erows) > 2 and erows[0] == hdr:
            shutil.copy(os.path.join(O, extra), os.path.join(P, extra))
            for r in erows[2:]:
                if len(r) > ki:
                    seen[r[ki]] = r
    out = []
    for k, r in seen.items():
        e = OrderedDict(kernel=k)
        for w in WANT:
            if w in hdr:
                e[w] = r[hdr.index(w)] + " " + units[hdr.index(w)]
        out.append(e)
        short = ("redo" if "detect_fused" in k else "seg" if "seg_" in k else "filter" if "filter_" in k else "nms" if "nms_kernel" in k
                 else "conf_cut" if "conf_cut" in k else k[:20])
        try:
            def b(m):
                v = float(r[hdr.index(m)].replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index(m)], 1)
            t = traffic.setdefault(f"cfg{c}", OrderedDict())
            t[short] = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            t["images_per_launch"] = PROF_BS[c]
            t["source"] = f"profiles/{R}_ncu_full_cfg{c}_raw.csv (ncu --set full, bs {PROF_BS[c]}, last capture of each kernel)"
        except Exception:
            pass
    summary[f"cfg{c}"] = out
# ---- Detect head conv (tcgen05): timing log + --set full capture of the three levels ------------------------------------
for name in (f"{R}_detconv.log",):
    if os.path.exists(os.path.join(O, name)):
        shutil.copy(os.path.join(O, name), os.path.join(P, name))
raw = os.path.join(O, f"{R}_ncu_full_detconv_raw.csv")
if os.path.exists(raw) and os.path.getsize(raw) > 0:
    rows = list(csv.reader(open(raw, errors="replace")))
    hdr, units = rows[0], rows[1]
    shutil.copy(raw, os.path.join(P, f"{R}_ncu_full_detconv_raw.csv"))
    out = []
    for r in rows[2:]:
        if len(r) > hdr.index("Kernel Name"):
            e = OrderedDict(kernel=r[hdr.index("Kernel Name")])
            for w in WANT + ["sm__inst_executed_pipe_tensor.sum"]:
                if w in hdr:
                    e[w] = r[hdr.index(w)] + " " + units[hdr.index(w)]
            out.append(e)
    summary["detconv_fp16_bs64_levels_80_40_20"] = out
# ---- source pages kept as evidence (gzip, small) --------------------------------------------------------------------------
for name in os.listdir(O):
    if name.startswith(f"{R}_ncu_src_") and name.endswith(".csv.gz") and os.path.getsize(os.path.join(O, name)) < 400 * 1024:
        os.makedirs(os.path.join(P, "src_pages"), exist_ok=True)
        shutil.copy(os.path.join(O, name), os.path.join(P, "src_pages", name))
if summary:
    json.dump(summary, open(os.path.join(P, f"{R}_ncu_full_summary.json"), "w"), indent=1)
    traffic["note"] = ("PER LAUNCH: dram__bytes_read.sum + dram__bytes_write.sum of one launch at `images_per_launch`, from ncu "
                       "--set full; bench.py reports it as roofline.traffic only when its own batch has that size.  Writes still "
                       "dirty in the 126 MB L2 at kernel end are not counted by ncu")
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print("profiles/ updated:", sorted(os.listdir(P)))
