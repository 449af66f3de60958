"""Ranks the CUDA source lines of one kernel by executed warp instructions and stall samples, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K --launch-count 1 > f.csv`.
python profiles/experiments/src_hot.py f.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr, out, cur = None, None, {}, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= max(ie, ss):
        continue
    if r[0]:                                  # a source line (its own row carries the aggregated metrics)
        cur = (fname, int(r[0]), r[1][:100])
        try:
            out[cur] = [float(r[ie] or 0), float(r[ss] or 0)]
        except ValueError:
            out[cur] = [0.0, 0.0]
tot = sum(v[0] for v in out.values()) or 1
tots = sum(v[1] for v in out.values()) or 1
print("total warp instructions", int(tot), "samples", int(tots))
for k, v in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:16s} {k[1]:4d} {100 * v[0] / tot:5.1f}% inst {100 * v[1] / tots:5.1f}% smp | {k[2]}")
