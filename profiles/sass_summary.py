"""Opcode summary of the built library's SASS (run in the build container, no GPU needed):
    python profiles/sass_summary.py > profiles/r2_sass_opcodes.txt
Per kernel: instruction count and the counts of the mnemonics that evidence the Blackwell-specific paths (UBLKCP = TMA bulk
copy, UTMALDG = TMA tensor-map copy, SYNCS = mbarrier, FFMA2 / FMUL2 = packed fp32x2 arithmetic, REDUX = warp reduction, MATCH = warp match, UCGABAR =
cluster barrier, UTCHMMA / LDTM / UTCBAR = tcgen05.mma / tcgen05.ld / tcgen05.commit, LDG/STG widths)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "desenet_b200", "lib", "libdesenet_b200.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "REDUX", "MATCH", "UCGABAR", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS.128",
         "STS.128", "UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "ATOMS", "ATOMG", "RED", "MUFU", "BAR.SYNC", "SHFL", "VOTE", "DMUL", "DFMA", "LDGSTS", "ERRBAR", "MEMBAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kern, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            kern[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            op = m.group(1)
            kern[name]["_total"] += 1
            for w in WATCH:
                if w.count(".") >= 2:                      # LDG.E.128: family + width, whatever cache-hint modifiers sit between
                    fam, width = w.split(".")[0], w.split(".")[-1]
                    if op.split(".")[0] == fam and width in op.split("."):
                        kern[name][w] += 1
                elif op == w or op.startswith(w + "."):
                    kern[name][w] += 1
    print(f"# SASS opcode summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); columns = static instruction counts")
    tot = collections.Counter()
    for k, c in kern.items():
        hot = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{k[:110]:110s} total={c['_total']:6d} {hot}")
        tot.update(c)
    print("# whole library:", " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]), f"total={tot['_total']}")


if __name__ == "__main__":
    main()
