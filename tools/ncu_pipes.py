"""Executed warp instructions per source function split by issue pipe (ALU / FMA / LSU / other).
    ncu -i REPORT --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_pipes.py src.csv units
"""
import collections
import csv
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_regions import function_ranges  # noqa: E402

ALU = {"LOP3", "SHF", "PRMT", "ISETP", "SEL", "IADD3", "VIADD", "LEA", "PLOP3", "POPC", "FLO", "BREV",
       "IABS", "IMNMX", "VIMNMX", "MOV", "CS2R", "P2R", "R2P", "VOTE", "SGXT", "BMSK", "ISCADD"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
LSU = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "LD", "ST", "SHFL", "ATOMS", "ATOM", "RED", "LDSM"}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2])
    fname = hdr = cur = None
    agg = collections.defaultdict(lambda: collections.Counter())
    cache = {}
    seen_sass = False
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            ci = hdr.index("Instructions Executed")
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]))
            continue
        try:
            n = int(r[ci])
        except (ValueError, IndexError):
            continue
        if n == 0:
            continue
        op = r[3].strip().split()
        if not op:
            continue
        op = op[1] if op[0].startswith("@") else op[0]
        op = op.split(".")[0]
        pipe = "alu" if op in ALU else "fma" if op in FMA else "lsu" if op in LSU else "other"
        f, ln = cur
        if f not in cache:
            try:
                cache[f] = function_ranges(f)
            except OSError:
                cache[f] = []
        name = f.split("/")[-1]
        for fn, a, b in cache[f]:
            if a <= ln <= b:
                name = f.split("/")[-1].split(".")[0] + ":" + fn
        agg[name][pipe] += n
        agg[name]["all"] += n
    tot = collections.Counter()
    for c in agg.values():
        tot.update(c)
    print("%-40s %8s %8s %8s %8s %8s" % ("function", "all", "alu", "fma", "lsu", "other"))
    print("%-40s %8.1f %8.1f %8.1f %8.1f %8.1f" % ("TOTAL", *[tot[k] / units for k in ("all", "alu", "fma", "lsu", "other")]))
    for name, c in sorted(agg.items(), key=lambda kv: -kv[1]["all"])[:30]:
        print("%-40s %8.1f %8.1f %8.1f %8.1f %8.1f" % (name[:40], *[c[k] / units for k in ("all", "alu", "fma", "lsu", "other")]))


if __name__ == "__main__":
    main()
