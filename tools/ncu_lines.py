"""
Aggregate an ncu report's executed instructions and stall samples per CUDA
source line.

    ncu -i REPORT.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_lines.py src.csv [units] [top]

`units` divides the instruction counts (e.g. the number of envs = warps).
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    rows = list(csv.reader(open(path)))
    fname, lines, cur, hdr = None, collections.OrderedDict(), None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]))
            lines.setdefault(cur, [0, 0, r[1].strip()])
            continue
        try:
            lines[cur][0] += int(r[ci])
            lines[cur][1] += int(r[cs])
        except (ValueError, IndexError):
            pass
    tot_i = sum(v[0] for v in lines.values())
    tot_s = sum(v[1] for v in lines.values())
    print("total instructions %.1f per unit, samples %d" % (tot_i / units, tot_s))
    print("%9s %6s %6s  %s" % ("instr/u", "inst%", "smpl%", "line"))
    ranked = sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]
    for (f, ln), (ni, ns, src) in ranked:
        print("%9.1f %6.1f %6.1f  %s:%d  %s" % (ni / units, 100.0 * ni / max(tot_i, 1),
                                                100.0 * ns / max(tot_s, 1), f, ln, src[:90]))


if __name__ == "__main__":
    main()
