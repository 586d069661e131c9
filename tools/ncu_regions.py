"""Executed warp instructions per unit, grouped by source function (line ranges
found by scanning the source for function headers).
    python tools/ncu_regions.py src.csv units
"""
import collections
import csv
import re
import sys


def function_ranges(path):
    ranges, cur = [], None
    pat = re.compile(r"^(?:template.*)?\s*(?:__global__|__device__|inline|static).*?\b(\w+)\s*\(")
    for no, line in enumerate(open(path), 1):
        m = pat.match(line)
        if m and not line.strip().startswith("//") and ";" not in line.split("(")[0]:
            if cur:
                ranges.append((cur[0], cur[1], no - 1))
            cur = (m.group(1), no)
    if cur:
        ranges.append((cur[0], cur[1], 10 ** 9))
    return ranges


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2])
    fname = hdr = cur = None
    lines = collections.OrderedDict()
    paths = {}
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
            lines.setdefault(cur, 0)
            continue
        try:
            lines[cur] += int(r[ci])
        except (ValueError, IndexError):
            pass
    agg = collections.Counter()
    cache = {}
    for (f, ln), n in lines.items():
        if f not in cache:
            try:
                cache[f] = function_ranges(f)
            except OSError:
                cache[f] = []
        name = f.split("/")[-1]
        for fn, a, b in cache[f]:
            if a <= ln <= b:
                name = f.split("/")[-1] + ":" + fn
        agg[name] += n
    tot = sum(agg.values())
    print("total %.1f instr/unit" % (tot / units))
    for k, v in agg.most_common(25):
        print("%8.1f  %5.1f%%  %s" % (v / units, 100.0 * v / tot, k))


if __name__ == "__main__":
    main()
