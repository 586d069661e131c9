"""Per-SASS-instruction executed counts and stall samples of the first kernel in
    ncu -i REPORT.ncu-rep --page source --print-source sass --csv > sass.csv
    python tools/ncu_sass.py sass.csv [units] [top] [--dump]
"""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, kernels = None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernels.append((r[1], []))
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr) and kernels:
            kernels[-1][1].append(r)
    return hdr, kernels


def main():
    hdr, kernels = load(sys.argv[1])
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    name, data = kernels[0]
    tot_i = sum(int(r[ci]) for r in data)
    tot_s = sum(int(r[cs]) for r in data)
    print(name, "sass instructions", len(data), "executed/unit %.1f" % (tot_i / units), "samples", tot_s)
    if "--dump" in sys.argv:
        for k, r in enumerate(data):
            print("%5d %9.2f %6d  %s" % (k, int(r[ci]) / units, int(r[cs]), r[1].strip()))
        return
    ops = {}
    for r in data:
        op = r[1].strip().split()
        op = op[1] if op[0].startswith("@") else op[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ci])
    print("by opcode (executed/unit):")
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
        print("   %-10s %8.1f %5.1f%%" % (op, n / units, 100.0 * n / tot_i))
    print("top by samples:")
    order = sorted(range(len(data)), key=lambda k: -int(data[k][cs]))[:top]
    for k in order:
        r = data[k]
        print("%5d %9.2f %6d %5.1f%%  %s" % (k, int(r[ci]) / units, int(r[cs]),
                                             100.0 * int(r[cs]) / max(tot_s, 1), r[1].strip()))


if __name__ == "__main__":
    main()
