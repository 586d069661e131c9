"""SASS opcode histogram of libsafelife_b200.so (cuobjdump -sass), whole library and
per step kernel; the evidence for which hardware paths the code uses (UBLKCP = 1-D
bulk TMA copy, SYNCS = mbarrier, BAR = CTA / named barriers, no tensor-core ops).
    python tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = re.compile(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)")
KEY = ("UBLKCP", "SYNCS", "UTMACMDFLUSH", "UTMALDG", "UTMASTG", "BAR", "ATOMS", "VOTE", "SHFL", "LOP3",
       "PRMT", "SHF", "IMAD", "POPC", "FLO", "LDS", "STS", "LDG", "STG", "LD", "ST", "HMMA", "IMMA",
       "UTCHMMA", "UTCQMMA", "TCGEN05")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "safelife_b200", "libsafelife_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
    total = collections.Counter()
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        if "Function :" in line:
            name = subprocess.run(["c++filt", line.split("Function :")[1].strip()], stdout=subprocess.PIPE,
                                  text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = LINE.match(line)
        if m and cur is not None:
            total[m.group(1)] += 1
            cur[m.group(1)] += 1
    print("# %s: %d SASS instructions in %d kernels (sm_100a)" % (os.path.basename(lib), sum(total.values()), len(per)))
    print("# whole library, by opcode")
    for op, n in total.most_common():
        print("%8d %s" % (n, op))
    tensor = [op for op in total if "MMA" in op or op.startswith("UTC") or "TCGEN" in op]
    print("# tensor-core opcodes: %s" % (", ".join(tensor) if tensor else "none (integer / bitwise path)"))
    print("# per kernel: total | " + " ".join(KEY[:21]))
    for name, c in per.items():
        print("%-110s %6d | %s" % (name[:110], sum(c.values()), " ".join("%s=%d" % (k, c[k]) for k in KEY[:21] if c[k])))


if __name__ == "__main__":
    main()
