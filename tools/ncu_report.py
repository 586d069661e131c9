"""Text summary of one ncu report: key metrics, stall reasons, hottest source lines.
    python tools/ncu_report.py REPORT.ncu-rep [units] > summary.txt
`units` divides the per-line instruction counts (e.g. the number of envs).
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), stdout=subprocess.PIPE,
                          stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2]
    print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-66s %s %s" % (k, vals[i], unit[i]))
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(vals[i]), h))
            except ValueError:
                pass
    print("# warp stall reasons (warps per issue-active cycle)")
    for v, h in sorted(stalls, reverse=True)[:10]:
        print("  %8.3f %s" % (v, h))
    # hottest source lines
    rows = list(csv.reader(ncu(rep, "--page", "source", "--print-source", "cuda,sass", "--csv").splitlines()))
    fname = hdr = cur = None
    lines = collections.OrderedDict()
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
            try:
                cur = (fname, int(r[0]), r[1].strip()[:90])
            except ValueError:
                cur = None
            continue
        if cur is None:
            continue
        try:
            e = lines.setdefault(cur, [0.0, 0.0])
            e[0] += float(r[ci])
            e[1] += float(r[cs])
        except (ValueError, IndexError):
            pass
    tot_i = sum(v[0] for v in lines.values()) or 1.0
    tot_s = sum(v[1] for v in lines.values()) or 1.0
    by_file = collections.Counter()
    for (f, l, t), v in lines.items():
        by_file[f] += v[0]
    print("# executed warp instructions per unit (%g units) by source file" % units)
    for f, n in by_file.most_common(8):
        print("  %10.1f  %s" % (n / units, f))
    print("# hottest source lines: instr/unit  inst%  stall-sample%  file:line  text")
    for (f, l, t), v in sorted(lines.items(), key=lambda kv: -(kv[1][0] / tot_i + kv[1][1] / tot_s))[:int(sys.argv[3]) if len(sys.argv) > 3 else 28]:
        print("  %9.1f %5.1f %5.1f  %s:%d  %s" % (v[0] / units, 100 * v[0] / tot_i, 100 * v[1] / tot_s, f, l, t))


if __name__ == "__main__":
    main()
