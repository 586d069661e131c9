"""Record the DRAM traffic of one captured kernel launch in profiles/r2_traffic.json.
    python tools/ncu_traffic.py REPORT.ncu-rep "PLAN TEXT" NxHxW
PLAN TEXT is the kernel description of sl_env_step_obs_plan without the grid, e.g.
"k_env_step_warp<Narrow, goals in place>".  The file is stamped with the hash of the
CUDA sources, so bench.py drops the figures once the kernels change.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import kernel_source_hash
    rep, plan, shape = sys.argv[1:4]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(vals[i]) * scale
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    data = {"source_sha1": kernel_source_hash(), "entries": {}}
    if os.path.exists(path):
        old = json.load(open(path))
        if old.get("source_sha1") == data["source_sha1"]:
            data = old
    data["entries"]["%s|%s" % (plan, shape)] = {
        "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
        "duration_us": float(vals[hdr.index("gpu__time_duration.sum")]),
        "report": os.path.basename(rep)}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data["entries"], indent=1))


if __name__ == "__main__":
    main()
