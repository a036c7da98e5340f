"""tools/ncu_traffic.py -- pull the DRAM traffic per launch of the dominant kernels out of `ncu --set full` reports and write
profiles/r2_traffic.json, which bench.py quotes as `roofline.traffic` (so the number in the bench line is a counter of a committed
capture keyed to a commit, not a literal in the source).

    python tools/ncu_traffic.py gemv_m1_12288=gpurun_out/r2_slab_prof.ncu-rep:gemv_slab [gemm_m4096_12288=...:gemm_ts] ...
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(rep, kernel_substr):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    want = {"dram__bytes_read.sum": None, "dram__bytes_write.sum": None, "gpu__time_duration.sum": None}
    idx = {k: hdr.index(k) for k in want}
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3}
    picked = [r for r in rows[2:] if kernel_substr in r[name_i]]
    if not picked:
        raise SystemExit(f"no kernel matching {kernel_substr!r} in {rep}")
    r = picked[-1]   # the last captured launch (warm instruction cache)
    vals = {k: float(r[i].replace(",", "")) * scale.get(units[i], 1.0) for k, i in idx.items()}
    return {"dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], "dram_read": vals["dram__bytes_read.sum"],
            "dram_write": vals["dram__bytes_write.sum"], "ncu_duration_us": vals["gpu__time_duration.sum"], "kernel": r[name_i][:120]}


def main():
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        rep, kern = rest.rsplit(":", 1)
        d = read(rep, kern)
        d["source"] = f"ncu --set full --clock-control none, {os.path.basename(rep)}, after commit {commit} (tools/ncu_traffic.py)"
        data[key] = d
        print(key, d)
    json.dump(data, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
