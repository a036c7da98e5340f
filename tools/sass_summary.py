"""tools/sass_summary.py -- per-kernel counts of the SASS mnemonics that prove what a kernel is built from (B200_PROFILING.md:
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA tensor / bulk copies, LDGSTS = cp.async, HMMA / IMMA =
mma.sync, SYNCS = mbarrier ops), from `cuobjdump -sass` of the built library.  Writes profiles/r2_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bitblas_b200", "lib", "libbitblas_b200.so")
KEYS = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "HMMA", "IMMA",
        "LOP3", "PRMT", "LDS", "STS", "LDG", "STG", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    fn = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = re.sub(r"\(anonymous namespace\)::", "", fn)
            fn = re.sub(r"\(.*$", "", fn)[:110]
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            counts[fn]["total"] += 1
            for k in KEYS:
                if op == k or (k in ("UTCHMMA", "UTCIMMA", "UTCQMMA") and op.startswith(k)):
                    counts[fn][k] += 1
    path = os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
    with open(path, "w") as f:
        f.write("# cuobjdump -sass bitblas_b200/lib/libbitblas_b200.so : instruction counts per kernel (tools/sass_summary.py)\n")
        for fn, c in counts.items():
            if c["total"] < 50:
                continue
            f.write(fn + "\n    " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]) + "\n")
    print(path)


if __name__ == "__main__":
    main()
