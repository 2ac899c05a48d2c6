"""Static SASS instruction counts per kernel of the built library (the evidence for TMA / packed-FP32 / mbarrier use):

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "glimpse_b200", "libglimpse_b200.so")
KEEP = ("UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "VIMNMX", "DFMA", "DMUL", "DADD", "MUFU", "REDUX", "ATOMS", "BAR", "LDS", "STS",
        "LDG", "STG")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, cnt, tot = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            tot[cur] += 1
            if op in KEEP:
                cnt[cur][op] += 1
    names = {k: subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0].replace("gb::", "") for k in tot}
    print("# SASS instruction counts per kernel (static), cuobjdump -sass of glimpse_b200/libglimpse_b200.so built by glimpse_b200/build.py for sm_100a")
    print("# UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = cp.async.bulk.tensor.2d (tiled TMA), SYNCS = mbarrier ops, FFMA2/FADD2 = packed FP32 "
          "(sm_100), VIMNMX = integer min/max (u16x2 median network)")
    print("%-46s %7s " % ("kernel", "total") + " ".join("%7s" % c for c in KEEP))
    for k in sorted(tot, key=lambda k: -tot[k]):
        print("%-46s %7d " % (names[k][:46], tot[k]) + " ".join("%7d" % cnt[k][c] for c in KEEP))


if __name__ == "__main__":
    sys.exit(main())
