#!/usr/bin/env python
"""profiles/rNN_sass_summary.md: per kernel of libmf_b200.so -- registers, stack (spills), shared memory, and the counts of the SASS
mnemonics that show which hardware path it uses (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM),
UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = cp.async.bulk, HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async.

    python scripts/sass_summary.py [round tag, default r02]   (needs only cuobjdump: runs in the build container)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mere_fusion_b200", "libmf_b200.so")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
WATCH = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCIMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "LDSM", "LDGSTS",
         "SYNCS", "LDG", "STG", "LDS", "STS")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.search(r"Function ([^:]+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    counts = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    fn = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            total[fn] += 1
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    counts[fn][w] += 1
                    break
    names = demangle(sorted(total))
    rows = []
    for fn in sorted(total, key=lambda f: -total[f]):
        c = counts[fn]
        utc = sum(c[k] for k in c if k.startswith("UTC") and k.endswith("MMA"))
        reg, stack, shared = usage.get(fn, (0, 0, 0))
        short = re.sub(r"\(.*", "", names.get(fn, fn))
        rows.append((short, total[fn], reg, stack, shared, utc, c["LDTM"] + c["STTM"], c["UTMALDG"] + c["UTMASTG"], c["UBLKCP"], c["HMMA"], c["LDSM"], c["LDGSTS"]))
    out = [f"# SASS summary of libmf_b200.so ({TAG}; scripts/sass_summary.py, cuobjdump -sass / -res-usage, sm_100a)", "",
           "UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMA* = TMA tensor copies, UBLKCP = cp.async.bulk, HMMA = mma.sync, "
           "LDSM = ldmatrix, LDGSTS = cp.async.  STACK > 0 = local memory (spills or out-of-line calls).", "",
           "| kernel | SASS instr | regs | stack B | static smem B | UTC*MMA | LDTM/STTM | UTMALDG/STG | UBLKCP | HMMA | LDSM | LDGSTS |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        out.append("| " + " | ".join(str(x) for x in r) + " |")
    path = os.path.join(ROOT, "profiles", f"{TAG}_sass_summary.md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()
