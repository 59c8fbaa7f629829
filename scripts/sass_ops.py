"""Counts the Blackwell-specific SASS mnemonics per kernel of libomniswarm_b200.so (cuobjdump -sass) and writes a
markdown table: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, STG.*256 = 256-bit stores, HMMA = legacy mma.sync (must be absent).
usage: python scripts/sass_ops.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "omni-swarm_b200", "csrc", "libomniswarm_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "STG.E.ENL2.256", "HMMA", "FFMA", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    dem = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    names = list(counts)
    d = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip(names, d))
    rows = []
    for k, c in counts.items():
        if not any(c[o] for o in OPS[:8]) and c["FFMA"] + c["DFMA"] < 50:
            continue
        name = re.sub(r"\(.*", "", dem.get(k, k)).replace("osb::", "")
        rows.append((name, c))
    rows.sort(key=lambda r: -(r[1]["UTCHMMA"] * 1000 + r[1]["UTMALDG"] * 10 + r[1]["FFMA"]))
    lines = ["# SASS op counts per kernel (libomniswarm_b200.so, sm_100a)", "",
             "`cuobjdump -sass` of the built library; static instruction counts (not executions).  UTCHMMA = `tcgen05.mma`, "
             "LDTM = `tcgen05.ld`, UTMALDG = TMA loads, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier operations, "
             "STG.E.ENL2.256 = 256-bit global stores.  HMMA (legacy `mma.sync`) does not occur.", "",
             "| kernel | " + " | ".join(OPS) + " |", "|---|" + "---|" * len(OPS)]
    for name, c in rows:
        lines.append(f"| `{name}` | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    tot = collections.Counter()
    for _, c in counts.items():
        tot.update(c)
    lines += ["", "Library totals: " + ", ".join(f"{o} x{tot[o]}" for o in OPS if tot[o])]
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
