"""Per-kernel SASS opcode summary of librtg_splat_b200.so (evidence for the TMA / cp.async / atomics claims of DESIGN.md):
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
Counts static instructions (cuobjdump -sass); TMA bulk copies show as UBLKCP, mbarrier waits as SYNCS, cp.async as LDGSTS,
reductions as RED / REDG, tensor-map TMA would show as UTMALDG (none: there is no 2-D tile on this path)."""
import os, re, subprocess, sys
from collections import Counter, defaultdict
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(here, "rtg_slam_b200", "librtg_splat_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "REDG", "RED", "ATOMG", "ATOMS", "MUFU.EX2", "MUFU.RCP", "SHFL", "VOTE", "LDS", "STS",
         "LDG", "STG", "BAR", "HMMA", "UTCHMMA", "DFMA", "DADD"]
per = defaultdict(Counter)
total = Counter()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
                break
def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return r.split("(")[0].replace("rtg::", "")
print("# static SASS opcode counts per kernel of librtg_splat_b200.so (sm_100a), tools/sass_opcodes.py")
print("# " + " ".join(f"{w:>8s}" for w in ["insts"] + WATCH) + "  kernel")
for k in sorted(total, key=lambda k: -total[k]):
    print("  " + " ".join(f"{v:8d}" for v in [total[k]] + [per[k][w] for w in WATCH]) + "  " + demangle(k))
