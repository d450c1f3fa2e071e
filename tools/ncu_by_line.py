"""Aggregate the per-SASS-instruction counts of an ncu report by CUDA source line.

ncu's CSV export of the source page carries metrics only in the SASS view; this joins it, instruction by
instruction, with `nvdisasm -g` of the same kernel (built with -lineinfo) and sums "Instructions Executed" and
stall samples per source line (inlined code is attributed to the innermost line).

    python tools/ncu_by_line.py gpurun_out/prof_r01.ncu-rep preprocess_fwd_kernel preprocess [top_n]
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kern, unit = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(here, "rtg_slam_b200", "librtg_splat_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", unit, so], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout
# instruction -> line list for the kernel's .text section
lines, cur, inside = [], None, False
for l in sass.splitlines():
    if l.startswith(".text."):
        inside = kern in l
        continue
    if l.startswith("//----") and ".text." in l:
        inside = False
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k",
                         "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hdr = rows[1]
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        if data: break  # a second launch of the same kernel follows: keep the first
        continue
    data.append((int(r[ie]), int(r[isamp])))
print(f"# {kern}: {len(data)} SASS instructions in the report, {len(lines)} in the disassembly")
n = min(len(data), len(lines))
agg = defaultdict(lambda: [0, 0])
for (e, s), ln in zip(data[:n], lines[:n]):
    agg[ln][0] += e
    agg[ln][1] += s
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
src_cache = {}
def src(ln):
    if ln is None: return ""
    f = ln[0]
    if f not in src_cache:
        p = os.path.join(here, "rtg_slam_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    t = src_cache[f]
    return t[ln[1] - 1].strip()[:90] if 0 < ln[1] <= len(t) else ""
print(f"{'inst %':>7} {'samp %':>7}  line")
for ln, (e, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100*e/tot:6.1f}% {100*s/tots:6.1f}%  {ln[0] if ln else '?'}:{ln[1] if ln else 0}  {src(ln)}")
