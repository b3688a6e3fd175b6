"""Development aid: per-source-line view of an ncu report (--set full --import-source on) of ONE kernel launch.

ncu's CSV source page is per SASS instruction and carries no line numbers; this joins it (by instruction offset) with the
line info `nvdisasm -gi` prints for the same function of the library the capture ran, and sums executed warp / thread
instructions and stall samples per innermost source line and per enclosing function region.

usage: ncu_source_lines.py <report.ncu-rep> <library.so> <mangled kernel name substring> [top N]"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=tmp, check=True, capture_output=True)
cubin = next(tmp.glob("*.cubin"))
sass = subprocess.run(["nvdisasm", "-gi", str(cubin)], capture_output=True, text=True).stdout.splitlines()
# the kernel's section
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kname in l)
lines_of = {}           # offset -> [(file, line), ...] innermost first
chain = []
fresh = True
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((Path(m.group(1)).name, int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        lines_of[int(m.group(1), 16)] = list(chain)
        fresh = True
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(csvtxt))
# a report with several launches: one section per launch, headed by "Kernel Name"; take the first whose demangled name
# matches the part of the mangled name after its length prefix (k_...)
want = re.search(r"k_[a-z0-9_]+", kname).group(0)
sect = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if sect:
    pick = next(i for i in sect if want + "(" in rows[i][1] or want + "<" in rows[i][1])
    end = next((j for j in sect if j > pick), len(rows))
    rows = rows[pick:end]
hdr = next(r for r in rows if r and r[0] == "Address")
data = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(data[0]["Address"], 16)
tot_w = tot_t = tot_s = 0
by_line = defaultdict(lambda: [0, 0, 0, 0])
by_outer = defaultdict(lambda: [0, 0, 0, 0])
for d in data:
    off = int(d["Address"], 16) - base
    w, t, s = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])
    ls = int(d.get("stall_long_sb", 0) or 0)
    tot_w += w; tot_t += t; tot_s += s
    ch = lines_of.get(off) or [("?", 0)]
    for key, agg in ((ch[0], by_line), (tuple(ch[:3]), by_outer)):
        a = agg[key]
        a[0] += w; a[1] += t; a[2] += s; a[3] += ls
print(f"total: {tot_w} warp instructions, {tot_t / max(tot_w, 1):.2f} lanes, {tot_s} samples")
print("---- by innermost line: % warp instr, lanes, % samples, % of samples long-scoreboard")
for key, a in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot_w * 100:5.1f}%  lanes {a[1] / max(a[0], 1):5.1f}  smp {a[2] / max(tot_s, 1) * 100:5.1f}%  lsb {a[3] / max(a[2], 1) * 100:4.0f}%   {key[0]}:{key[1]}")
print("---- by inline chain (3 innermost frames)")
for key, a in sorted(by_outer.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot_w * 100:5.1f}%  lanes {a[1] / max(a[0], 1):5.1f}  smp {a[2] / max(tot_s, 1) * 100:5.1f}%   " + " < ".join(f"{f}:{n}" for f, n in key))
