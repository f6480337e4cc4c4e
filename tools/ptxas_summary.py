"""Compact table of ptxas -v output: kernel, registers, spills, smem.  Usage:
   python -m deep3d_aerial_b200.build --force --ptxas 2>&1 | python tools/ptxas_summary.py [filter]"""
import re
import subprocess
import sys

flt = sys.argv[1] if len(sys.argv) > 1 else ""
cur = None
rows = []
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1), "spill": 0, "regs": 0, "smem": 0}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes spill stores", line)
    if m:
        cur["spill"] = int(m.group(1))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(m2.group(1)) if m2 else 0
names = subprocess.run(["c++filt"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.split("\n")
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n).replace("d3d::", "").replace("void ", "")
    if flt in n:
        print("%-60s regs %3d  spill %3d  smem %5d" % (n, r["regs"], r["spill"], r["smem"]))
