"""Top stall sites of an ncu report's SASS source page: python tools/ncu_source.py rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {h: sum(int(r[idx[h]] or 0) for r in body) for h in stall_cols}
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
top = sorted(body, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]
for r in sorted(top, key=lambda r: int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") else 0):
    s = int(r[idx["# Samples"]] or 0)
    reasons = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print("%6s %5.2f%%  %-70s %s" % (r[idx["Address"]][-5:], 100.0 * s / tot, r[idx["Source"]][:70],
                                      " ".join("%s:%d" % (nm, v) for v, nm in reasons if v)))
