"""Stall samples aggregated by opcode: python tools/ncu_stall_by_op.py rep [reason-substring]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
execs = collections.Counter()
for r in body:
    src = r[idx["Source"]].strip().split()
    if not src: continue
    op = src[1] if src[0].startswith("@") else src[0]
    op = op.split(".")[0]
    execs[op] += int(r[idx["Instructions Executed"]] or 0)
    for h in stall_cols:
        agg[op][h[6:]] += int(r[idx[h]] or 0)
tot = sum(sum(c.values()) for c in agg.values())
te = sum(execs.values())
print("%-10s %7s %7s  top reasons (samples per executed warp-inst = cycles)" % ("op", "exec%", "smpl%"))
for op, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:22]:
    s = sum(c.values())
    print("%-10s %6.1f%% %6.1f%%  %s" % (op, 100.0 * execs[op] / te, 100.0 * s / tot,
          ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in c.most_common(4))))
