"""Executed-instruction histogram of an ncu report's SASS page, run-length grouped by execution count:
python tools/ncu_exec.py rep [--list]   (per-voxel figures need VOX=<voxels> in the environment)"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
vox = float(os.environ.get("VOX", "122585088"))
tot = sum(int(r[idx["Instructions Executed"]] or 0) for r in body)
print("total warp-inst %.3fG = %.1f per voxel; SASS lines %d" % (tot / 1e9, tot / vox, len(body)))
# group
groups = []
for i, r in enumerate(body):
    n = int(r[idx["Instructions Executed"]] or 0)
    if groups and abs(groups[-1][2] - n) <= 0.02 * max(n, 1):
        groups[-1][1] = i; groups[-1][3] += n; groups[-1][4] += int(r[idx["# Samples"]] or 0)
    else:
        groups.append([i, i, n, n, int(r[idx["# Samples"]] or 0)])
ts = sum(g[4] for g in groups)
for a, b, n, s, smp in groups:
    if s / tot > 0.004:
        print("lines %5d-%5d  n=%4d  exec/inst %.3e  total %5.2f/voxel (%4.1f%%)  samples %4.1f%%   %s" % (
            a, b, b - a + 1, n, s / vox, 100.0 * s / tot, 100.0 * smp / ts, body[a][idx["Source"]].strip()[:50]))
if "--list" in sys.argv:
    for i, r in enumerate(body):
        print(i, r[idx["Instructions Executed"]], r[idx["# Samples"]], r[idx["Source"]].strip())
