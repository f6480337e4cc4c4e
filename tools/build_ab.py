"""Builds an A/B copy of libd3dsweep.so with extra -D macros on chosen translation units (everything else is linked from the
regular build's objects): python tools/build_ab.py <tag> <-DNAME=VALUE ...> -- <csrc glob ...>
-> deep3d_aerial_b200/_ab/libd3dsweep_<tag>.so; run with D3D_SWEEP_LIB=<that path> (deep3d_aerial_b200/_lib.py)."""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import build as B  # noqa: E402

tag = sys.argv[1]
cut = sys.argv.index("--")
defs, pats = sys.argv[2:cut], sys.argv[cut + 1:]
B.build()
out_dir = os.path.join(B.HERE, "_ab")
obj_dir = os.path.join(out_dir, "obj_" + tag)
os.makedirs(obj_dir, exist_ok=True)
mine = sorted({s for p in pats for s in glob.glob(os.path.join(B.CSRC, p))})
nvcc = B._nvcc()


def run(src):
    obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc] + B.ARCH + [f for f in B.FLAGS if f != "--use_fast_math=false"] + defs + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stdout + r.stderr)
    return obj


with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(run, mine))
names = {os.path.basename(o) for o in objs}
rest = [o for o in glob.glob(os.path.join(B.OBJ, "*.o")) if os.path.basename(o) not in names]
lib = os.path.join(out_dir, "libd3dsweep_%s.so" % tag)
subprocess.check_call([nvcc] + B.ARCH + ["-shared", "-o", lib] + objs + rest + ["-cudart", "static"])
print(lib)
