"""ORACLE — TEST INFRASTRUCTURE ONLY.

Golden vectors for SURVEY.md §8 row f3 (depth-map fusion consistency check), made by running the LIVE reference
`fuse/consistency_check_n.py:ConsistencyChecker.check` (imported from /root/reference, never copied) on seeded
synthetic scenes.  Run here, in the authoring container:  `python -m oracle.make_golden_fuse`.

The reference imports `cupy`, which is not in this image: `sys.modules['cupy']` is bound to a shim that forwards
every attribute to numpy (`cp.int` -> int, `cp.asnumpy` -> np.asarray).  numpy raises on the out-of-bounds
indices CuPy would wrap, so the scenes use padded source maps (`synth.fusion_scene(src_pad=...)`, no invalid
depths): every projection lands inside.  The accumulation around `check` is the reference's
`fusion_3d_normal.py:449-455, 525-537` restated line by line in `_accumulate` (that file needs `IO.*`,
`tools.*`, matplotlib and a parsed command line to import).

Writes tests/golden/fuse_pair.npz and tests/golden/fuse_view.npz.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from deep3d_aerial_b200 import synth  # noqa: E402

SCENE = dict(num_src=3, height=40, width=56, seed=5, src_pad=12, invalid=0.0)
THRESHOLDS = dict(position_threshold=1.0, depth_threshold=0.01, normal_threshold=10.0, confidence_threshold=0.2)


def load_live_checker():
    shim = types.ModuleType("cupy")
    shim.__getattr__ = lambda name: getattr(np, name)          # module-level __getattr__ (PEP 562)
    shim.int = int
    shim.asnumpy = np.asarray
    sys.modules["cupy"] = shim
    sys.path.insert(0, "/root/reference/fuse")
    import consistency_check_n
    return consistency_check_n.ConsistencyChecker


def main():
    Checker = load_live_checker()
    chk = Checker(THRESHOLDS["position_threshold"], THRESHOLDS["depth_threshold"], THRESHOLDS["normal_threshold"],
                  THRESHOLDS["confidence_threshold"])
    sc = synth.fusion_scene(**SCENE)
    d, n, k, e, prob = sc["ref"]
    pair, view = {}, {}
    height, width = d.shape
    # fusion_3d_normal.py:443-455, 466
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    xa, ya, da = x_ref.reshape([-1]), y_ref.reshape([-1]), d.reshape([-1])
    cam = np.matmul(np.linalg.inv(k), np.vstack((xa, ya, np.ones_like(xa))) * da)
    world = np.matmul(np.linalg.inv(e), np.vstack((cam, np.ones_like(xa))))[:3]
    all_xyz = world.reshape([-1, height, width]).astype(np.float32)
    conf = 0 + np.ones_like(all_xyz)
    count = 0 + np.ones([height, width], dtype=np.int32)
    for s, (ds, ns, ks, es) in enumerate(sc["src"]):
        mask, rep, removed, xyz, angle = chk.check(d, n, k, e, ds, ns, ks, es, prob)
        for name, arr in (("mask", mask), ("depth_reprojected", rep), ("depth_src_out", removed), ("xyz", xyz),
                          ("angle", angle)):
            pair["%d.%s" % (s, name)] = arr
        count += mask.astype(np.int32)                           # :525-527
        all_xyz += (angle * xyz).astype(np.float32)
        conf += angle
    view["count"] = count
    view["xyz"] = (all_xyz / conf).astype(np.float32)            # :533-534
    for m in (2, 3, 4):
        view["final_mask.%d" % m] = np.array(count >= m)         # :537
    np.savez_compressed(os.path.join(OUT, "fuse_pair.npz"), **pair)
    np.savez_compressed(os.path.join(OUT, "fuse_view.npz"), **view)
    print("masks", [float(pair["%d.mask" % s].mean()) for s in range(len(sc["src"]))], "count", np.bincount(count.ravel()))


if __name__ == "__main__":
    main()
