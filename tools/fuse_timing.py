"""Times d3d_consistency_fuse at the full depth-map size (1856 x 2752, S source views) with CUDA events.
usage: python tools/fuse_timing.py [S] [reps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from deep3d_aerial_b200 import fusion, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t = time.time()
sc = synth.fusion_scene(num_src=S, height=2752, width=1856, seed=11, focal=4000.0)
print("scene %.1f s" % (time.time() - t))
dev = torch.device("cuda", 0)
d, n, k, e, prob = sc["ref"]
up = lambda x: torch.from_numpy(x).to(dev)  # noqa: E731
geom = torch.from_numpy(fusion.pair_geometry(k, e, [v[2] for v in sc["src"]], [v[3] for v in sc["src"]])).to(dev)
args = (up(d), up(n), up(prob), geom, [up(v[0]) for v in sc["src"]], [up(v[1]) for v in sc["src"]])
kw = dict(normal_threshold_cos=0.98480775, min_consistent=4)
for per_source, update in ((False, True), (False, False), (True, True)):
    out = fusion.fuse_view(*args, per_source=per_source, update_sources=update, **kw)
    for _ in range(3):
        fusion.fuse_view(*args, per_source=per_source, update_sources=update, out=out, **kw)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fusion.fuse_view(*args, per_source=per_source, update_sources=update, out=out, **kw)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    px = 2752 * 1856
    print("S=%d per_source=%s update_sources=%s: %.3f ms per reference view, %.1f Mpixel-pairs/s, final mask %.3f"
          % (S, per_source, update, ms, px * S / ms / 1e3, float(out["final_mask"].float().mean())))
