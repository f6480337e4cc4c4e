"""A/B of sweep-kernel variants on one cascade-stage shape (CUDA events; results compared with variant 0 bit for bit).
usage: python tools/stage_ab.py <stage 1|2|3> <variant> [<variant> ...]"""
import sys

import torch

sys.path.insert(0, ".")
from deep3d_aerial_b200 import sweep, synth  # noqa: E402

mode = sweep.AGG_WEIGHTED_PRODUCT
if sys.argv[1] == "variance":                  # python tools/stage_ab.py variance <stage> <variant> ...
    mode = sweep.AGG_VARIANCE
    del sys.argv[1]
stage = int(sys.argv[1])
variants = [int(v) for v in sys.argv[2:]] or [0]
scale, c, d, ratio = {1: (4, 32, 48, 4.0), 2: (2, 16, 32, 2.0), 3: (1, 8, 8, 1.0)}[stage]
dev = torch.device("cuda", 0)
rig = synth.make_rig(num_views=5)
h, w = 2752 // scale, 1856 // scale
g = torch.Generator().manual_seed(3)
feats = torch.randn(5, c, h, w, generator=g).to(dev)
tex = sweep.to_texels(feats)
pose = sweep.relative_poses(torch.from_numpy(rig.proj(scale)).to(dev))
if stage == 1:
    hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d).to(dev)
else:
    cur = synth.smooth_depth_map(rig, h, w, seed=1).to(dev)
    hyps = synth.per_pixel_hypotheses(cur, d, ratio * (rig.dmax - rig.dmin) / 384).contiguous()
weights = torch.rand(4, h, w, generator=g).to(dev)
rays = sweep.rays_for(pose, h, w)
base = None
for v in [0] + [x for x in variants if x != 0]:
    out = sweep.cost_volume(tex, pose, hyps, mode, weights=(weights if mode == sweep.AGG_WEIGHTED_PRODUCT else None), plane_major=True, rays=rays, variant=v)
    for _ in range(3):
        sweep.cost_volume(tex, pose, hyps, mode, weights=(weights if mode == sweep.AGG_WEIGHTED_PRODUCT else None), plane_major=True, rays=rays, variant=v, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        sweep.cost_volume(tex, pose, hyps, mode, weights=(weights if mode == sweep.AGG_WEIGHTED_PRODUCT else None), plane_major=True, rays=rays, variant=v, out=out)
    b.record()
    torch.cuda.synchronize()
    if base is None:
        base = out.clone()
    err = float((out - base).norm() / base.norm())
    print("stage %d (C=%d D=%d %dx%d) variant %2d: %.3f ms  identical to variant 0: %s (rel diff %.2e)"
          % (stage, c, d, h, w, v, a.elapsed_time(b) / 10, bool(torch.equal(out, base)), err))
