"""Times the pieces of the AdaMVS stage-1 'prepare' span (cfg3): relayout, hypothesis planes, pair volumes (dot cache vs
coefficient cache), the four pair softmax regressions.  usage: python tools/stage1_parts.py"""
import sys

import torch

sys.path.insert(0, ".")
from deep3d_aerial_b200 import sweep, synth  # noqa: E402

dev = torch.device("cuda", 0)
rig = synth.make_rig(num_views=5)
c, d, h, w = 32, 48, 688, 464
g = torch.Generator().manual_seed(3)
feats = torch.randn(5, c, h, w, generator=g).to(dev)
pose = sweep.relative_poses(torch.from_numpy(rig.proj(4)).to(dev))
rays = sweep.rays_for(pose, h, w)
logits = (4.0 * torch.randn(4, d, h, w, generator=g)).to(dev)


def timed(name, fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("%-46s %.3f ms" % (name, a.elapsed_time(b) / n))


tex = sweep.to_texels(feats)
hyps = sweep.depth_samples(sweep.SAMPLES_RANGE, d, (h, w), device=dev, dmin=rig.dmin, dmax=rig.dmax)
timed("relayout, 5 maps", lambda: sweep.to_texels(feats, out=tex))
timed("hypothesis planes [48,688,464]", lambda: sweep.depth_samples(sweep.SAMPLES_RANGE, d, (h, w), device=dev, dmin=rig.dmin, dmax=rig.dmax))
for v in (0, 48):
    timed("pair volumes, variant %d" % v, lambda v=v: sweep.cost_volume(tex, pose, hyps, sweep.AGG_PAIR_MEAN, rays=rays, variant=v))
timed("4 pair softmax regressions", lambda: [sweep.depth_regress(logits[k], hyps, want_index=False) for k in range(4)])
timed("1 pair softmax regression", lambda: sweep.depth_regress(logits[0], hyps, want_index=False))
