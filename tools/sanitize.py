"""Small launches of every production sweep kernel family, for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck|racecheck|synccheck --error-exitcode 3 python tools/sanitize.py
Prints, per shape, the largest difference of each A/B form from the production form."""
import sys

import torch

sys.path.insert(0, ".")
from deep3d_aerial_b200 import sweep, synth  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
for (v, c, d, h, w) in ((5, 8, 8, 45, 67), (3, 16, 11, 33, 50), (5, 8, 3, 70, 6), (5, 32, 16, 32, 48)):
    rig = synth.tiny_rig(num_views=v, width=w * 4, height=h * 4)
    feats = synth.make_features(v, c, h, w, seed=1).to(dev)
    cur = synth.smooth_depth_map(rig, h, w, seed=1)
    hyps = synth.per_pixel_hypotheses(cur, d, (rig.dmax - rig.dmin) / (4 * d)).to(dev).contiguous()
    uni = synth.uniform_hypotheses(rig.dmin, rig.dmax, d).to(dev)
    tex = sweep.to_texels(feats)
    pose = sweep.relative_poses(torch.from_numpy(rig.proj(4)).to(dev))
    wt = torch.rand(v - 1, h, w).to(dev)
    outs = [sweep.cost_volume(tex, pose, hyps, sweep.AGG_WEIGHTED_PRODUCT, weights=wt, variant=k) for k in (0, 10, 12, 13, 14, 11)]
    var = [sweep.cost_volume(tex, pose, hy, sweep.AGG_VARIANCE, variant=k) for hy in (uni, hyps) for k in (0, 1)]
    extra = []
    if c == 32:
        extra = [sweep.cost_volume(tex, pose, uni, sweep.AGG_GROUP_CORR, groups=8), sweep.cost_volume(tex, pose, hyps, sweep.AGG_PAIR_MEAN)]
    # the same views as slots of a texel pool allocated to the byte: the reference in the LAST slot, a source in the first
    slots = [v + 1, 0] + list(range(2, v))
    pool = torch.empty((v + 2, h, w, c), device=dev)
    for i, sl in enumerate(slots):
        pool[sl].copy_(tex[i])
    pooled = [sweep.cost_volume(pool, pose, hyps, sweep.AGG_WEIGHTED_PRODUCT, weights=wt, view_slots=slots),
              sweep.cost_volume(pool, pose, uni, sweep.AGG_VARIANCE, view_slots=slots),
              sweep.cost_volume(pool, pose, hyps, sweep.AGG_VARIANCE, view_slots=slots, variant=1)]
    if c == 32:
        pooled += [sweep.cost_volume(pool, pose, uni, sweep.AGG_GROUP_CORR, groups=8, view_slots=slots),
                   sweep.cost_volume(pool, pose, hyps, sweep.AGG_PAIR_MEAN, view_slots=slots)]
    same = ["%.1e" % float((a - b).abs().max() / b.abs().max()) for a, b in zip(pooled, [outs[0], var[0], var[3]] + extra)]
    torch.cuda.synchronize()
    print("texel pool vs the dense block (0 = the same kernel served both; sweep_lean's shapes fall through to sweep_base):", same)
    print(v, c, d, h, w, "weighted product:", ["%.1e" % float((o - outs[0]).abs().max()) for o in outs],
          "variance (uniform, per pixel) vs baseline kernel:", "%.1e %.1e" % (float((var[0] - var[1]).abs().max()),
                                                                            float((var[2] - var[3]).abs().max())),
          "finite:", all(bool(torch.isfinite(x).all()) for x in extra))
