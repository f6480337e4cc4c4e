"""Diagnostic (GPU box): per-view warp parity against CUDA-ATen at the cascade stages' production sizes.
    python tools/diag_stage_parity.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import sweep, synth  # noqa: E402
from oracle import sweep_torch  # noqa: E402  (diagnostic tool, not the product path)

torch.set_grad_enabled(False)
dev = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


rig = synth.make_rig(num_views=5)
for scale, c, d in ((2, 16, 32), (1, 8, 8), (4, 32, 48)):
    h, w = 2752 // scale, 1856 // scale
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(5, c, h, w, generator=g).to(dev)
    proj = torch.from_numpy(rig.proj(scale)).unsqueeze(0).to(dev)
    hy = synth.uniform_hypotheses(rig.dmin, rig.dmax, d, device=dev)
    sub = hy[3:5].unsqueeze(0)
    for v in range(1, 5):
        wv = sweep_torch.warp_source(feats[v:v + 1], proj[:, v], proj[:, 0], sub)[0]
        t2 = sweep.to_texels(torch.stack([feats[0], feats[v]], 0))
        p2 = sweep.relative_poses(torch.stack([proj[0, 0], proj[0, v]], 0))
        gw = sweep.cost_volume(t2, p2, hy, sweep.AGG_WARP, d_begin=3, d_count=2)
        bad = (gw - wv).abs() > 1e-5
        print("scale %d (%dx%d) warp view %d: rel %.3e  max|diff| %.3e  elements off by > 1e-5: %.5f" % (
            scale, h, w, v, rel(gw, wv), float((gw - wv).abs().max()), float(bad.float().mean())))
        if bad.any():
            idx = bad.nonzero()
            ys, xs = idx[:, 2].float(), idx[:, 3].float()
            print("      rows %.0f..%.0f (mean %.0f)  cols %.0f..%.0f (mean %.0f)" % (ys.min(), ys.max(), ys.mean(), xs.min(), xs.max(), xs.mean()))
        del wv, gw
