"""Diagnostic (GPU box): are the relative poses bit-identical however the reference forms them?
    python tools/diag_pose.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import sweep, synth
torch.set_grad_enabled(False)
dev = "cuda"
rig = synth.make_rig(num_views=5)
for scale in (2, 4):
    proj = torch.from_numpy(rig.proj(scale)).unsqueeze(0).to(dev)
    ours = sweep.relative_poses(proj[0])
    projs = torch.unbind(proj, 1)
    for i in range(1, 5):
        ref = torch.matmul(projs[i], torch.inverse(projs[0]))[0]
        ref2 = torch.matmul(proj[0, i:i + 1], torch.inverse(proj[0, 0:1]))[0]
        print("scale %d view %d: ours vs unbind-matmul mismatch %d  max|diff| %.3e   slice-matmul vs unbind-matmul %d" % (
            scale, i, int((ours[i - 1] != ref).sum()), float((ours[i - 1] - ref).abs().max()), int((ref2 != ref).sum())))
    inv_a = torch.inverse(projs[0]); inv_b = torch.linalg.inv_ex(proj[0, 0:1], check_errors=False).inverse
    print("  inverse vs inv_ex mismatch", int((inv_a != inv_b).sum()))
