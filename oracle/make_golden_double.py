"""ORACLE — TEST INFRASTRUCTURE ONLY.

Golden vectors of `homo_warping_double` (mvs/mvs_cas/models/module.py:560-601) from the LIVE reference on CPU: the warp
whose coordinate arithmetic runs in fp64 (fp64 projection matrices in, fp32 features and depths).

    python -m oracle.make_golden_double      # -> tests/golden/warp_double_{uniform,perpixel}.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from deep3d_aerial_b200 import synth  # noqa: E402
from oracle import ref_live  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref = ref_live.load()
    torch.set_grad_enabled(False)
    v, c, d, h, w = 3, 4, 5, 20, 24
    rig = synth.tiny_rig(num_views=v, width=w * 4, height=h * 4)
    proj = torch.from_numpy(rig.proj(4)).unsqueeze(0).double()          # [1,V,4,4] fp64
    proj[:, 2, 0, 3] += 4.0 * rig.z_mean                                # view 2: part of the samples leave the image
    feats = synth.make_features(v, c, h, w, seed=21)
    for name, perpixel in (("warp_double_uniform", False), ("warp_double_perpixel", True)):
        if perpixel:
            hyps = synth.per_pixel_hypotheses(synth.smooth_depth_map(rig, h, w, seed=3), d, 0.35).unsqueeze(0)
        else:
            hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d).unsqueeze(0)
        outs = [ref.module.homo_warping_double(feats[i:i + 1], proj[:, i], proj[:, 0], hyps) for i in range(1, v)]
        assert outs[0].dtype == torch.float32
        np.savez(os.path.join(OUT, name + ".npz"), feats=feats.numpy(), proj=proj.numpy(), hyps=hyps.numpy(),
                 warped=torch.stack(outs, 1).numpy())
        print(name, tuple(outs[0].shape))


if __name__ == "__main__":
    main()
