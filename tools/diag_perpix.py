"""Diagnostic (GPU box): where do the kernel's sample coordinates differ from CUDA-ATen's at a cascade stage
(per-pixel hypotheses)?  Source "features" are coordinate ramps (channel 0 = x, 1 = y), so the warped volume IS
the sampling position; it is compared with the grid the reference's chain (module.py:528-546) produces.

    python tools/diag_perpix.py [scale=2]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import sweep, synth  # noqa: E402

torch.set_grad_enabled(False)
dev = "cuda"


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ratio, d = {4: (4, 48), 2: (2, 32), 1: (1, 8)}[scale]
    rig = synth.make_rig(num_views=5)
    h, w = 2752 // scale, 1856 // scale
    proj = torch.from_numpy(rig.proj(scale)).to(dev)
    cur = synth.smooth_depth_map(rig, h, w, seed=scale).to(dev)
    hyps = sweep.depth_samples(sweep.SAMPLES_AROUND, d, (h, w), cur=cur, interval=ratio * (rig.dmax - rig.dmin) / 384)
    ys, xs = torch.meshgrid(torch.arange(0, h, dtype=torch.float32, device=dev),
                            torch.arange(0, w, dtype=torch.float32, device=dev), indexing="ij")
    ramp = torch.stack([xs, ys, torch.zeros_like(xs), torch.zeros_like(xs)], 0)       # [4,H,W]
    for v in range(1, 5):
        # the reference chain, verbatim (module.py:528-546), planes 0..3
        sub = hyps[:4]
        p = torch.matmul(proj[v:v + 1], torch.inverse(proj[0:1]))
        rot, trans = p[:, :3, :3], p[:, :3, 3:4]
        xyz = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dev))).unsqueeze(0)
        rot_xyz = torch.matmul(rot, xyz)
        rdx = rot_xyz.unsqueeze(2).repeat(1, 1, 4, 1) * sub.view(1, 1, 4, -1)
        pxyz = rdx + trans.view(1, 3, 1, 1)
        pxy = pxyz[:, :2] / pxyz[:, 2:3]
        gx = pxy[:, 0] / ((w - 1) / 2) - 1
        gy = pxy[:, 1] / ((h - 1) / 2) - 1
        ix = ((gx + 1) / 2) * (w - 1)
        iy = ((gy + 1) / 2) * (h - 1)
        tex = sweep.to_texels(torch.stack([ramp, ramp], 0))
        pose = sweep.relative_poses(torch.stack([proj[0], proj[v]], 0))
        for variant in (0, 1):
            got = sweep.cost_volume(tex, pose, sub.contiguous(), sweep.AGG_WARP, variant=variant)     # [4,4,H,W]
            inside = (ix > 1) & (ix < w - 2) & (iy > 1) & (iy < h - 2)
            dx = (got[0].reshape(1, 4, -1) - ix).abs()[inside]
            dy = (got[1].reshape(1, 4, -1) - iy).abs()[inside]
            print("view %d variant %d: |dx| max %.3e  frac>1e-5 %.5f   |dy| max %.3e  frac>1e-5 %.5f   (ulp at %d: %.1e)"
                  % (v, variant, dx.max().item(), (dx > 1e-5).float().mean().item(), dy.max().item(),
                     (dy > 1e-5).float().mean().item(), h, 2.0 ** -23 * 1024))


if __name__ == "__main__":
    main()
