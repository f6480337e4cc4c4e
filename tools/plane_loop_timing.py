"""Wall time of the plane-at-a-time loop, eager vs CUDA graph (row f1): AdaMVS stage-1 shape, a recurrent
stand-in regulariser with ~12 launches per plane (the reference's SliceCostRegNetRED has more).

    python tools/plane_loop_timing.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import depthnets, sweep, synth  # noqa: E402


class GruLike(torch.nn.Module):
    def forward(self, x, s1, s2):
        s1 = torch.tanh(0.7 * s1 + 0.3 * x[:, :8])
        s2 = 0.5 * s2 + 0.5 * torch.nn.functional.avg_pool2d(torch.cat([s1, x[:, 8:16]], 1), 2)
        logit = 1.5 * s1.mean(1, keepdim=True) + torch.nn.functional.interpolate(s2.mean(1, keepdim=True), scale_factor=2)
        return torch.nn.functional.interpolate(logit, scale_factor=2, mode="nearest"), s1, s2


def main():
    torch.set_grad_enabled(False)
    dev = "cuda"
    v, c, d, h, w = 5, 32, 48, 688, 464
    rig = synth.make_rig(num_views=v)
    proj = torch.from_numpy(rig.proj(4)).unsqueeze(0).to(dev)
    feats = [torch.randn(1, c, h, w, device=dev) for _ in range(v)]
    hyps = sweep.depth_samples(sweep.SAMPLES_RANGE, d, (h, w), device=torch.device(dev), dmin=rig.dmin,
                               dmax=rig.dmax).unsqueeze(0)
    conf = [torch.rand(1, 1, h, w, device=dev) for _ in range(v - 1)]

    class Net:
        pass

    net = Net()
    net.in_up, net.reg, net.reg_fuse = True, None, GruLike()
    for graphs in (False, True):
        depthnets.PLANE_LOOP_GRAPHS = graphs
        for _ in range(2):
            depthnets.ada_infer_forward(net, feats, proj, hyps, d, confidence_map=conf)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            out = depthnets.ada_infer_forward(net, feats, proj, hyps, d, confidence_map=conf)
        torch.cuda.synchronize()
        print("%-6s %.2f ms per view (stage 1, D=%d planes)" % ("graph" if graphs else "eager", (time.perf_counter() - t0) / n * 1e3, d))
    depthnets.PLANE_LOOP_GRAPHS = False
    float(out["depth"][0, 0, 0])


if __name__ == "__main__":
    main()
