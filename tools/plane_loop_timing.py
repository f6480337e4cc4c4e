"""Wall time of the plane-at-a-time part of an AdaMVS view (row f1) WITH a regulariser of the reference's shape
(synth.SliceRegulariser: the layer list of SliceCostRegNetRED, adamvs.py:403-427, seeded weights), at the stage-1 and
stage-2 shapes of the 1856 x 2752 configuration: the eager loop of adamvs.py:492-529, the same loop recorded as a CUDA
graph (depthnets.PLANE_LOOP_GRAPHS), and the stateless convolutions batched over all planes
(depthnets.BATCH_STATELESS_CONVS).  Each figure is the whole ada_infer_forward call: weighted-product sweep + regulariser +
streaming soft-argmax.

    python tools/plane_loop_timing.py
"""
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import depthnets, sweep, synth  # noqa: E402


def main():
    torch.set_grad_enabled(False)
    dev = "cuda"
    v = 5
    rig = synth.make_rig(num_views=v)
    for name, scale, c, d in (("stage 1", 4, 32, 48), ("stage 2", 2, 16, 32)):
        h, w = 2752 // scale, 1856 // scale
        proj = torch.from_numpy(rig.proj(scale)).unsqueeze(0).to(dev)
        feats = [torch.randn(1, c, h, w, device=dev) for _ in range(v)]
        hyps = sweep.depth_samples(sweep.SAMPLES_RANGE, d, (h, w), device=torch.device(dev), dmin=rig.dmin,
                                   dmax=rig.dmax).unsqueeze(0)
        conf = [torch.rand(1, 1, h, w, device=dev) for _ in range(v - 1)]
        net = types.SimpleNamespace(in_up=True, reg=None, reg_fuse=synth.SliceRegulariser(c, up=True).to(dev).eval())
        for label, graphs, batched in (("eager loop", False, False), ("CUDA graph", True, False),
                                       ("batched stateless convs", False, True)):
            depthnets.PLANE_LOOP_GRAPHS, depthnets.BATCH_STATELESS_CONVS = graphs, batched
            for _ in range(2):
                depthnets.ada_infer_forward(net, feats, proj, hyps, d, confidence_map=conf)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 5
            for _ in range(n):
                out = depthnets.ada_infer_forward(net, feats, proj, hyps, d, confidence_map=conf)
            torch.cuda.synchronize()
            print("%s (C=%d, D=%d, %dx%d)  %-24s %.2f ms per view" % (name, c, d, w, h, label, (time.perf_counter() - t0) / n * 1e3))
        depthnets.PLANE_LOOP_GRAPHS = depthnets.BATCH_STATELESS_CONVS = False
        float(out["depth"][0, 0, 0])
        del feats, hyps, conf, net
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
