"""Like-for-like GPU baseline: the reference's own ATen path -- homo_warping / grid_sample per
source view, elementwise variance, softmax + regression -- on the SAME B200, at the cfg2 shape, chunked over
depth so the 47 GB of temporaries fit.  Not a test (pytest does not collect it) and not the bench's reference arm
(that one is the CPU path, by contract); it lives under tests/ because it executes the oracle.

    python tests/aten_gpu_baseline.py [plane_chunk=16]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import synth  # noqa: E402
from oracle import sweep_torch  # noqa: E402


def main():
    torch.set_grad_enabled(False)
    dev = "cuda"
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    v, c, d, h, w = 5, 32, 384, 688, 464
    rig = synth.make_rig(num_views=v)
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(v, c, h, w, generator=g).to(dev)
    views = [feats[i:i + 1] for i in range(v)]
    proj = torch.from_numpy(rig.proj(4)).unsqueeze(0).to(dev)
    hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d, device=dev).unsqueeze(0)
    logits = 4.0 * torch.randn(1, d, h, w, generator=g).to(dev)
    volume = torch.empty((1, c, d, h, w), device=dev)

    def step():
        for d0 in range(0, d, chunk):
            volume[:, :, d0:d0 + chunk] = sweep_torch.variance_volume(views, proj, hyps[:, d0:d0 + chunk])
        return sweep_torch.regress_maxprob(logits, hyps)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    vox = d * h * w
    print("reference ATen path on this GPU (cfg2, planes in chunks of %d): %.1f ms per view = %.2f Gvoxel/s, peak memory %.1f GB"
          % (chunk, ms, vox / ms / 1e6, torch.cuda.max_memory_allocated() / 1e9))
    float(out[0][0, 0, 0])


if __name__ == "__main__":
    main()
