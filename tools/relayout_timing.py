import torch, sys
sys.path.insert(0, "/root/repo")
from deep3d_aerial_b200 import sweep
for c, h, w in ((32, 688, 464), (16, 1376, 928), (8, 2752, 1856), (32, 43, 29)):
    x = torch.randn(1, c, h, w, device="cuda")
    out = sweep.to_texels(x)
    assert torch.equal(out[0], x[0].permute(1, 2, 0).contiguous()), (c, h, w)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3): sweep.to_texels(x, out=out)
    ev[0].record()
    for _ in range(20): sweep.to_texels(x, out=out)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 20
    print("C=%d %dx%d: %.1f us, %.2f TB/s" % (c, h, w, ms * 1e3, 2 * 4 * c * h * w / ms / 1e9))
