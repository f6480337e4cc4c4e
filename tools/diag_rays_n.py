"""Diagnostic (GPU box): for which column counts N, and from which column on, does torch.matmul([1,3,3] @ [1,3,N])
leave the "x, y, 1 fma chain" order?   python tools/diag_rays_n.py
"""
import torch

torch.set_grad_enabled(False)
dev = "cuda"


def f32(x):
    return x.to(torch.float32)


def fma(a, b, c):
    return f32(a.double() * b.double() + c.double())


rot = torch.tensor([[[0.9961947, 0.01, -7.3], [0.002, 1.0038, 161.5], [1.1e-5, -2.3e-5, 1.01]]], device=dev)
w = 928
for n in (1 << 20, (1 << 20) + 128, 1200000, 1276928, 1500000, (1 << 21) - 64, (1 << 21) + 4096, 3000000, 5107712, 319232):
    idx = torch.arange(n, device=dev)
    x, y = (idx % w).float(), (idx // w).float()
    one = torch.ones_like(x)
    ray = torch.matmul(rot, torch.stack((x, y, one)).unsqueeze(0))[0]
    r = rot[0]
    a = torch.stack([fma(r[i, 2], one, fma(r[i, 1], y, f32(r[i, 0] * x))) for i in range(3)])
    b = torch.stack([fma(r[i, 2], one, fma(r[i, 0], x, f32(r[i, 1] * y))) for i in range(3)])
    bad_a = (a != ray).any(0)
    bad_b = (b != ray).any(0)
    first_a = int(idx[bad_a].min()) if bad_a.any() else -1
    # where does order "y, x, 1" explain everything that order "x, y, 1" does not?
    both = (bad_a & bad_b)
    print("N=%8d: order xy1 wrong in %7d cols (first %8d); order yx1 wrong in %7d cols (last %8d); neither explains %d" % (
        n, int(bad_a.sum()), first_a, int(bad_b.sum()), int(idx[bad_b].max()) if bad_b.any() else -1, int(both.sum())))
