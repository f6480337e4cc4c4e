"""Diagnostic (GPU box): which fp32 operation order does CUDA-ATen use along module.py:528-546?

Emulates candidate orders with exact fp64 products (fma(a,b,c) == f32(f64(a)*f64(b)+f64(c)) up to rare
double-rounding ties) and reports the fraction of elements that differ from what torch computes on the
device.  Drives the choices in csrc/sweep*.cuh::project().  Not part of the product or the tests.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep3d_aerial_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)
dev = "cuda"


def f32(x):
    return x.to(torch.float32)


def fma(a, b, c):
    return f32(a.double() * b.double() + c.double())


def frac(a, b, name):
    bad = (a != b)
    print("  %-46s mismatch %.6f   max|diff| %.3e" % (name, bad.float().mean().item(), (a.double() - b.double()).abs().max().item()))


def main():
    rig = synth.make_rig(num_views=5)
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 4          # 4, 2, 1 = cascade stage 1, 2, 3
    h, w = 2752 // scale, 1856 // scale
    print("stage scale %d: %d x %d" % (scale, h, w))
    proj = torch.from_numpy(rig.proj(scale)).to(dev)
    ref = proj[0:1]
    print("relative pose: batched matmul vs the reference's per-view [1,4,4] matmul")
    inv = torch.inverse(ref)
    batched = torch.matmul(proj[1:], inv)
    single = torch.cat([torch.matmul(proj[i:i + 1], torch.inverse(ref)) for i in range(1, 5)], 0)
    frac(batched, single, "matmul(proj[1:], inv) vs per view")
    pose64 = f32(proj[1:].double() @ torch.inverse(ref.double()))
    frac(single, pose64, "per view vs fp64 pose")

    ys, xs = torch.meshgrid(torch.arange(0, h, dtype=torch.float32, device=dev),
                            torch.arange(0, w, dtype=torch.float32, device=dev), indexing="ij")
    x, y = xs.reshape(-1), ys.reshape(-1)
    one = torch.ones_like(x)
    xyz = torch.stack((x, y, one)).unsqueeze(0)
    for v in range(4):
        rot = single[v:v + 1, :3, :3]
        trans = single[v:v + 1, :3, 3:4]
        ray = torch.matmul(rot, xyz)[0]                 # [3,HW]  (module.py:538)
        r = rot[0]
        print("view %d: rot @ [x,y,1]" % v)
        for name, fn in (
            ("A fma(r2,1,fma(r1,y,r0*x))", lambda i: fma(r[i, 2], one, fma(r[i, 1], y, f32(r[i, 0] * x)))),
            ("B fma(r0,x,fma(r1,y,r2))", lambda i: fma(r[i, 0], x, fma(r[i, 1], y, r[i, 2] * one))),
            ("C (r0*x + r1*y) + r2  no fma", lambda i: (r[i, 0] * x + r[i, 1] * y) + r[i, 2]),
            ("D fma(r1,y,fma(r0,x,r2))", lambda i: fma(r[i, 1], y, fma(r[i, 0], x, r[i, 2] * one))),
            ("E fma(r2,1,r0*x) then fma(r1,y,.)", lambda i: fma(r[i, 1], y, fma(r[i, 2], one, f32(r[i, 0] * x)))),
        ):
            got = torch.stack([fn(i) for i in range(3)])
            frac(got, ray, name)
        if v == 0:
            d = torch.tensor(437.5, device=dev)
            pts = ray * d + trans[0]
            X, Y, Z = pts[0], pts[1], pts[2]
            u = X / Z
            print("division X/Z")
            frac(u, f32(X.double() / Z.double()), "torch X/Z vs correctly rounded")
            frac(u, X * (1.0 / Z), "torch X/Z vs X*(1/Z)")
            print("u / ((W-1)/2) - 1")
            gx = u / ((w - 1) / 2) - 1
            inv_s = torch.tensor(1.0, device=dev) / torch.tensor((w - 1) / 2, device=dev)
            frac(gx, u * inv_s - 1, "vs u*f32(1/s) - 1")
            frac(gx, f32(u.double() / ((w - 1) / 2)) - 1, "vs correctly rounded u/s - 1")
            frac(gx, f32(u.double() * (1.0 / ((w - 1) / 2))) - 1, "vs u*f64(1/s) rounded - 1")
            print("x.div_(5)")
            q = torch.randn(1 << 20, device=dev) * 3
            frac(q.clone().div_(5), q * torch.tensor(0.2, device=dev), "vs q*0.2f")
            frac(q.clone().div_(5), f32(q.double() / 5), "vs correctly rounded q/5")
            print("unnormalise ((g+1)/2)*(W-1) as grid_sample does: compare to explicit ops")
            ix = ((gx + 1) / 2) * (w - 1)
            frac(ix, ((gx + 1) * 0.5) * (w - 1), "vs ((g+1)*0.5)*(W-1)")
    # CPU vs CUDA for the same chain
    print("CPU-ATen vs CUDA-ATen for the whole grid (view 0, one plane)")
    rot, trans = single[0:1, :3, :3], single[0:1, :3, 3:4]
    ray_g = torch.matmul(rot, xyz)
    ray_c = torch.matmul(rot.cpu(), xyz.cpu())
    frac(ray_g.cpu(), ray_c, "rays")
    pg = ray_g * 437.5 + trans
    pc = ray_c * 437.5 + trans.cpu()
    ug, uc = pg[:, 0] / pg[:, 2], pc[:, 0] / pc[:, 2]
    frac(ug.cpu(), uc, "u = X/Z")
    gg, gc = ug / ((w - 1) / 2) - 1, uc / ((w - 1) / 2) - 1
    frac(gg.cpu(), gc, "gx")


if __name__ == "__main__":
    main()
