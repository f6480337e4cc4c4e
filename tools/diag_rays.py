"""Diagnostic (GPU box): which fp32 operation order does torch.matmul(rot[1,3,3], xyz[1,3,N]) use, as a function
of the column index?  (At N = 1376*928 the columns past 2^20 were found to round differently.)
    python tools/diag_rays.py [scale=2]
"""
import itertools
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


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    rig = synth.make_rig(num_views=5)
    h, w = 2752 // scale, 1856 // scale
    n = h * w
    proj = torch.from_numpy(rig.proj(scale)).to(dev)
    ys, xs = torch.meshgrid(torch.arange(0, h, dtype=torch.float32, device=dev),
                            torch.arange(0, w, dtype=torch.float32, device=dev), indexing="ij")
    x, y = xs.reshape(-1), ys.reshape(-1)
    one = torch.ones_like(x)
    xyz = torch.stack((x, y, one)).unsqueeze(0)
    col = torch.arange(n, device=dev)
    for v in range(1, 5):
        p = torch.matmul(proj[v:v + 1], torch.inverse(proj[0:1]))
        rot = p[:, :3, :3]
        ray = torch.matmul(rot, xyz)[0]
        r = rot[0]
        terms = lambda i: [(r[i, 0], x), (r[i, 1], y), (r[i, 2], one)]
        a_model = torch.stack([fma(r[i, 2], one, fma(r[i, 1], y, f32(r[i, 0] * x))) for i in range(3)])
        bad = (a_model != ray).any(0)
        print("view %d: model A mismatch columns %d of %d; first %d last %d (2^20 = %d)" % (
            v, int(bad.sum()), n, int(col[bad].min()) if bad.any() else -1, int(col[bad].max()) if bad.any() else -1, 1 << 20))
        if not bad.any():
            continue
        lo = int(col[bad].min())
        # region boundaries: mismatch density per 64k-column block
        dens = [(b, float(bad[b:b + 65536].float().mean())) for b in range(0, n, 65536)]
        print("   blocks with mismatches:", [(b >> 16, round(d, 4)) for b, d in dens if d > 0])
        tail = slice(lo - (lo % 128), n)
        cands = {}
        for perm in itertools.permutations(range(3)):
            for mode in ("fma_chain", "mul_then_fma", "all_rounded"):
                vals = []
                for i in range(3):
                    t = terms(i)
                    (a0, b0), (a1, b1), (a2, b2) = t[perm[0]], t[perm[1]], t[perm[2]]
                    if mode == "fma_chain":       # fma(a2,b2, fma(a1,b1, a0*b0))
                        vals.append(fma(a2, b2[tail], fma(a1, b1[tail], f32(a0 * b0[tail]))))
                    elif mode == "mul_then_fma":  # (a0*b0 rounded + a1*b1 rounded) then fma(a2,b2,.)
                        vals.append(fma(a2, b2[tail], f32(a0 * b0[tail]) + f32(a1 * b1[tail])))
                    else:                         # every product and sum rounded
                        vals.append((f32(a0 * b0[tail]) + f32(a1 * b1[tail])) + f32(a2 * b2[tail]))
                got = torch.stack(vals)
                cands[(perm, mode)] = float((got != ray[:, tail]).float().mean())
        best = sorted(cands.items(), key=lambda kv: kv[1])[:5]
        print("   tail columns [%d, %d): best candidate orders:" % (tail.start, n))
        for (perm, mode), m in best:
            print("      order %s %-13s mismatch %.6f" % ("".join("xy1"[k] for k in perm), mode, m))


if __name__ == "__main__":
    main()
