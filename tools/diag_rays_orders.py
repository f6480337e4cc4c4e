"""Diagnostic (GPU box): which accumulation order explains torch.matmul([1,3,3] @ [1,3,N]) column by column at the sizes
where it leaves the kernels' order fma(r2,1,fma(r1,y,r0*x))?   python tools/diag_rays_orders.py [N ...]"""
import itertools
import sys

import torch

torch.set_grad_enabled(False)
dev = "cuda"


def f32(x):
    return x.to(torch.float32)


def fma(a, b, c):
    return f32(a.double() * b.double() + c.double())


rots = [torch.tensor([[[0.9961947, 0.01, -7.3], [0.002, 1.0038, 161.5], [1.1e-5, -2.3e-5, 1.01]]], device=dev),
        torch.tensor([[[1.0123, -0.0431, 35.7], [0.0377, 0.9911, -88.2], [-3.1e-5, 1.7e-5, 0.97]]], device=dev)]
w = 928
for n in [int(a) for a in sys.argv[1:]] or [1276928]:
    idx = torch.arange(n, device=dev)
    x, y = (idx % w).float(), (idx // w).float()
    one = torch.ones_like(x)
    ops = {"x": x, "y": y, "1": one}
    for rot in rots:
        ray = torch.matmul(rot, torch.stack((x, y, one)).unsqueeze(0))[0]
        r = rot[0]
        col = {"x": 0, "y": 1, "1": 2}
        cands = {}
        for perm in itertools.permutations("xy1"):
            a, b, c = perm
            # chain: fma(rc, c, fma(rb, b, ra*a))
            cands["fma(%s,fma(%s,%s*))" % (c, b, a)] = torch.stack(
                [fma(r[i, col[c]], ops[c], fma(r[i, col[b]], ops[b], f32(r[i, col[a]] * ops[a]))) for i in range(3)])
            # pairwise: (ra*a) + fma(rb, b, rc*c)  -- two partial sums added (split-K)
            cands["%s* + fma(%s,%s*)" % (a, b, c)] = torch.stack(
                [f32(f32(r[i, col[a]] * ops[a]) + fma(r[i, col[b]], ops[b], f32(r[i, col[c]] * ops[c]))) for i in range(3)])
        base = cands["fma(1,fma(y,x*))"]
        bad = (base != ray).any(0)
        print("N=%d: kernel order wrong in %d columns, first %d, last %d" % (n, int(bad.sum()), int(idx[bad].min()) if bad.any() else -1,
                                                                        int(idx[bad].max()) if bad.any() else -1))
        if not bad.any():
            continue
        lo = int(idx[bad].min())
        tail = idx >= lo - (lo % 32)
        print("   tail = columns >= %d (%d columns)" % (lo - lo % 32, int(tail.sum())))
        for name, c in cands.items():
            wrong_tail = ((c != ray).any(0) & tail)
            wrong_head = ((c != ray).any(0) & ~tail)
            print("   %-22s wrong in tail: %7d   wrong in head: %7d" % (name, int(wrong_tail.sum()), int(wrong_head.sum())))
        # per 32-column block of the tail: is one candidate right for the whole block?
        names = list(cands)
        okb = torch.stack([~(cands[k] != ray).any(0) for k in names]).float()            # [K, n]
        tb = tail.nonzero()[0].item()
        nblk = (n - tb) // 32
        blk = okb[:, tb:tb + nblk * 32].view(len(names), nblk, 32).min(2).values          # block fully explained
        best = blk.argmax(0)
        none = int((blk.max(0).values == 0).sum())
        hist = torch.bincount(best[blk.max(0).values > 0], minlength=len(names))
        print("   32-column blocks of the tail: %d, explained by none: %d; first fully-explaining candidate per block:" % (nblk, none))
        for k, name in enumerate(names):
            if int(hist[k]):
                print("      %-22s %d blocks" % (name, int(hist[k])))
