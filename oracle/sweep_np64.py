"""ORACLE — TEST INFRASTRUCTURE ONLY (arbiter).

An independent float64 numpy restatement of warp + bilinear sample + aggregation, written from
the geometry rather than from ATen calls, used to arbitrate when the fp32 ATen path and the CUDA
kernels disagree at the 1e-4 level.  Follows
`mvs/mvs_cas/models/module.py:528-555` for the projection and
`torch/include/ATen/native/GridSampler.h:27-32` (+ `cuda/GridSampler.cuh`) for the
un-normalisation, corner weights and zero padding.  Pinned by tests/test_oracle.py against the
torch restatement (which is pinned against the live reference).  Small inputs only.
"""
from __future__ import annotations

import numpy as np


def relative_pose(src_proj, ref_proj):
    rel = src_proj.astype(np.float64) @ np.linalg.inv(ref_proj.astype(np.float64))
    return rel[:3, :3], rel[:3, 3]


def pixel_coords(rot, trans, hyps, h, w):
    """Source-image pixel coordinates (ix, iy), each [D,h,w], for hypotheses [D] or [D,h,w]."""
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    ray = np.einsum("ij,jhw->ihw", rot, np.stack([xs, ys, np.ones_like(xs)]))
    d = hyps.astype(np.float64)
    d = d.reshape(-1, 1, 1) if d.ndim == 1 else d
    pts = ray[:, None] * d[None] + trans.reshape(3, 1, 1, 1)
    u = pts[0] / pts[2]
    v = pts[1] / pts[2]
    # normalise to [-1,1] (module.py:543-544) and back (GridSampler.h:31): identity in exact math
    gx = u / ((w - 1) / 2) - 1
    gy = v / ((h - 1) / 2) - 1
    return ((gx + 1) / 2) * (w - 1), ((gy + 1) / 2) * (h - 1)


def bilinear_zero_pad(fea, ix, iy):
    """fea [C,h,w]; ix, iy [D,h,w] -> [C,D,h,w]; corners outside the image contribute zero."""
    c, h, w = fea.shape
    fea = fea.astype(np.float64)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    out = np.zeros((c,) + ix.shape, dtype=np.float64)
    bad = ~(np.isfinite(ix) & np.isfinite(iy))
    for dy in (0, 1):
        for dx in (0, 1):
            xc = x0 + dx
            yc = y0 + dy
            wgt = (1 - np.abs(ix - xc)) * (1 - np.abs(iy - yc))
            ok = (xc >= 0) & (xc <= w - 1) & (yc >= 0) & (yc <= h - 1) & ~bad
            xi = np.where(ok, xc, 0).astype(np.int64)
            yi = np.where(ok, yc, 0).astype(np.int64)
            out += fea[:, yi, xi] * np.where(ok, wgt, 0.0)[None]
    return out


def warp_source(src_fea, src_proj, ref_proj, hyps):
    rot, trans = relative_pose(src_proj, ref_proj)
    _, h, w = src_fea.shape
    ix, iy = pixel_coords(rot, trans, hyps, h, w)
    return bilinear_zero_pad(src_fea, ix, iy)


def variance_volume(features, projs, hyps):
    """features [V,C,h,w], projs [V,4,4], hyps [D] or [D,h,w] -> [C,D,h,w] (cas_mvsnet.py:46-60)."""
    v = features.shape[0]
    d = hyps.shape[0]
    ref = np.repeat(features[0].astype(np.float64)[:, None], d, axis=1)
    s, q = ref.copy(), ref ** 2
    for i in range(1, v):
        wv = warp_source(features[i], projs[i], projs[0], hyps)
        s += wv
        q += wv ** 2
    return q / v - (s / v) ** 2


def pair_mean_volumes(features, projs, hyps):
    """[V-1,D,h,w]: mean over channels of ref * warped_i (adamvs.py:466-475)."""
    ref = features[0].astype(np.float64)[:, None]
    return np.stack([(ref * warp_source(features[i], projs[i], projs[0], hyps)).mean(0)
                     for i in range(1, features.shape[0])])


def groupwise_volume(features, projs, hyps, groups):
    c = features.shape[1]
    ref = features[0].astype(np.float64)[:, None]
    acc = 0
    for i in range(1, features.shape[0]):
        p = ref * warp_source(features[i], projs[i], projs[0], hyps)
        acc = acc + p.reshape(groups, c // groups, *p.shape[1:]).mean(1)
    return acc / (features.shape[0] - 1)


def weighted_product_volume(features, projs, hyps, weights, eps_in_numerator=False):
    """weights [V-1,h,w] already at feature resolution (adamvs.py:492-509 / 287-301)."""
    ref = features[0].astype(np.float64)[:, None]
    num = 0
    den = 0
    for i in range(1, features.shape[0]):
        wt = weights[i - 1].astype(np.float64)[None, None]
        num = num + ref * warp_source(features[i], projs[i], projs[0], hyps) * wt
        den = den + wt
    return (num + 1e-5) / den if eps_in_numerator else num / (den + 1e-5)


def softmax_regress(logits, hyps):
    """logits [D,h,w]; returns prob, depth, idx_expect(float)."""
    x = logits.astype(np.float64)
    e = np.exp(x - x.max(0, keepdims=True))
    p = e / e.sum(0, keepdims=True)
    d = hyps.astype(np.float64)
    d = d.reshape(-1, 1, 1) if d.ndim == 1 else d
    k = np.arange(x.shape[0], dtype=np.float64).reshape(-1, 1, 1)
    return p, (p * d).sum(0), (p * k).sum(0)
