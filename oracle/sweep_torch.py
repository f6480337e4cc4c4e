"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.

A CPU (or any-device) restatement, in plain PyTorch ops, of the reference's
plane-sweep hot path.  The arithmetic of that path lives in a third-party
dependency of the reference, PyTorch itself (`F.grid_sample`, `torch.inverse`,
`F.softmax`, `F.interpolate`; the reference pins pytorch 2.3.1 in
`environment.yml:83`, this image has 2.11.0), so the restatement calls the same
ATen ops in the same order and therefore reproduces the reference bit for bit on
the same device.  It is pinned two ways (tests/test_oracle.py):

  * against golden vectors produced by importing the *live* reference
    (`oracle/make_golden.py`, fixtures under `tests/golden/`), and
  * against the live reference itself whenever `/root/reference` is present.

The reference has no tests, golden vectors or fixtures of its own for this path
(SURVEY.md §4, §8c), so beyond those two anchors parity is *unpinned by the
reference's own tests*.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module.

Every function cites the reference lines it follows (paths relative to
`/root/reference/mvs/mvs_cas/models/`).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- a1
def relative_pose(src_proj, ref_proj):
    """`module.py:528-530`: P = P_src @ inverse(P_ref); returns (rot[B,3,3], trans[B,3,1])."""
    rel = torch.matmul(src_proj, torch.inverse(ref_proj))
    return rel[:, :3, :3], rel[:, :3, 3:4]


def sampling_grid(rot, trans, depth_values, height, width):
    """`module.py:532-546`: normalised sampling grid [B, D, H*W, 2] for one source view.

    Order of fp32 roundings kept: rot @ [x,y,1] -> * depth -> + trans -> x/z, y/z ->
    / ((size-1)/2) -> - 1.
    """
    batch = rot.shape[0]
    num_depth = depth_values.shape[1]
    dev = rot.device
    ys, xs = torch.meshgrid(torch.arange(0, height, dtype=torch.float32, device=dev),
                            torch.arange(0, width, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, device=dev)))
    pix = pix.unsqueeze(0).repeat(batch, 1, 1)                      # [B,3,HW]
    ray = torch.matmul(rot, pix)                                    # [B,3,HW]
    pts = ray.unsqueeze(2).repeat(1, 1, num_depth, 1) * depth_values.view(batch, 1, num_depth, -1)
    pts = pts + trans.view(batch, 3, 1, 1)                          # [B,3,D,HW]
    uv = pts[:, :2] / pts[:, 2:3]
    gx = uv[:, 0] / ((width - 1) / 2) - 1
    gy = uv[:, 1] / ((height - 1) / 2) - 1
    return torch.stack((gx, gy), dim=3)


def warp_source(src_fea, src_proj, ref_proj, depth_values):
    """`homo_warping_float`, `module.py:516-557`: [B,C,H,W] -> [B,C,D,H,W].

    `depth_values` is [B,D] or [B,D,H,W].  Bilinear, zero padding, align_corners=True
    (the branch torch>=1.3 selects at `module.py:548-553`).
    """
    batch, channels, height, width = src_fea.shape
    num_depth = depth_values.shape[1]
    with torch.no_grad():
        rot, trans = relative_pose(src_proj, ref_proj)
        grid = sampling_grid(rot, trans, depth_values, height, width)
    out = F.grid_sample(src_fea, grid.view(batch, num_depth * height, width, 2), mode="bilinear",
                        padding_mode="zeros", align_corners=True)
    return out.view(batch, channels, num_depth, height, width)


def warp_source_double(src_fea, src_proj, ref_proj, depth_values):
    """`module.py:560-601` (homo_warping_double): the same warp with the coordinate arithmetic in fp64 -- fp64
    projection matrices (upstream multiplies `rot` by an fp64 pixel grid; torch.matmul does not promote), fp32 depths
    widened by `.double()`, the normalised grid cast back to fp32 for the fp32 `grid_sample`."""
    batch, channels, height, width = src_fea.shape
    num_depth = depth_values.shape[1]
    rot, trans = relative_pose(src_proj, ref_proj)                   # fp64
    dev = src_fea.device
    ys, xs = torch.meshgrid(torch.arange(0, height, dtype=torch.float32, device=dev),
                            torch.arange(0, width, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, device=dev)))
    pix = pix.unsqueeze(0).repeat(batch, 1, 1).double()
    ray = torch.matmul(rot, pix)
    pts = ray.unsqueeze(2).repeat(1, 1, num_depth, 1) * depth_values.view(batch, 1, num_depth, -1).double()
    pts = pts + trans.view(batch, 3, 1, 1)
    uv = pts[:, :2] / pts[:, 2:3]
    grid = torch.stack((uv[:, 0] / ((width - 1) / 2) - 1, uv[:, 1] / ((height - 1) / 2) - 1), dim=3).float()
    warped = F.grid_sample(src_fea, grid.view(batch, num_depth * height, width, 2), mode="bilinear",
                           padding_mode="zeros", align_corners=True)
    return warped.view(batch, channels, num_depth, height, width)


# --------------------------------------------------------------------------- a3
def variance_volume(features, proj_matrices, depth_values):
    """`cas_mvsnet.py:46-60` (same lines in `msrednet.py:217-230`, `ucsnet.py:119-134`).

    features: list of V tensors [B,C,H,W] (index 0 = reference view);
    proj_matrices: [B,V,4,4]; returns sq/V - (sum/V)^2 as [B,C,D,H,W].
    """
    projs = torch.unbind(proj_matrices, 1)
    num_views = len(features)
    num_depth = depth_values.shape[1]
    ref = features[0].unsqueeze(2).repeat(1, 1, num_depth, 1, 1)
    acc = ref
    acc_sq = ref ** 2
    for fea, proj in zip(features[1:], projs[1:]):
        w = warp_source(fea, proj, projs[0], depth_values)
        acc = acc + w
        acc_sq = acc_sq + w.pow(2)
    return acc_sq.div_(num_views).sub_(acc.div_(num_views).pow_(2))


# --------------------------------------------------------------------------- a5
def pair_mean_volumes(features, proj_matrices, depth_values):
    """`adamvs.py:466-475`: per source view, mean over channels of ref*warped -> [V-1] x [B,D,H,W]."""
    projs = torch.unbind(proj_matrices, 1)
    out = []
    for fea, proj in zip(features[1:], projs[1:]):
        planes = []
        for d in range(depth_values.shape[1]):
            w = warp_source(fea, proj, projs[0], depth_values[:, d:d + 1])
            planes.append((features[0].unsqueeze(2) * w).mean(dim=1).squeeze(1))
        out.append(torch.stack(planes, dim=1))
    return out


# --------------------------------------------------------------------------- a6
def groupwise_correlation_volume(features, proj_matrices, depth_values, groups):
    """Group-wise correlation averaged over source views.

    The reference only has a commented-out call `groupwise_correlation(ref, warped, 8, 1)`
    (`adamvs.py:271, 295`) whose body is not in the repository; by analogy with the pair
    volume (`adamvs.py:473`, which is the G=1 case) each source view contributes
    (ref*warped).view(B,G,C/G,D,H,W).mean(2); views are averaged.  PARITY UNPINNED BY THE
    REFERENCE for this function.
    """
    projs = torch.unbind(proj_matrices, 1)
    b, c, h, w = features[0].shape
    d = depth_values.shape[1]
    acc = 0
    for fea, proj in zip(features[1:], projs[1:]):
        wv = warp_source(fea, proj, projs[0], depth_values)
        acc = acc + (features[0].unsqueeze(2) * wv).view(b, groups, c // groups, d, h, w).mean(2)
    return acc / (len(features) - 1)


# --------------------------------------------------------------------------- a4
def resize_weight(weight, h, w):
    """`adamvs.py:502`: bilinear resize of a [B,1,h',w'] view-weight map, align_corners=False."""
    return F.interpolate(weight, [h, w], mode="bilinear", align_corners=False)


def weighted_product_volume(features, proj_matrices, depth_values, view_weights, eps_in_numerator=False):
    """AdaMVS visibility-weighted product volume -> [B,C,D,H,W].

    Inference form (`adamvs.py:492-509`): sum_i (warp_i*ref)*w_i / (1e-5 + sum_i w_i).
    Training form (`adamvs.py:262, 287-301`, eps_in_numerator=True):
    (1e-5 + sum_i prod_i*w_i) / sum_i w_i.
    view_weights: list of V-1 maps [B,1,h',w'], resized to the feature resolution first.
    """
    projs = torch.unbind(proj_matrices, 1)
    _, _, h, w = features[0].shape
    planes = []
    for d in range(depth_values.shape[1]):
        num = 1e-5 if eps_in_numerator else 0
        den = 0 if eps_in_numerator else 1e-5
        for i, (fea, proj) in enumerate(zip(features[1:], projs[1:])):
            wv = warp_source(fea, proj, projs[0], depth_values[:, d:d + 1])
            wt = resize_weight(view_weights[i], h, w).unsqueeze(1)
            num = num + (wv * features[0].unsqueeze(2)) * wt
            den = den + wt
        planes.append(num / den)
    return torch.cat(planes, dim=2)


# --------------------------------------------------------------------------- a7
def expectation(prob, depth_values):
    """`depth_regression`, `module.py:605-613`: sum_d p*d; 4-D hypotheses are bilinearly resized
    to p's H x W first (align_corners=False)."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1)
    else:
        depth_values = F.interpolate(depth_values, [prob.shape[2], prob.shape[3]], mode="bilinear",
                                     align_corners=False)
    return torch.sum(prob * depth_values, 1)


# --------------------------------------------------------------------------- a8
def regress_window4(logits, depth_values):
    """`cas_mvsnet.py:69-76` (= `ucsnet.py:137-146`): softmax over D, expected depth and the
    4-plane window confidence p[i-1]+p[i]+p[i+1]+p[i+2] at i = clamp(floor(sum p*k))."""
    num_depth = logits.shape[1]
    prob = F.softmax(logits, dim=1)
    depth = expectation(prob, depth_values)
    padded = F.pad(prob.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2))
    win = 4 * F.avg_pool3d(padded, (4, 1, 1), stride=1, padding=0).squeeze(1)
    idx = expectation(prob, torch.arange(num_depth, device=prob.device, dtype=torch.float)).long()
    idx = idx.clamp(min=0, max=num_depth - 1)
    conf = torch.gather(win, 1, idx.unsqueeze(1)).squeeze(1)
    return depth, conf, idx


# --------------------------------------------------------------------------- a9
def regress_maxprob(logits, depth_values):
    """`msrednet.py:234-238`, `adamvs.py:306-310, 478-483`: softmax, expected depth, max prob, argmax."""
    prob = F.softmax(logits, dim=1)
    depth = expectation(prob, depth_values)
    conf, idx = prob.max(1)
    return depth, conf, idx


# -------------------------------------------------------------------------- a10
def regress_streaming(logit_slices, depth_slices, upsample2=False):
    """`adamvs.py:514-529`, `msrednet.py:418-437`: un-normalised exp accumulated plane by plane.

    logit_slices: iterable of [B,1,H',W']; depth_slices: iterable of [B,1,h,w].
    upsample2 (`adamvs.py:519-520`): the depth slice is bilinearly doubled first.
    """
    s = m = acc = None
    for x, d in zip(logit_slices, depth_slices):
        e = x.exp()
        if s is None:
            s, m, acc = torch.zeros_like(e), torch.zeros_like(e), torch.zeros_like(e)
        flag = (m < e).float()
        m = flag * e + (1 - flag) * m
        if upsample2:
            d = F.interpolate(d, [e.shape[2], e.shape[3]], mode="bilinear", align_corners=False)
        acc = d * e + acc
        s = s + e
    tot = s + 1e-10
    return (acc / tot).squeeze(1), (m / tot).squeeze(1)


# -------------------------------------------------------------------------- a11
def depth_range_samples(cur_depth, ndepth, interval, shape):
    """`get_depth_range_samples` / `get_cur_depth_range_samples`, `module.py:616-650` -> [B,D,H,W].

    cur_depth [B,2+] (a range: first/last entries are dmin/dmax) or [B,H,W] (previous estimate).
    """
    if cur_depth.dim() == 2:
        lo = cur_depth[:, 0]
        hi = cur_depth[:, -1]
        step = (hi - lo) / (ndepth - 1)
        k = torch.arange(0, ndepth, device=cur_depth.device, dtype=cur_depth.dtype).reshape(1, -1)
        s = lo.unsqueeze(1) + k * step.unsqueeze(1)
        return s.unsqueeze(-1).unsqueeze(-1).repeat(1, 1, shape[1], shape[2])
    lo = cur_depth - ndepth / 2 * interval
    hi = cur_depth + ndepth / 2 * interval
    step = (hi - lo) / (ndepth - 1)
    k = torch.arange(0, ndepth, device=cur_depth.device, dtype=cur_depth.dtype).reshape(1, -1, 1, 1)
    return lo.unsqueeze(1) + k * step.unsqueeze(1)


def cascade_stage_hypotheses(cur_depth, ndepth, interval, full_hw, stage_scale):
    """Cas-MVSNet / RED-Net stage glue, `cas_mvsnet.py:206-226` (= `msrednet.py:495-515`):
    previous depth [B,h',w'] (or the [B,2] range) -> bilinear to full res -> samples at full
    res -> trilinear down to the stage's [D, H/s, W/s] (all align_corners=False)."""
    fh, fw = full_hw
    if cur_depth.dim() == 3:
        cur_depth = F.interpolate(cur_depth.unsqueeze(1), [fh, fw], mode="bilinear",
                                  align_corners=False).squeeze(1)
    full = depth_range_samples(cur_depth, ndepth, interval, [cur_depth.shape[0], fh, fw])
    dv = F.interpolate(full.unsqueeze(1), [ndepth, fh // int(stage_scale), fw // int(stage_scale)],
                       mode="trilinear", align_corners=False)
    return dv.squeeze(1)


# -------------------------------------------------------------------------- a12
def exp_variance(prob, depth_values, depth, lamb):
    """`ucsnet.py:148-149`: lamb * sqrt(sum_d p * (d - depth)^2)."""
    return lamb * torch.sum((depth_values - depth.unsqueeze(1)) ** 2 * prob, dim=1) ** 0.5


def uncertainty_samples(cur_depth, spread, ndepth):
    """`ucsnet.py:30-53`, stage>=2 branch: ndepth samples over [depth-spread, depth+spread] (+1e-12).
    cur_depth, spread: [B,1,H,W] -> [B,D,H,W]."""
    lo = cur_depth - spread
    hi = cur_depth + spread
    step = (hi - lo) / (float(ndepth) - 1)
    return torch.cat([lo + step * i + 1e-12 for i in range(int(ndepth))], 1)


# ------------------------------------------------------------- timed CPU baseline
def cpu_step_variance(features, proj_matrices, depth_values, logits, plane_chunk=None):
    """One pass of the hot path the way the reference runs it on CPU: variance volume
    (depth-sliced like `msrednet.py:400-414` when plane_chunk is given, so temporaries fit), then
    softmax + regression + max-prob confidence on the given logit volume [B,D,H,W] (the CNN
    regulariser between the two is out of scope on both arms, so the logits are an input).
    Returns (depth, conf, checksum of the variance volume)."""
    d_total = depth_values.shape[1]
    step = plane_chunk or d_total
    check = 0.0
    for d0 in range(0, d_total, step):
        var = variance_volume(features, proj_matrices, depth_values[:, d0:d0 + step])
        check += float(var[..., ::7, ::5].sum())
    depth, conf, _ = regress_maxprob(logits, depth_values)
    return depth, conf, check
