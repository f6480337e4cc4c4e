"""Drop-in replacements for the `DepthNet` / `InferDepthNet` forward passes of the reference's cascade
networks (SURVEY.md §8b level b3) -- same signatures, asserts, return-dict keys and list quirks -- with
the warp + aggregation and the softmax + regression done by libd3dsweep.  The CNN regularisers are
NOT here: they are passed in (Cas-MVSNet, RED-Net, UCS-Net) or looked up on `self` (`self.reg`,
`self.reg_fuse` of the reference's AdaMVS InferDepthNet) and run in PyTorch, unchanged.

    cas_depthnet_forward      cas_mvsnet.py:35-78      DepthNet.forward
    red_depthnet_forward      msrednet.py:206-241      DepthNet.forward (whole-volume form)
    red_infer_forward         msrednet.py:377-438      InferDepthNet.forward (plane-at-a-time GRU)
    ada_infer_forward         adamvs.py:436-531        InferDepthNet.forward
    ada_depthnet_forward      adamvs.py:247-312        DepthNet.forward (whole-volume form)
    ucs_compute_depth         ucsnet.py:99-151         compute_depth
    ucs_uncertainty_samples   ucsnet.py:30-53          uncertainty_aware_samples

`deep3d_aerial_b200.install()` binds these onto the reference's classes; the thin
`nn.Module` wrappers at the bottom serve callers that build the networks themselves.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import sweep

# Opt-in: record the D-plane "regulariser + streaming soft-argmax" loop of the
# plane-at-a-time models in a CUDA graph per (module, shape) and replay it per reference view (graphs.py).
# Off by default: the regulariser must be capture-safe, which holds for the reference's modules but is the
# caller's promise for anything else.  B = 1 per graph replay keeps the reference's semantics.
PLANE_LOOP_GRAPHS = False


def _plane_loop(owner, tag, step, batch, planes, channels, hw, out_hw, state_shapes, hyps_hw, device):
    from .graphs import PlaneLoop

    cache = owner.__dict__.setdefault("_d3d_plane_loops", {})
    key = (tag, id(step), batch, planes, channels, tuple(hw), tuple(out_hw), tuple(hyps_hw), str(device))
    if key not in cache:
        cache[key] = PlaneLoop(step, batch, planes, channels, hw, out_hw, state_shapes, device, hyps_hw=hyps_hw)
    return cache[key]


def _volume_into(out, features, proj_matrices, depth_values, mode, scenes=None, **kw):
    """Plane-major cost volume of every batch item written straight into `out` [B,D,C,h,w]."""
    with torch.no_grad():
        for b in range(features[0].shape[0]):
            texels, pose, rays, slots = scenes[b] if scenes is not None else _scene(features, proj_matrices, b)
            w = kw.get("weights")
            sweep.cost_volume(texels, pose, depth_values[b].contiguous(), mode, plane_major=True, out=out[b], rays=rays,
                              view_slots=slots, **{**kw, "weights": None if w is None else w[b]})


def _inference_only(what, *tensors):
    """The kernels have no backward: a caller that expects gradients must hear about it, not train on constants.
    (install() rebinds the classes the reference's train.py uses too.)"""
    if not torch.is_grad_enabled():
        return
    for t in tensors:
        items = t if isinstance(t, (list, tuple)) else (t,)
        for x in items:
            if isinstance(x, torch.Tensor) and x.requires_grad:
                raise RuntimeError(
                    "deep3d_aerial_b200: %s is inference only (the fused sweep kernels have no autograd backward) but was "
                    "called with grad enabled on an input that requires grad; wrap the call in torch.no_grad(), or do "
                    "not install() in a training process" % what)


def _check(features, proj_matrices, depth_values, num_depth):
    assert len(features) == len(proj_matrices), "Different number of images and projection matrices"
    assert depth_values.shape[1] == num_depth, "depth_values.shape[1]:{}  num_depth:{}".format(
        depth_values.shape[1], num_depth)


# Opt-in (row f4): a texel_pool.TexelPool.  Set where the feature maps of an image are the same tensor objects for every
# reference view that uses the image (predict.py sets it together with FeatureCache): an image is then laid out once and the
# sweeps name their views by pool slot instead of laying out a dense [V,H,W,C] block per reference view and stage.
TEXEL_POOL = None


def _scene(features, proj_matrices, b):
    """What every sweep over one stage of batch item b shares: channels-last texels [V,H,W,C], relative poses
    [V-1,4,4] and the rays rot @ [x,y,1] [V-1,3,H*W] where the kernels cannot form them bit-exactly (sweep.rays_for).  AdaMVS sweeps a stage
    twice (pair volumes, then the weighted product): it builds the scene once."""
    with torch.no_grad():
        slots = None
        if TEXEL_POOL is not None and features[0].shape[0] == 1:        # (a batch item's slice is a new tensor every call)
            found = TEXEL_POOL.lookup(features)
            if found is not None:
                texels, slots = found
        if slots is None:
            texels = sweep.to_texels([f[b] for f in features])
        pose = sweep.relative_poses(torch.stack([p[b] for p in proj_matrices], 0))
        rays = sweep.rays_for(pose, texels.shape[1], texels.shape[2])
    return texels, pose, rays, slots


def _volume(features, proj_matrices, depth_values, mode, plane_major=False, scenes=None, **kw):
    """Cost volume for every batch item: [B,Cout,D,H,W] (or [B,D,Cout,H,W] plane-major)."""
    with torch.no_grad():
        vols = []
        for b in range(features[0].shape[0]):
            texels, pose, rays, slots = scenes[b] if scenes is not None else _scene(features, proj_matrices, b)
            w = kw.get("weights")
            vols.append(sweep.cost_volume(texels, pose, depth_values[b].contiguous(), mode, plane_major=plane_major,
                                          rays=rays, view_slots=slots, **{**kw, "weights": None if w is None else w[b]}))
        return torch.stack(vols, 0) if len(vols) > 1 else vols[0].unsqueeze(0)


def _resize_pair_weights(confidence_map, n_src, img_h, img_w):
    """adamvs.py:291-302, :498-503: the first V-1 pair-confidence maps [B,1,h',w'] resized (bilinear, align_corners=False)
    to the stage's resolution -> (list of V-1 maps [B,1,h,w], stacked weights [B,V-1,h,w]).  On the device the maps are
    stacked at THEIR size and one launch (sweep.resize_bilinear: ATen's formula) writes the stacked weights; the list
    holds views of them (upstream: V-1 upsample launches and a concatenation at the stage's full size)."""
    maps = [confidence_map[i] for i in range(n_src)]
    if (all(m.is_cuda and m.dtype == torch.float32 and m.dim() == 4 and m.shape[1] == 1 for m in maps)
            and len({tuple(m.shape) for m in maps}) == 1):
        stacked = torch.cat(maps, 1) if n_src > 1 else maps[0]
        weights = stacked if tuple(stacked.shape[-2:]) == (img_h, img_w) else sweep.resize_bilinear(stacked, (img_h, img_w))
        return [weights[:, i:i + 1] for i in range(n_src)], weights
    resized = [F.interpolate(m, [img_h, img_w], mode='bilinear', align_corners=False) for m in maps]
    return resized, torch.cat(resized, 1)


def _regress(logits, depth_values, **kw):
    """depth_regress per batch item; logits [B,D,H,W]; returns dict of stacked [B,...] tensors."""
    outs = [sweep.depth_regress(logits[b], depth_values[b].contiguous(), **kw) for b in range(logits.shape[0])]
    return {k: torch.stack([o[k] for o in outs], 0) for k in outs[0]}


# ------------------------------------------------------------------------------------- Cas-MVSNet
def cas_depthnet_forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization,
                         prob_volume_init=None):
    proj_matrices = torch.unbind(proj_matrices, 1)
    _check(features, proj_matrices, depth_values, num_depth)
    _inference_only("cas_depthnet_forward", features, depth_values)
    volume_variance = _volume(features, proj_matrices, depth_values, sweep.AGG_VARIANCE)     # :46-60
    cost_reg = cost_regularization(volume_variance)                                             # :63
    prob_volume_pre = cost_reg.squeeze(1)
    if prob_volume_init is not None:
        prob_volume_pre += prob_volume_init
    with torch.no_grad():
        r = _regress(prob_volume_pre, depth_values, conf_mode=sweep.CONF_WINDOW4)              # :69-76
    return {"depth": r["depth"], "photometric_confidence": r["conf"]}


# ---------------------------------------------------------------------------------------- RED-Net
def red_depthnet_forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization,
                         prob_volume_init=None):
    proj_matrices = torch.unbind(proj_matrices, 1)
    _check(features, proj_matrices, depth_values, num_depth)
    _inference_only("red_depthnet_forward", features, depth_values)
    volume_variance = _volume(features, proj_matrices, depth_values, sweep.AGG_VARIANCE)     # :217-230
    prob_volume_pre = cost_regularization(volume_variance).squeeze(1)
    if prob_volume_init is not None:
        prob_volume_pre += prob_volume_init
    with torch.no_grad():
        r = _regress(prob_volume_pre, depth_values, conf_mode=sweep.CONF_MAX_PROB)             # :234-238
    return {"depth": r["depth"], "photometric_confidence": r["conf"]}


# Planes of regulariser output folded into the streaming soft-argmax accumulators per launch.  Upstream updates
# its three accumulator maps after every plane (adamvs.py:514-525) because the whole design streams through an
# 11 GB GPU; on a B200 the last K planes simply stay alive (K x 20 MB at 1856 x 2752) and ONE launch folds them in,
# in plane order with the same fp32 operations -- bit-identical, and the accumulators are read and written once per
# K planes instead of once per plane (145 -> 33 MB of HBM traffic per plane at the stage-2 shape).  1 = upstream's
# cadence.  At most 16 (D3D_REGRESS_MAX_PLANES).
STREAM_BATCH_PLANES = 16


# Opt-in (row f1, second half): where the plane-at-a-time regulariser has the reference's layer list
# (SliceCostRegNetRED, adamvs.py:403-427), only its two GRU cells carry state from plane to plane, and the two
# recurrences do not feed each other's state.  So the stateless layers run ONCE over all D planes as a batch -- conv1
# before GRU 1's walk, conv2 between the two walks, upconv1 + skip + ReLU + the output convolution after GRU 2's -- and
# the soft-argmax is one launch over the whole logit volume.  Same layers, same weights, same per-plane operations;
# cuDNN may pick another algorithm for a batch of D than for a batch of 1, so logits agree to rounding, not bit for bit.
BATCH_STATELESS_CONVS = False
_SLICE_REG_LAYERS = ("conv1", "conv_gru1", "conv2", "conv_gru2", "upconv1", "upconv2d")


def batched_slice_regulariser(reg, volume, state1, state2):
    """volume [D,C,h,w] (one batch item, plane-major) -> logits [D,1,H,W] through `reg`'s own layers (see above)."""
    c1 = reg.conv1(volume)                                   # stateless: every plane at once
    r1 = []
    for k in range(volume.shape[0]):                         # recurrence 1
        out, state1 = reg.conv_gru1(c1[k:k + 1], state1)
        r1.append(out)
    r1 = torch.cat(r1, 0)
    c2 = reg.conv2(r1)                                       # stateless
    r2 = []
    for k in range(volume.shape[0]):                         # recurrence 2
        out, state2 = reg.conv_gru2(c2[k:k + 1], state2)
        r2.append(out)
    up = F.relu(torch.add(reg.upconv1(torch.cat(r2, 0)), r1))   # adamvs.py:423-424
    return reg.upconv2d(up)


class _Stream:
    """Streaming soft-argmax accumulators of the plane-at-a-time models (adamvs.py:456-462, 514-529)."""

    def __init__(self, batch, h, w, device):
        self.state = [torch.zeros((3, h, w), device=device, dtype=torch.float32) for _ in range(batch)]
        self.pending = [[] for _ in range(batch)]
        self.out = None

    def update(self, d, num_depth, reg_cost, depth_values):
        """reg_cost [B,1,H,W] = the regulariser's output for plane d."""
        last = d == num_depth - 1
        k = max(1, min(int(STREAM_BATCH_PLANES), 16))
        outs = []
        for b in range(reg_cost.shape[0]):
            held = self.pending[b]
            held.append(reg_cost[b, 0])
            if len(held) < k and not last:
                continue
            outs.append(sweep.depth_regress(list(held), depth_values[b].contiguous(), softmax_mode=sweep.SOFTMAX_RAW_EXP,
                                            d_begin=d + 1 - len(held), state=self.state[b], finalize=last))
            held.clear()
        if last:
            self.out = {k_: torch.stack([o[k_] for o in outs], 0) for k_ in ("depth", "conf")}


def red_infer_forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization):
    proj_matrices = torch.unbind(proj_matrices, 1)
    _check(features, proj_matrices, depth_values, num_depth)
    _inference_only("red_infer_forward", features, depth_values)
    ref = features[0]
    b_num, _, img_h, img_w = ref.shape
    dev = ref.device
    state1 = torch.zeros((b_num, 8, img_h, img_w), device=dev)                                  # :391-394
    state2 = torch.zeros((b_num, 16, int(img_h / 2), int(img_w / 2)), device=dev)
    state3 = torch.zeros((b_num, 32, int(img_h / 4), int(img_w / 4)), device=dev)
    state4 = torch.zeros((b_num, 64, int(img_h / 8), int(img_w / 8)), device=dev)
    if PLANE_LOOP_GRAPHS and depth_values.dim() == 4:
        loop = _plane_loop(self, "red", cost_regularization, b_num, num_depth, ref.shape[1], (img_h, img_w),
                           (img_h, img_w), [tuple(s.shape) for s in (state1, state2, state3, state4)],
                           tuple(depth_values.shape[2:]), dev)
        _volume_into(loop.volume, features, proj_matrices, depth_values, sweep.AGG_VARIANCE)
        loop.hyps.copy_(depth_values)
        depth, conf = loop.replay()
        return {"depth": depth.clone(), "photometric_confidence": conf.clone()}
    # all D variance planes in one launch, plane-major so each plane is a contiguous [C,H,W] slice
    volume = _volume(features, proj_matrices, depth_values, sweep.AGG_VARIANCE, plane_major=True)
    acc = _Stream(b_num, img_h, img_w, dev)
    for d in range(num_depth):
        reg_cost, state1, state2, state3, state4 = cost_regularization(volume[:, d], state1, state2, state3, state4)
        acc.update(d, num_depth, reg_cost, depth_values)                                         # :418-437
    return {"depth": acc.out["depth"], "photometric_confidence": acc.out["conf"]}


# ----------------------------------------------------------------------------------------- AdaMVS
def ada_infer_forward(self, features, proj_matrices, depth_values, num_depth, confidence_map=None):
    proj_matrices = torch.unbind(proj_matrices, 1)
    _check(features, proj_matrices, depth_values, num_depth)
    _inference_only("ada_infer_forward", features, depth_values, confidence_map)
    ref = features[0]
    b_num, _, img_h, img_w = ref.shape
    dev = ref.device
    n_src = len(features) - 1
    pair_confidence = []
    pair_results = []
    state1 = torch.zeros((b_num, 8, img_h, img_w), device=dev)                                  # :451-452
    state2 = torch.zeros((b_num, 16, int(img_h / 2), int(img_w / 2)), device=dev)
    up = 2 if self.in_up else 1
    acc = _Stream(b_num, img_h * up, img_w * up, dev)
    scenes = [_scene(features, proj_matrices, b) for b in range(b_num)]

    if confidence_map is None:                                                                  # :465-489
        pairs = _volume(features, proj_matrices, depth_values, sweep.AGG_PAIR_MEAN, scenes=scenes)   # [B,V-1,D,h,w]
        for i in range(n_src):
            score_volume = self.reg(pairs[:, i])
            with torch.no_grad():
                r = _regress(score_volume, depth_values, conf_mode=sweep.CONF_MAX_PROB, want_index=False)
            pair_results.append(r["depth"])
            pair_confidence.append(r["conf"].unsqueeze(1))
        confidence_map = pair_confidence

    # :498-503 -- zip() stops at the V-1 source views, so only the first V-1 maps are consumed; every
    # depth iteration appends the resized maps again (kept: callers index the list)
    resized, weights = _resize_pair_weights(confidence_map, n_src, img_h, img_w)                # [B,V-1,h,w]
    if PLANE_LOOP_GRAPHS and depth_values.dim() == 4:
        loop = _plane_loop(self, "ada", self.reg_fuse, b_num, num_depth, ref.shape[1], (img_h, img_w),
                           (img_h * up, img_w * up), [tuple(state1.shape), tuple(state2.shape)],
                           tuple(depth_values.shape[2:]), dev)
        _volume_into(loop.volume, features, proj_matrices, depth_values, sweep.AGG_WEIGHTED_PRODUCT, weights=weights,
                     scenes=scenes)
        loop.hyps.copy_(depth_values)
        depth, conf = loop.replay()
        for d in range(num_depth):
            pair_confidence.extend(resized)                                                      # :503, the list quirk
        return {"depth": depth.clone(), "photometric_confidence": conf.clone(),
                "pair_confidence": pair_confidence, "pair_result": pair_results}
    similarity = _volume(features, proj_matrices, depth_values, sweep.AGG_WEIGHTED_PRODUCT, plane_major=True,
                         weights=weights, scenes=scenes)                                         # [B,D,C,h,w]
    if BATCH_STATELESS_CONVS and all(hasattr(self.reg_fuse, n) for n in _SLICE_REG_LAYERS):
        outs = []
        for b in range(b_num):
            logits = batched_slice_regulariser(self.reg_fuse, similarity[b], state1[b:b + 1], state2[b:b + 1])
            outs.append(sweep.depth_regress(logits[:, 0], depth_values[b].contiguous(), softmax_mode=sweep.SOFTMAX_RAW_EXP,
                                            want_index=False))
        for d in range(num_depth):
            pair_confidence.extend(resized)
        return {"depth": torch.stack([o["depth"] for o in outs], 0),
                "photometric_confidence": torch.stack([o["conf"] for o in outs], 0),
                "pair_confidence": pair_confidence, "pair_result": pair_results}
    for d in range(num_depth):
        pair_confidence.extend(resized)
        reg_cost, state1, state2 = self.reg_fuse(similarity[:, d], state1, state2)               # :512
        acc.update(d, num_depth, reg_cost, depth_values)                                         # :514-529
    return {"depth": acc.out["depth"], "photometric_confidence": acc.out["conf"],
            "pair_confidence": pair_confidence, "pair_result": pair_results}


def ada_depthnet_forward(self, features, proj_matrices, depth_values, num_depth, confidence_map=None):
    """Training-form DepthNet (adamvs.py:247-312): whole volumes, epsilon in the numerator, 3-D
    regulariser `self.reg_fuse`, softmax + max-prob confidence."""
    proj_matrices = torch.unbind(proj_matrices, 1)
    _check(features, proj_matrices, depth_values, num_depth)
    _inference_only("ada_depthnet_forward", features, depth_values, confidence_map)
    n_src = len(features) - 1
    _, _, img_h, img_w = features[0].shape
    pair_confidence, pair_results = [], []
    scenes = [_scene(features, proj_matrices, b) for b in range(features[0].shape[0])]
    if confidence_map is None:
        pairs = _volume(features, proj_matrices, depth_values, sweep.AGG_PAIR_MEAN, scenes=scenes)
        for i in range(n_src):
            score_volume = self.reg(pairs[:, i])
            r = _regress(score_volume, depth_values, conf_mode=sweep.CONF_MAX_PROB, want_index=False)
            pair_results.append(r["depth"])
            pair_confidence.append(r["conf"].unsqueeze(1))
        weights = torch.cat(pair_confidence, 1)
    else:
        _, weights = _resize_pair_weights(confidence_map, n_src, img_h, img_w)
        pair_confidence = confidence_map            # adamvs.py:299: the caller's list, NOT the resized maps
    fused = _volume(features, proj_matrices, depth_values, sweep.AGG_WEIGHTED_PRODUCT, weights=weights,
                    eps_in_numerator=True, scenes=scenes)
    prob_volume_pre = self.reg_fuse(fused).squeeze(1)
    r = _regress(prob_volume_pre, depth_values, conf_mode=sweep.CONF_MAX_PROB, want_index=False)
    return {"depth": r["depth"], "photometric_confidence": r["conf"], "pair_confidence": pair_confidence,
            "pair_result": pair_results}


# ---------------------------------------------------------------------------------------- UCS-Net
def ucs_compute_depth(feats, proj_mats, depth_samps, cost_reg, lamb, is_training=False):
    proj_mats = torch.unbind(proj_mats, 1)
    assert len(proj_mats) == len(feats), "Different number of images and projection matrices"
    _inference_only("ucsnet.compute_depth", feats, depth_samps)
    volume_variance = _volume(feats, proj_mats, depth_samps, sweep.AGG_VARIANCE)               # :119-134
    prob_volume_pre = cost_reg(volume_variance).squeeze(1)
    with torch.no_grad():
        r = _regress(prob_volume_pre, depth_samps, conf_mode=sweep.CONF_WINDOW4, lamb=float(lamb))   # :137-149
    return {"depth": r["depth"], "photometric_confidence": r["conf"], "variance": r["exp_variance"]}


def ucs_uncertainty_samples(cur_depth, exp_var, ndepth, device, dtype, shape):
    """ucsnet.py:30-53.  The first-stage branch is get_depth_range_samples' range branch; the per-pixel branch
    (cur -/+ exp_var in ndepth steps, + 1e-12) is d3d_depth_samples' SPREAD mode: one launch writes the [D,H,W] planes."""
    if cur_depth.dim() == 2:
        from .module import get_depth_range_samples
        return get_depth_range_samples(cur_depth, ndepth, 0.0, device, dtype, shape)
    assert ndepth > 1
    with torch.no_grad():
        return torch.stack([sweep.depth_samples(sweep.SAMPLES_SPREAD, int(ndepth), cur_depth.shape[-2:],
                                                cur=cur_depth[b, 0].contiguous(), spread=exp_var[b, 0].contiguous())
                            for b in range(cur_depth.shape[0])], 0)


# ------------------------------------------------------------------ nn.Module wrappers (parameter-free)
class DepthNet(nn.Module):
    """cas_mvsnet.DepthNet."""
    forward = cas_depthnet_forward


class REDDepthNet(nn.Module):
    """msrednet.DepthNet."""
    forward = red_depthnet_forward


class REDInferDepthNet(nn.Module):
    """msrednet.InferDepthNet."""
    forward = red_infer_forward


class AdaInferDepthNet(nn.Module):
    """adamvs.InferDepthNet with the caller's regularisers: reg = CostRegNet2D(in_depths, base),
    reg_fuse = SliceCostRegNetRED(in_channels, in_up, base) (adamvs.py:430-434)."""

    def __init__(self, reg, reg_fuse, in_up=True):
        super().__init__()
        self.in_up = in_up
        self.reg = reg
        self.reg_fuse = reg_fuse

    forward = ada_infer_forward
