"""Tensor-level entry points of the plane-sweep engine (one call = one libd3dsweep launch).

PyTorch owns every buffer and the stream; this module only checks tensors, fills the C structs of
`include/d3d_sweep.h` and calls the library on `torch.cuda.current_stream()`.  Inputs must be fp32
CUDA tensors: there is no CPU path.

Batch is handled above this layer (`module.py` loops over B, which is 1 in the reference's
inference script, mvs/mvs_cas/predict.py:49).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (AGG_GROUP_CORR, AGG_PAIR_MEAN, AGG_VARIANCE, AGG_WARP, AGG_WEIGHTED_PRODUCT,  # noqa: F401
                   CONF_MAX_PROB, CONF_WINDOW4, HYPS_PER_PIXEL, HYPS_RESIZED, HYPS_UNIFORM,
                   SAMPLES_AROUND, SAMPLES_CASCADE, SAMPLES_RANGE, SAMPLES_SPREAD, SOFTMAX_NONE, SOFTMAX_RAW_EXP,
                   SOFTMAX_STABLE)


def _need(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: the sweep engine only runs on CUDA (no CPU fallback)" % (name, t.device))
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def to_texels(features, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Feature maps of all views, [V] x [C,H,W] (or one [V,C,H,W] tensor) -> channels-last [V,H,W,C]."""
    views = list(features) if not isinstance(features, torch.Tensor) else list(features.unbind(0))
    views = [_need(v[0] if v.dim() == 4 else v, "features[%d]" % i) for i, v in enumerate(views)]
    c, h, w = views[0].shape
    if out is None:
        out = torch.empty((len(views), h, w, c), device=views[0].device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(views[0].device):
        for i, v in enumerate(views):
            if tuple(v.shape) != (c, h, w):
                raise ValueError("features[%d] has shape %s, expected %s" % (i, tuple(v.shape), (c, h, w)))
            _lib.check(lib.d3d_nchw_to_nhwc(v.data_ptr(), out[i].data_ptr(), c, h, w, _stream()))
    return out


def resize_bilinear(maps: torch.Tensor, hw: Tuple[int, int]) -> torch.Tensor:
    """F.interpolate(maps, hw, mode="bilinear", align_corners=False) for [..., h, w] fp32 maps (ATen's formula, one launch
    for the whole stack): the resize of the AdaMVS pair confidences between stages (adamvs.py:291-302)."""
    maps = _need(maps, "maps")
    h, w = int(hw[0]), int(hw[1])
    if maps.dim() < 2:
        raise ValueError("maps must be [..., h, w]")
    n = maps.numel() // (maps.shape[-2] * maps.shape[-1])
    out = torch.empty(tuple(maps.shape[:-2]) + (h, w), device=maps.device, dtype=torch.float32)
    with torch.cuda.device(maps.device):
        _lib.check(_lib.load().d3d_resize_bilinear(maps.data_ptr(), out.data_ptr(), n, maps.shape[-2], maps.shape[-1], h, w,
                                                   _stream()))
    return out


def relative_poses(proj: torch.Tensor) -> torch.Tensor:
    """[V,4,4] projection matrices (view 0 = reference) -> [V-1,4,4] P_src @ inverse(P_ref).

    Computed with the very calls of the reference (mvs/mvs_cas/models/module.py:528, one [1,4,4] matmul
    per source view) so the kernel sees bit for bit the matrices the reference sees: a batched matmul
    rounds 1.6 % of the entries differently on B200 (tools/diag_coords.py), which is enough to move
    projected coordinates by an ulp and the cost volume by 1e-4 at production image sizes."""
    # linalg.inv_ex is torch.inverse without the host-side error check (no device->host sync)
    inv = torch.linalg.inv_ex(proj[0:1], check_errors=False).inverse
    return torch.cat([torch.matmul(proj[i:i + 1], inv) for i in range(1, proj.shape[0])], 0).contiguous()


_PIXEL_GRIDS = {}


def _pixel_grid(h: int, w: int, device) -> torch.Tensor:
    """[1,3,H*W] homogeneous pixel coordinates, built with the reference's calls (module.py:532-537)."""
    key = (h, w, str(device))
    if key not in _PIXEL_GRIDS:
        y, x = torch.meshgrid([torch.arange(0, h, dtype=torch.float32, device=device),
                               torch.arange(0, w, dtype=torch.float32, device=device)], indexing="ij")
        y, x = y.contiguous().view(h * w), x.contiguous().view(h * w)
        _PIXEL_GRIDS[key] = torch.unsqueeze(torch.stack((x, y, torch.ones_like(x))), 0)
    return _PIXEL_GRIDS[key]


def reference_rays(pose: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[V-1,4,4] relative poses -> [V-1,3,H*W] = rot_i @ [x,y,1], by the reference's own call (module.py:538:
    one [1,3,3] @ [1,3,H*W] torch.matmul per source view).

    cuBLAS does not round this 3-term product the same way at every problem size: at H*W = 1376*928 (cascade
    stage 2 of the 1856x2752 configuration) the columns past 2^20 - 32 leave the fma(r2,1,fma(r1,y,r0*x))
    order it uses everywhere else (tools/diag_rays.py, diag_rays_n.py), which moves 5 % of those rays by an ulp
    and the cost volume by 1e-4.  Handing the kernel the reference's own product makes the sampling
    coordinates bit-identical whatever the library does."""
    xyz = _pixel_grid(h, w, pose.device)
    rays = torch.empty((pose.shape[0], 3, h * w), device=pose.device, dtype=torch.float32)
    for i in range(pose.shape[0]):
        torch.matmul(pose[i:i + 1, :3, :3], xyz, out=rays[i:i + 1])
    return rays


def kernel_rays(pose: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[V-1,3,H*W] rays as the sweep kernels form them when none are handed in: fma(r2, 1, fma(r1, y, r0*x))."""
    pose = _need(pose, "pose")
    out = torch.empty((pose.shape[0], 3, h * w), device=pose.device, dtype=torch.float32)
    with torch.cuda.device(pose.device):
        _lib.check(_lib.load().d3d_pixel_rays(pose.data_ptr(), pose.shape[0], h, w, out.data_ptr(), _stream()))
    return out


# (H, W, device, matmul settings, library versions) -> does cuBLAS round the reference's rot @ [x,y,1] in the kernels'
# own order at this size?
_RAY_ORDER = {}


def _ray_key(h: int, w: int, device) -> tuple:
    """What the verdict of `rays_for` depends on besides the image size: the device, and everything that can make
    torch.matmul pick another cuBLAS kernel or another arithmetic for an fp32 product -- TF32 permission, the
    float32 matmul precision, the CUDA / torch build.  A change of any of them re-runs the probe."""
    return (h, w, torch.device(device).index, bool(torch.backends.cuda.matmul.allow_tf32),
            torch.get_float32_matmul_precision(), torch.version.cuda, torch.__version__)
_PROBE = ((0.9931, -0.0713, 13.71, 0.0), (0.0527, 1.0117, -22.37, 0.0), (1.3e-5, -2.1e-5, 1.0031, 0.0), (0.0, 0.0, 0.0, 1.0))


def rays_for(pose: torch.Tensor, h: int, w: int) -> Optional[torch.Tensor]:
    """The `rays` argument of `cost_volume` for this pose: None where the kernels' own rays are bit-identical to the
    reference's matmul, else `reference_rays(pose, h, w)`.

    Which it is depends on the kernel cuBLAS picks for a [1,3,3] @ [1,3,H*W] product, i.e. on the image size, not
    on the values: the first call at a size computes both forms, for a generic dense probe rotation AND for the
    pose at hand, and compares them bit for bit (one host sync per size and device, so call it outside CUDA-graph
    capture); every later call at that size reuses the verdict.  At 1856 x 2752 the skipped matmuls are 0.9 ms of
    a 10.7 ms AdaMVS view.  `D3D_RAYS=matmul` in the environment forces the matmul everywhere."""
    if os.environ.get("D3D_RAYS", "") == "matmul":           # escape hatch: always the reference's own product
        return reference_rays(pose, h, w)
    key = _ray_key(h, w, pose.device)
    same = _RAY_ORDER.get(key)
    if same is None:
        probe = torch.tensor(_PROBE, dtype=torch.float32, device=pose.device).unsqueeze(0)
        ref = reference_rays(pose, h, w)
        same = bool(torch.equal(reference_rays(probe, h, w), kernel_rays(probe, h, w))) and \
            bool(torch.equal(ref, kernel_rays(pose, h, w)))
        _RAY_ORDER[key] = same
        return None if same else ref
    return None if same else reference_rays(pose, h, w)


def cost_volume(texels: torch.Tensor, pose: torch.Tensor, hyps: torch.Tensor, mode: int = AGG_VARIANCE, *,
                groups: int = 0, weights: Optional[torch.Tensor] = None, eps_in_numerator: bool = False,
                d_begin: int = 0, d_count: int = 0, out: Optional[torch.Tensor] = None,
                plane_major: bool = False, variant: int = 0, rays: Optional[torch.Tensor] = None,
                exact_rays: bool = True, view_slots=None) -> torch.Tensor:
    """Fused warp + aggregate.  texels [V,H,W,C]; pose [V-1,4,4]; hyps [D] or [D,H,W].

    view_slots: `texels` is then a POOL [S,H,W,C] of per-image maps (laid out once per image, `to_texels(..., out=pool[s])`)
    and view v of this sweep is pool slot view_slots[v] (v = 0: the reference) -- the reference views of a scene block share
    their images, so nothing is relaid out or copied per view.

    rays: `rays_for(pose, H, W)` if the caller already has them (plane-slice callers compute them once per view);
    with exact_rays (default) that call is made here -- the reference's own matmul wherever the kernel's rounding
    order is not known to match it; with exact_rays=False the kernel forms them unconditionally.

    Returns [Cout, Dn, H, W] (or [Dn, Cout, H, W] with plane_major=True, the layout whose planes are
    contiguous [Cout,H,W] slices for the plane-at-a-time regularisers), Dn = planes computed.
    """
    texels = _need(texels, "texels")
    pose = _need(pose, "pose")
    hyps = _need(hyps, "hyps")
    v, h, w, c = texels.shape
    slots = None
    if view_slots is not None:                    # `texels` is a pool [S,h,w,c] of per-image maps; the views are named by slot
        slots = [int(x) for x in view_slots]
        if not 2 <= len(slots) <= 9 or any(x < 0 or x >= v for x in slots):
            raise ValueError("view_slots must name 2..9 slots of the %d-slot texel pool, got %s" % (v, slots))
        pool_slots, v = v, len(slots)
    if tuple(pose.shape) != (v - 1, 4, 4):
        raise ValueError("pose must be [%d,4,4], got %s" % (v - 1, tuple(pose.shape)))
    if hyps.dim() == 1:
        per_pixel = 0
    elif hyps.dim() == 3 and tuple(hyps.shape[1:]) == (h, w):
        per_pixel = 1
    else:
        raise ValueError("hyps must be [D] or [D,%d,%d], got %s" % (h, w, tuple(hyps.shape)))
    d = hyps.shape[0]
    dn = d - d_begin if d_count <= 0 else d_count
    if mode == AGG_GROUP_CORR:
        cout = groups
    elif mode == AGG_PAIR_MEAN:
        cout = v - 1
    else:
        cout = c
    if weights is not None:
        weights = _need(weights, "weights")
        if tuple(weights.shape) != (v - 1, h, w):
            raise ValueError("weights must be [%d,%d,%d], got %s" % (v - 1, h, w, tuple(weights.shape)))
    shape = (dn, cout, h, w) if plane_major else (cout, dn, h, w)
    if out is None:
        out = torch.empty(shape, device=texels.device, dtype=torch.float32)
    else:
        _need(out, "out")
        if tuple(out.shape) != shape or not out.is_contiguous():
            raise ValueError("out must be a contiguous %s tensor" % (shape,))
    a = _lib.CostVolumeArgs()
    a.struct_size = C.sizeof(a)
    a.mode, a.num_views, a.channels, a.height, a.width, a.num_depth = mode, v, c, h, w, d
    a.d_begin, a.d_count, a.hyps_per_pixel = d_begin, dn, per_pixel
    a.groups, a.eps_in_numerator, a.variant = groups, int(eps_in_numerator), variant
    if slots is not None:
        a.texel_slots = pool_slots
        for i, x in enumerate(slots):
            a.view_slot[i] = x
    if rays is None and exact_rays:
        rays = rays_for(pose, h, w)
    if rays is not None:
        rays = _need(rays, "rays")
        if tuple(rays.shape) != (v - 1, 3, h * w):
            raise ValueError("rays must be [%d,3,%d], got %s" % (v - 1, h * w, tuple(rays.shape)))
    a.feats, a.pose, a.hyps, a.weights, a.out = texels.data_ptr(), pose.data_ptr(), hyps.data_ptr(), _ptr(weights), out.data_ptr()
    a.rays = _ptr(rays)
    if plane_major:
        a.out_stride_c, a.out_stride_d = h * w, cout * h * w
    else:
        a.out_stride_c, a.out_stride_d = dn * h * w, h * w
    with torch.cuda.device(texels.device):
        _lib.check(_lib.load().d3d_cost_volume(C.byref(a), _stream()))
    return out


def depth_regress(logits: torch.Tensor, hyps: torch.Tensor, *, conf_mode: int = CONF_MAX_PROB,
                  softmax_mode: int = SOFTMAX_STABLE, num_depth: Optional[int] = None, d_begin: int = 0,
                  state: Optional[torch.Tensor] = None, finalize: bool = True, want_index: bool = True,
                  lamb: Optional[float] = None, next_num_depth: int = 0, next_interval: float = 0.0):
    """Fused softmax + expectation + confidence over logits [Dn,H,W] (a strided view along D is fine).

    hyps: [D] (uniform), [D,H,W] (per pixel) or [D,h',w'] (bilinearly resized to H x W, align_corners
    False).  Returns a dict with "depth", "conf" and optionally "index", "exp_variance", "next_hyps",
    "state".  With softmax_mode=SOFTMAX_RAW_EXP the call may cover a slice of planes starting at
    d_begin and carries its accumulators in `state` [3,H,W]; the slice may then also be a LIST of up to 16
    separately allocated [H,W] planes (what a plane-at-a-time regulariser returns call after call).
    """
    planes = None
    if isinstance(logits, (list, tuple)):
        if softmax_mode == SOFTMAX_STABLE or not 1 <= len(logits) <= _lib.REGRESS_MAX_PLANES:
            raise ValueError("a list of planes needs RAW_EXP / NONE and 1..%d entries" % _lib.REGRESS_MAX_PLANES)
        planes = []
        for t in logits:
            if not t.is_cuda or t.dtype != torch.float32:
                raise RuntimeError("logits must be fp32 CUDA tensors (no CPU fallback)")
            if t.dim() < 2 or t.numel() != t.shape[-2] * t.shape[-1] or (planes and t.shape[-2:] != planes[0].shape):
                raise ValueError("every plane must be one [H,W] map of the same extent")
            planes.append(t.reshape(t.shape[-2], t.shape[-1]).contiguous())
        logits = planes[0].unsqueeze(0)
        dn_list = len(planes)
    if logits.dim() != 3:
        raise ValueError("logits must be [D,H,W]")
    if not logits.is_cuda or logits.dtype != torch.float32:
        raise RuntimeError("logits must be an fp32 CUDA tensor (no CPU fallback)")
    dn, h, w = logits.shape
    if planes is not None:
        dn = dn_list
    if logits.stride(2) != 1 or logits.stride(1) != w:
        logits = logits.contiguous()
    stride_d = logits.stride(0) if dn > 1 and planes is None else h * w
    hyps = _need(hyps, "hyps")
    d = int(num_depth) if num_depth is not None else hyps.shape[0]
    if hyps.shape[0] != d:
        raise ValueError("hyps has %d planes, expected %d" % (hyps.shape[0], d))
    a = _lib.RegressArgs()
    a.struct_size = C.sizeof(a)
    if hyps.dim() == 1:
        a.hyps_mode = HYPS_UNIFORM
    elif hyps.dim() == 3 and tuple(hyps.shape[1:]) == (h, w):
        a.hyps_mode = HYPS_PER_PIXEL
    elif hyps.dim() == 3:
        a.hyps_mode, a.hyps_height, a.hyps_width = HYPS_RESIZED, hyps.shape[1], hyps.shape[2]
    else:
        raise ValueError("hyps must be [D], [D,H,W] or [D,h,w]")
    dev = logits.device
    res = {}
    raw = softmax_mode != SOFTMAX_STABLE
    if raw and state is None and not (d_begin == 0 and dn == d):
        raise ValueError("a RAW_EXP plane slice needs the `state` tensor")
    if state is not None:
        state = _need(state, "state")
        if tuple(state.shape) != (3, h, w):
            raise ValueError("state must be [3,%d,%d]" % (h, w))
        res["state"] = state
    do_final = finalize or not raw
    if do_final:
        res["depth"] = torch.empty((h, w), device=dev, dtype=torch.float32)
        res["conf"] = torch.empty((h, w), device=dev, dtype=torch.float32)
        if want_index and not raw:
            res["index"] = torch.empty((h, w), device=dev, dtype=torch.int32)
        if lamb is not None:
            res["exp_variance"] = torch.empty((h, w), device=dev, dtype=torch.float32)
        if next_num_depth > 0:
            res["next_hyps"] = torch.empty((next_num_depth, h, w), device=dev, dtype=torch.float32)
    a.num_depth, a.height, a.width, a.d_begin, a.d_count = d, h, w, d_begin, dn
    a.softmax_mode, a.conf_mode, a.finalize = softmax_mode, conf_mode, int(do_final)
    a.next_num_depth = next_num_depth if do_final else 0
    a.next_interval = float(next_interval)
    a.lamb = float(lamb) if lamb is not None else 0.0
    a.logits, a.logits_stride_d, a.hyps = (logits.data_ptr() if planes is None else None), stride_d, hyps.data_ptr()
    if planes is not None:
        for k, t in enumerate(planes):
            a.logit_planes[k] = t.data_ptr()
    a.depth, a.conf, a.index = _ptr(res.get("depth")), _ptr(res.get("conf")), _ptr(res.get("index"))
    a.state, a.exp_variance, a.next_hyps = _ptr(state), _ptr(res.get("exp_variance")), _ptr(res.get("next_hyps"))
    with torch.cuda.device(dev):
        _lib.check(_lib.load().d3d_depth_regress(C.byref(a), _stream()))
    return res


def depth_samples(mode: int, num_depth: int, hw: Tuple[int, int], *, device=None, cur: Optional[torch.Tensor] = None,
                  interval: float = 0.0, dmin: float = 0.0, dmax: float = 0.0,
                  full_hw: Optional[Sequence[int]] = None, spread: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Depth hypotheses [D,H,W] for a stage (include/d3d_sweep.h: D3D_SAMPLES_*)."""
    h, w = int(hw[0]), int(hw[1])
    if cur is not None:
        cur = _need(cur, "cur")
        device = cur.device
    if device is None:
        raise ValueError("device is required when cur is None")
    if torch.device(device).type != "cuda":
        raise RuntimeError("depth_samples asked for %s: the sweep engine only runs on CUDA (no CPU fallback)" % (device,))
    out = torch.empty((num_depth, h, w), device=device, dtype=torch.float32)
    a = _lib.SamplesArgs()
    a.struct_size = C.sizeof(a)
    a.mode, a.num_depth, a.height, a.width = mode, num_depth, h, w
    if cur is not None:
        if cur.dim() != 2:
            raise ValueError("cur must be [h,w]")
        a.src_height, a.src_width = cur.shape
        if mode in (SAMPLES_AROUND, SAMPLES_SPREAD) and tuple(cur.shape) != (h, w):
            raise ValueError("cur must be [%d,%d] for SAMPLES_AROUND / SAMPLES_SPREAD" % (h, w))
    if mode == SAMPLES_SPREAD:
        if cur is None or spread is None:
            raise ValueError("SAMPLES_SPREAD needs cur and spread")
        spread = _need(spread, "spread")
        if tuple(spread.shape) != (h, w):
            raise ValueError("spread must be [%d,%d]" % (h, w))
    if full_hw is not None:
        a.full_height, a.full_width = int(full_hw[0]), int(full_hw[1])
    else:
        a.full_height, a.full_width = h, w
    a.interval, a.dmin, a.dmax = float(interval), float(dmin), float(dmax)
    a.cur, a.out, a.spread = _ptr(cur), out.data_ptr(), _ptr(spread if mode == SAMPLES_SPREAD else None)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().d3d_depth_samples(C.byref(a), _stream()))
    return out
