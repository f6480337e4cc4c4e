"""Drop-in replacements for the hot-path functions of the reference's
`mvs/mvs_cas/models/module.py` -- same names, argument order, shapes and error behaviour
(SURVEY.md §8b level b3), executed by libd3dsweep on the B200.

    homo_warping_float / homo_warping_double   module.py:516-601
    depth_regression                           module.py:605-613
    get_cur_depth_range_samples                module.py:616-630
    get_depth_range_samples                    module.py:633-650

The reference's model files do `from .module import *`; `deep3d_aerial_b200.install()` rebinds these
names inside the reference's modules.
"""
from __future__ import annotations

import torch

from . import sweep

__all__ = ["homo_warping_float", "homo_warping_double", "depth_regression", "get_cur_depth_range_samples",
           "get_depth_range_samples"]


def _hyps_item(depth_values, b):
    """[B,D] or [B,D,H,W] -> the [D] / [D,H,W] hypotheses of batch item b."""
    return depth_values[b]


def homo_warping_float(src_fea, src_proj, ref_proj, depth_values):
    """src_fea [B,C,H,W], src_proj/ref_proj [B,4,4], depth_values [B,D] or [B,D,H,W] -> [B,C,D,H,W].

    Materialises the warped volume, as the reference does; the cascade networks in this package never
    call it (they use the fused aggregation) -- it exists for callers that want the warp itself.
    """
    batch, channels, height, width = src_fea.shape
    num_depth = depth_values.shape[1]
    out = torch.empty((batch, channels, num_depth, height, width), device=src_fea.device, dtype=torch.float32)
    with torch.no_grad():
        proj = torch.matmul(src_proj, torch.inverse(ref_proj))      # module.py:528
        for b in range(batch):
            texels = sweep.to_texels([src_fea[b], src_fea[b]])      # slot 0 (reference view) is unused by WARP
            sweep.cost_volume(texels, proj[b:b + 1].contiguous(), _hyps_item(depth_values, b).contiguous(),
                              sweep.AGG_WARP, out=out[b])
    return out


def homo_warping_double(src_fea, src_proj, ref_proj, depth_values):
    """module.py:560-601: the warp with the coordinate arithmetic in fp64.  src_fea [B,C,H,W] fp32, src_proj / ref_proj
    [B,4,4] **fp64** (upstream multiplies `rot` by an fp64 pixel grid and torch.matmul does not promote, so fp32
    projections fail there with a dtype error; here too), depth_values [B,D] or [B,D,H,W] -> [B,C,D,H,W] fp32.
    No model calls it; d3d_homo_warp_f64 keeps the contract: fp64 up to the normalised grid, fp32 sampling."""
    import ctypes as C

    from . import _lib

    if src_proj.dtype != torch.float64 or ref_proj.dtype != torch.float64:
        raise RuntimeError("expected m1 and m2 to have the same dtype, but got: %s != double (homo_warping_double needs "
                           "fp64 projection matrices, as the reference does)" % str(src_proj.dtype).replace("torch.", ""))
    if not src_fea.is_cuda:
        raise RuntimeError("src_fea is on %s: the sweep engine only runs on CUDA (no CPU fallback)" % src_fea.device)
    batch, channels, height, width = src_fea.shape
    num_depth = depth_values.shape[1]
    out = torch.empty((batch, channels, num_depth, height, width), device=src_fea.device, dtype=torch.float32)
    with torch.no_grad():
        proj = torch.matmul(src_proj, torch.inverse(ref_proj)).contiguous()       # module.py:573, in fp64
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream(src_fea.device).cuda_stream)
        for b in range(batch):
            texels = sweep.to_texels([src_fea[b].float()])[0]
            hyps = _hyps_item(depth_values, b).float().contiguous()
            with torch.cuda.device(src_fea.device):
                _lib.check(lib.d3d_homo_warp_f64(texels.data_ptr(), proj[b].data_ptr(), hyps.data_ptr(), int(hyps.dim() == 3),
                                                 channels, num_depth, height, width, out[b].data_ptr(), stream))
    return out


def depth_regression(p, depth_values):
    """p [B,D,H,W] probability volume, depth_values [B,D] or [B,D,h,w] -> [B,H,W] = sum_d p*d.
    4-D hypotheses at another resolution are bilinearly resized (align_corners=False) on the fly."""
    batch = p.shape[0]
    out = []
    for b in range(batch):
        r = sweep.depth_regress(p[b], depth_values[b].contiguous(), softmax_mode=sweep.SOFTMAX_NONE, want_index=False)
        out.append(r["depth"])
    return torch.stack(out, 0)


def get_cur_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, shape, max_depth=192.0, min_depth=0.0):
    """cur_depth [B,H,W] -> [B,D,H,W]: cur -/+ ndepth/2*interval in ndepth steps (max/min unused upstream)."""
    assert cur_depth.shape == torch.Size(shape), "cur_depth:{}, input shape:{}".format(cur_depth.shape, shape)
    return torch.stack([sweep.depth_samples(sweep.SAMPLES_AROUND, ndepth, cur_depth.shape[1:], cur=cur_depth[b],
                                            interval=float(depth_inteval_pixel))
                        for b in range(cur_depth.shape[0])], 0)


def get_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, device, dtype, shape, max_depth=192.0,
                            min_depth=0.0):
    """cur_depth [B,H,W] (previous estimate) or [B,2+] (a [dmin .. dmax] range) -> [B,D,H,W]."""
    if cur_depth.dim() == 2:
        lo = cur_depth[:, 0].tolist()
        hi = cur_depth[:, -1].tolist()
        return torch.stack([sweep.depth_samples(sweep.SAMPLES_RANGE, ndepth, (shape[1], shape[2]),
                                                device=cur_depth.device, dmin=lo[b], dmax=hi[b])
                            for b in range(cur_depth.shape[0])], 0)
    return get_cur_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, shape, max_depth, min_depth)
