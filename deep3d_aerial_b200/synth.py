"""Seeded synthetic scenes for parity tests and the benchmark.

Nothing here is on the hot path: it only manufactures inputs with the tensor
contract that the reference's dataset hands to the cascade networks
(reference `mvs/mvs_cas/datasets/cas_normal_eval.py:138-162`): per-view 4x4
projection matrices ``[[K @ [R|t]], [0,0,0,1]]`` with rows 0-1 divided by
{4,2,1} for stage {1,2,3}, a ``[dmin, dmax]`` depth range and per-view feature
maps.  The rig is WHU-OMVS shaped: a nadir reference camera and source cameras
on a cross (+-x, +-y) converging on the scene centre.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Rig:
    """A reference camera plus V-1 source cameras, full-resolution intrinsics."""

    width: int            # full-resolution image width  (columns)
    height: int           # full-resolution image height (rows)
    focal: float
    z_mean: float
    dmin: float
    dmax: float
    proj_full: np.ndarray  # [V,4,4] float32, K @ [R|t] at full resolution

    def proj(self, scale: int) -> np.ndarray:
        """Projection matrices for a stage whose features are 1/scale of full res."""
        p = self.proj_full.copy()
        p[:, :2, :] = self.proj_full[:, :2, :] / np.float32(scale)
        return p

    @property
    def depth_range(self) -> np.ndarray:
        return np.array([self.dmin, self.dmax], dtype=np.float32)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def make_rig(num_views=5, width=1856, height=2752, focal=4000.0, z_mean=500.0,
             baseline_frac=0.08, range_frac=0.2) -> Rig:
    """Cross rig: ref at the origin looking down +z, sources at +-baseline on x
    then y, each rotated about the orthogonal axis so its optical axis passes
    through (0, 0, z_mean) (convergence atan(baseline/z_mean), 4.6 deg by default)."""
    K = np.array([[focal, 0, (width - 1) / 2.0],
                  [0, focal, (height - 1) / 2.0],
                  [0, 0, 1]], dtype=np.float64)
    b = baseline_frac * z_mean
    centres = [np.zeros(3)]
    rots = [np.eye(3)]
    offsets = [(+b, 0), (-b, 0), (0, +b), (0, -b), (+b, +b), (-b, -b), (+b, -b), (-b, +b)]
    for i in range(num_views - 1):
        ox, oy = offsets[i % len(offsets)]
        ring = 1 + i // len(offsets)
        ox, oy = ox * ring, oy * ring
        c = np.array([ox, oy, 0.0])
        # world->camera rotation turning the optical axis towards the scene centre
        R = _rot_y(math.atan2(ox, z_mean)).T @ _rot_x(-math.atan2(oy, z_mean)).T
        centres.append(c)
        rots.append(R)
    projs = []
    for R, c in zip(rots, centres):
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = -R @ c
        P = E.copy()
        P[:3, :4] = K @ E[:3, :4]
        projs.append(P.astype(np.float32))
    return Rig(width=width, height=height, focal=focal, z_mean=z_mean,
               dmin=(1 - range_frac) * z_mean, dmax=(1 + range_frac) * z_mean,
               proj_full=np.stack(projs))


def tiny_rig(num_views=3, width=160, height=128) -> Rig:
    """Config 1 of BASELINE.json: f=200, baselines +-0.5, depth 5..15."""
    return make_rig(num_views=num_views, width=width, height=height, focal=200.0,
                    z_mean=10.0, baseline_frac=0.05, range_frac=0.5)


def make_features(num_views, channels, h, w, seed=0, smooth=False, device="cpu"):
    """N(0,1) feature maps [V,C,h,w]; `smooth` low-passes them (avg_pool 5x5,
    re-normalised) so that bilinear errors are not hidden by white noise."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    f = torch.randn(num_views, channels, h, w, generator=g, dtype=torch.float32)
    if smooth:
        f = torch.nn.functional.avg_pool2d(f, 5, stride=1, padding=2)
        f = f / f.std()
    return f.to(device)


def uniform_hypotheses(dmin, dmax, num_depth, device="cpu"):
    """Stage-1 plane sweep: D fronto-parallel planes, [D]."""
    return torch.linspace(float(dmin), float(dmax), num_depth, dtype=torch.float32, device=device)


def smooth_depth_map(rig: Rig, h, w, seed=0, device="cpu"):
    """A tilted plane plus a low-frequency sinusoid inside the depth range, [h,w]."""
    ys = torch.linspace(-1, 1, h).view(h, 1)
    xs = torch.linspace(-1, 1, w).view(1, w)
    ph = 0.37 * (seed + 1)
    mid = 0.5 * (rig.dmin + rig.dmax)
    amp = 0.25 * (rig.dmax - rig.dmin)
    d = mid + amp * (0.5 * xs - 0.3 * ys + 0.4 * torch.sin(3.1 * xs + ph) * torch.cos(2.3 * ys - ph))
    return d.to(torch.float32).to(device)


def per_pixel_hypotheses(cur_depth, num_depth, interval):
    """Stage>=2 hypotheses around `cur_depth` [h,w] -> [D,h,w]
    (same arithmetic as reference module.py:616-630, stated independently here
    only to make inputs; parity of the real resampler is tested separately)."""
    lo = cur_depth - num_depth / 2 * interval
    hi = cur_depth + num_depth / 2 * interval
    step = (hi - lo) / (num_depth - 1)
    k = torch.arange(num_depth, dtype=cur_depth.dtype, device=cur_depth.device).view(-1, 1, 1)
    return lo.unsqueeze(0) + k * step.unsqueeze(0)


def planted_logits(num_depth, h, w, seed=0, device="cpu"):
    """Kernel-2 input: 4*N(0,1) with a +8 peak planted at a random plane per pixel."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = 4.0 * torch.randn(num_depth, h, w, generator=g, dtype=torch.float32)
    peak = torch.randint(0, num_depth, (1, h, w), generator=g)
    x.scatter_add_(0, peak, torch.full((1, h, w), 8.0))
    return x.to(device)


def fusion_scene(num_src=4, height=96, width=128, seed=0, focal=None, z_mean=500.0, baseline_frac=0.04,
                 noise=2e-3, outliers=0.1, invalid=0.03, shift=0.0, src_pad=0):
    """Depth-map fusion inputs with the dtypes `fuse/fusion_3d_normal.py:112-133, 405-515`
    hands `ConsistencyChecker.check`: per view a float32 depth map [H,W], camera-frame normals [H,W,3], a
    confidence map, float32 K [3,3] and Tcw [4,4].

    The scene is one tilted world plane seen by the cross rig of `make_rig`, so every view's depth is exact by
    ray-plane intersection and views agree wherever they are not disturbed: relative depth noise `noise`
    (threshold of the check: 1 %), a fraction `outliers` of pixels 5-30 % off, a fraction `invalid` at depth 0,
    normals = the plane normal plus noise, confidence uniform in [0, 1].  `shift` moves the source cameras
    sideways (in units of the image footprint) so projections leave the source maps (CuPy's wrap-around);
    `src_pad` grows the source maps by that many pixels on every side (principal point moved along), so that
    every projection lands inside them.
    -> dict(ref=(depth, normal, K, E, prob), src=[(depth, normal, K, E), ...])"""
    rng = np.random.default_rng(seed)
    focal = float(focal) if focal else 1.6 * width
    rig = make_rig(num_views=num_src + 1, width=width, height=height, focal=focal, z_mean=z_mean,
                   baseline_frac=baseline_frac)
    K = np.array([[focal, 0, (width - 1) / 2.0], [0, focal, (height - 1) / 2.0], [0, 0, 1]], dtype=np.float64)
    n_world = np.array([0.08, -0.05, -1.0])
    n_world /= np.linalg.norm(n_world)
    offset = float(n_world @ np.array([0.0, 0.0, z_mean]))           # plane: n . X = offset
    K_ref = K
    views = []
    for v in range(num_src + 1):
        E = np.linalg.inv(K_ref) @ rig.proj_full[v].astype(np.float64)[:3, :4]      # [R|t]
        pad = src_pad if v > 0 else 0
        height, width = rig.height + 2 * pad, rig.width + 2 * pad
        K = K_ref.copy()
        K[0, 2] += pad
        K[1, 2] += pad
        ys, xs = np.mgrid[0:height, 0:width]
        rays_cam = np.linalg.inv(K) @ np.stack([xs.ravel(), ys.ravel(), np.ones(height * width)]).astype(np.float64)
        R, t = E[:, :3].copy(), E[:, 3].copy()
        if v > 0 and shift:
            t[0] += shift * rig.width * z_mean / focal
        centre = -R.T @ t
        dirs = R.T @ rays_cam                                        # world directions with camera z = 1
        depth = ((offset - n_world @ centre) / (n_world @ dirs)).reshape(height, width)
        depth = depth * (1 + noise * rng.standard_normal((height, width)))
        bad = rng.random((height, width)) < outliers
        depth = np.where(bad, depth * (1 + rng.uniform(0.05, 0.3, (height, width)) * rng.choice([-1, 1], (height, width))),
                         depth)
        depth = np.where(rng.random((height, width)) < invalid, 0.0, depth).astype(np.float32)
        normal = (R @ n_world)[None, None, :] + 0.05 * rng.standard_normal((height, width, 3))
        flip = rng.random((height, width)) < 0.05                    # a few normals point the other way
        normal = np.where(flip[..., None], -normal, normal).astype(np.float32)
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        views.append((depth, normal, K.astype(np.float32), T.astype(np.float32)))
    prob = rng.random((rig.height, rig.width)).astype(np.float32)
    d, n, k, e = views[0]
    return {"ref": (d, n, k, e, prob), "src": views[1:]}


class _GruCell(torch.nn.Module):
    """A convolutional GRU cell of the shape the reference's recurrent regularisers use (module.py:5-50: one 3x3
    convolution for the reset / update gates, one for the candidate, `u*h + (1-u)*tanh(.)`).  Written from that
    description for the synthetic regulariser below; weights are seeded, not trained."""

    def __init__(self, channels):
        super().__init__()
        self.gates = torch.nn.Conv2d(2 * channels, 2 * channels, 3, padding=1)
        self.cand = torch.nn.Conv2d(2 * channels, channels, 3, padding=1)

    def forward(self, x, h):
        r, u = torch.sigmoid(self.gates(torch.cat((x, h), 1))).chunk(2, 1)
        c = torch.tanh(self.cand(torch.cat((x, r * h), 1)))
        out = u * h + (1 - u) * c
        return out, out


class _ConvRelu(torch.nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)

    def forward(self, x):
        return torch.nn.functional.relu(self.conv(x))


class SliceRegulariser(torch.nn.Module):
    """A seeded stand-in with the LAYER LIST of the reference's plane-at-a-time regulariser (SliceCostRegNetRED,
    adamvs.py:403-427: ConvReLU -> ConvGRU -> strided ConvReLU -> ConvGRU -> transposed conv + skip + ReLU -> output
    conv, x2 up-sampling when `up`), under the attribute names the reference uses, so that timing and batching code
    sees what it would see in production.  `forward(cost [B,C,h,w], state1, state2) -> (logit [B,1,H,W], state1, state2)`."""

    def __init__(self, in_channels, up=True, base=8, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.conv1 = _ConvRelu(in_channels, base, 1)
        self.conv_gru1 = _GruCell(base)
        self.conv2 = _ConvRelu(base, 2 * base, 2)
        self.conv_gru2 = _GruCell(2 * base)
        self.upconv1 = torch.nn.ConvTranspose2d(2 * base, base, 3, stride=2, padding=1, output_padding=1)
        self.upconv2d = (torch.nn.ConvTranspose2d(base, 1, 3, stride=2, padding=1, output_padding=1) if up
                         else torch.nn.Conv2d(base, 1, 3, padding=1))
        with torch.no_grad():
            for prm in self.parameters():
                prm.copy_(0.15 * torch.randn(prm.shape, generator=g))

    def forward(self, cost, state1, state2):
        c1 = self.conv1(cost)
        r1, state1 = self.conv_gru1(c1, state1)
        r2, state2 = self.conv_gru2(self.conv2(r1), state2)
        up = torch.nn.functional.relu(self.upconv1(r2) + r1)
        return self.upconv2d(up), state1, state2
