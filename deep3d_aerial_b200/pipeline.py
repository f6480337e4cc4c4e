"""Host-to-host view pipeline: reference views arrive as pinned HOST tensors, depth + confidence maps
leave as pinned HOST tensors, and the host<->device copies of neighbouring views overlap the sweep.

    pipe = ViewPipeline(num_views=5, channels=32, height=688, width=464, num_depth=384, device="cuda:0")
    for i, view in enumerate(views):
        pipe.submit(view.features, view.proj, view.hyps, logits_fn)     # returns at once
        if i: depth, conf = pipe.collect()                              # result of view i-1
    depth, conf = pipe.collect()

This mirrors what the reference's inference loop does per reference view (mvs/mvs_cas/predict.py:126-133:
`tocuda(sample)`, `model(...)`, `tensor2numpy(outputs)`), but with two input slots, a copy stream and
events instead of a synchronous round trip.  The CNN regulariser between the two kernels is the caller's
(`logits_fn(volume) -> [D,H,W] logits`, run on the compute stream).  CUDA only: there is no CPU path.
"""
from __future__ import annotations

import collections
from typing import Callable, Optional

import torch

from . import sweep


class ViewPipeline:
    def __init__(self, num_views: int, channels: int, height: int, width: int, num_depth: int, device,
                 mode: int = sweep.AGG_VARIANCE, groups: int = 0, conf_mode: int = sweep.CONF_MAX_PROB,
                 per_pixel_hyps: bool = False, variant: int = 0, slots: int = 2):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("ViewPipeline runs on CUDA only (no CPU fallback)")
        v, c, h, w, d = num_views, channels, height, width, num_depth
        self.shape = (v, c, h, w, d)
        self.mode, self.groups, self.conf_mode, self.variant = mode, groups, conf_mode, variant
        cout = groups if mode == sweep.AGG_GROUP_CORR else (v - 1 if mode == sweep.AGG_PAIR_MEAN else c)
        hyp_shape = (d, h, w) if per_pixel_hyps else (d,)
        with torch.cuda.device(self.dev):
            self.copy_stream = torch.cuda.Stream()
            self.compute_stream = torch.cuda.Stream()
            self.slots = [{
                "feats": torch.empty((v, c, h, w), device=self.dev),
                "proj": torch.empty((v, 4, 4), device=self.dev),
                "hyps": torch.empty(hyp_shape, device=self.dev),
                "out": torch.empty((2, h, w), dtype=torch.float32).pin_memory(),
                "copied": torch.cuda.Event(), "free": torch.cuda.Event(), "done": torch.cuda.Event(),
            } for _ in range(slots)]
            self.texels = torch.empty((v, h, w, c), device=self.dev)
            self.volume = torch.empty((cout, d, h, w), device=self.dev)
        self._next = 0
        self._pending = collections.deque()
        self.h2d_bytes = 4 * (v * c * h * w + 16 * v + int(torch.tensor(hyp_shape).prod()))
        self.d2h_bytes = 8 * h * w

    def submit(self, feats: torch.Tensor, proj: torch.Tensor, hyps: torch.Tensor,
               logits_fn: Callable[[torch.Tensor], torch.Tensor]) -> None:
        """Enqueue one reference view.  feats [V,C,H,W], proj [V,4,4], hyps [D] or [D,H,W]: pinned host
        tensors (device tensors are accepted too and copied device-to-device)."""
        if len(self._pending) == len(self.slots):
            raise RuntimeError("all slots are in flight: collect() a result first")
        s = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["free"])            # the previous user of this slot has read its inputs
            s["feats"].copy_(feats, non_blocking=True)
            s["proj"].copy_(proj, non_blocking=True)
            s["hyps"].copy_(hyps, non_blocking=True)
            # camera geometry of this view (the reference's own torch calls, module.py:528 and :538) also runs on
            # the copy stream, i.e. behind the previous view's sweep; the slot keeps the tensors alive
            s["pose"] = sweep.relative_poses(s["proj"])
            s["rays"] = sweep.rays_for(s["pose"], self.shape[2], self.shape[3])
            s["copied"].record()
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(s["copied"])
            sweep.to_texels(s["feats"], out=self.texels)
            sweep.cost_volume(self.texels, s["pose"], s["hyps"], self.mode, groups=self.groups, out=self.volume,
                              variant=self.variant, rays=s["rays"])
            logits = logits_fn(self.volume)
            r = sweep.depth_regress(logits, s["hyps"], conf_mode=self.conf_mode, want_index=False)
            s["free"].record()                                # inputs consumed: the slot may be refilled
            s["out"][0].copy_(r["depth"], non_blocking=True)
            s["out"][1].copy_(r["conf"], non_blocking=True)
            s["done"].record()
        self._pending.append(s)

    def collect(self):
        """Block until the oldest submitted view is finished; returns (depth [H,W], conf [H,W]) as views of
        a pinned host buffer that stays valid until the slot is reused (`slots` submissions later)."""
        if not self._pending:
            raise RuntimeError("nothing submitted")
        s = self._pending.popleft()
        s["done"].synchronize()
        return s["out"][0], s["out"][1]

    def drain(self):
        out = []
        while self._pending:
            out.append(self.collect())
        return out
