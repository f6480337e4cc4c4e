"""CUDA-graph capture of the plane-at-a-time part of AdaMVS / RED-Net inference.

After the fused sweep has written the whole cost volume in plane-major layout, what remains of a reference
view is D iterations of "recurrent regulariser on one [C,h,w] slice -> streaming soft-argmax update"
(adamvs.py:492-529, msrednet.py:400-437).  The regulariser is the caller's PyTorch module (a few dozen small
convolution / element-wise launches per plane), the update is one libd3dsweep launch; at D = 48 + 32 + 8 planes
per view the loop is bound by launch latency, not by the GPU.  `PlaneLoop` records the whole loop once -- every
launch of every plane, on static buffers -- and replays it per reference view:

    loop = PlaneLoop(step, batch=1, planes=48, channels=32, hw=(688, 464), out_hw=(1376, 928),
                     state_shapes=[(1, 8, 688, 464), (1, 16, 344, 232)], device="cuda:0")
    sweep.cost_volume(texels, pose, hyps, AGG_WEIGHTED_PRODUCT, plane_major=True, weights=w, out=loop.volume[0])
    loop.hyps[0].copy_(hyps)
    depth, conf = loop.replay()          # static tensors: clone() what must outlive the next replay

`step(slice [B,C,h,w], *states) -> (reg_cost [B,1,H,W], *new_states)` must be capture-safe: fixed shapes, no
host synchronisation, no `.cuda()` of host tensors inside (the reference's regularisers satisfy this; the
recurrent states are created here, on the device).  libd3dsweep launches are capture-safe by contract
(include/d3d_sweep.h).  CUDA only.
"""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import torch

from . import sweep


class PlaneLoop:
    def __init__(self, step: Callable, batch: int, planes: int, channels: int, hw: Tuple[int, int],
                 out_hw: Tuple[int, int], state_shapes: Sequence[Tuple[int, ...]], device,
                 hyps_hw: Tuple[int, int] = None, warmup: int = 2):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("PlaneLoop runs on CUDA only (no CPU fallback)")
        self.step, self.batch, self.planes = step, batch, planes
        h, w = hw
        hh, hw_ = hyps_hw if hyps_hw is not None else hw
        self.out_hw = tuple(out_hw)
        self.state_shapes = [tuple(s) for s in state_shapes]
        with torch.cuda.device(self.dev):
            self.volume = torch.empty((batch, planes, channels, h, w), device=self.dev)      # plane-major, static
            self.hyps = torch.empty((batch, planes, hh, hw_), device=self.dev)               # static
            self._acc = [torch.zeros((3,) + self.out_hw, device=self.dev) for _ in range(batch)]
            self.depth = self.conf = None
            self.graph = None
            self.volume.zero_()
            self.hyps.fill_(1.0)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up off the capture stream (allocator, cuDNN plans)
                for _ in range(max(1, warmup)):
                    self._run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.depth, self.conf = self._run()
            self.graph = g

    def _run(self):
        states = [torch.zeros(s, device=self.dev) for s in self.state_shapes]          # adamvs.py:451-452
        from .depthnets import STREAM_BATCH_PLANES
        k = max(1, min(int(STREAM_BATCH_PLANES), 16))          # planes folded into the accumulators per launch
        depth = conf = None
        held = [[] for _ in range(self.batch)]
        for d in range(self.planes):
            out = self.step(self.volume[:, d], *states)
            reg_cost, states = out[0], list(out[1:])
            last = d == self.planes - 1
            for b in range(self.batch):
                held[b].append(reg_cost[b, 0])
            if len(held[0]) < k and not last:
                continue
            res = [sweep.depth_regress(list(held[b]), self.hyps[b], softmax_mode=sweep.SOFTMAX_RAW_EXP,
                                       d_begin=d + 1 - len(held[b]), num_depth=self.planes, state=self._acc[b],
                                       finalize=last)
                   for b in range(self.batch)]
            for b in range(self.batch):
                held[b].clear()
            if last:
                depth = torch.stack([r["depth"] for r in res], 0)
                conf = torch.stack([r["conf"] for r in res], 0)
        return depth, conf

    def replay(self):
        """Run the recorded loop on the current contents of `volume` and `hyps`; returns (depth, conf), both
        [B,H,W] static tensors that the next replay overwrites."""
        self.graph.replay()
        return self.depth, self.conf
