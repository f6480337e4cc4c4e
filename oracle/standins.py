"""ORACLE — TEST INFRASTRUCTURE ONLY.

Parameter-free, deterministic stand-ins for the CNN regularisers (out of scope: they stay in
PyTorch unchanged on both sides).  The reference's DepthNets take the regulariser as a callable /
sub-module, so the golden generator plugs these in to (a) capture the cost volumes the reference
builds and (b) drive its softmax / regression code with known logits.  The same functions are
handed to our drop-in DepthNets in the tests, so both sides see identical logits.
"""
import torch
import torch.nn.functional as F


class Capture:
    """Wraps a callable and records every first positional input it is called with."""

    def __init__(self, fn):
        self.fn = fn
        self.seen = []

    def __call__(self, x, *rest):
        self.seen.append(x.detach().clone())
        return self.fn(x, *rest)


def reg3d(volume):
    """[B,C,D,H,W] -> [B,1,D,H,W] logits (stand-in for CostRegNet, cas_mvsnet.py:81-121)."""
    return -4.0 * volume.mean(1, keepdim=True)


def reg2d_pair(pair_volume):
    """[B,D,h,w] -> [B,D,h,w] scores (stand-in for CostRegNet2D, adamvs.py:198-238)."""
    return 6.0 * pair_volume


def slice_reg_up(similarity, state1, state2):
    """[B,C,h,w] -> [B,1,2h,2w] (stand-in for SliceCostRegNetRED with up=True, adamvs.py:403-427)."""
    logit = F.interpolate(1.5 * similarity.mean(1, keepdim=True), scale_factor=2, mode="nearest")
    return logit, state1, state2


def slice_reg_same(similarity, state1, state2):
    """[B,C,h,w] -> [B,1,h,w] (stand-in for SliceCostRegNetRED with up=False)."""
    return 1.5 * similarity.mean(1, keepdim=True), state1, state2


def slice_reg_red(variance, s1, s2, s3, s4):
    """[B,C,h,w] -> [B,1,h,w] (stand-in for slice_RED_Regularization, msrednet.py:337-370)."""
    return -2.0 * variance.mean(1, keepdim=True), s1, s2, s3, s4
