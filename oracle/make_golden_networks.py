"""ORACLE — TEST INFRASTRUCTURE ONLY.

Network-level golden vectors: runs the LIVE reference's three inference networks end to end on CPU --
`Infer_CascadeMVSNet.forward` (cas_mvsnet.py:183-241), `Infer_CascadeREDNet.forward` (msrednet.py:473-528) and
`Infer_AdaMVSNet.forward` (adamvs.py:565-617) -- with seeded random weights on a small seeded scene, and records
every tensor that crosses a DepthNet boundary:

    per stage: the feature maps FeatureNet produced, the projection matrices, the depth hypotheses the
               reference's stage glue built, every input and output of every regulariser call (whole volume
               for Cas-MVSNet, one call per plane for the recurrent RED / AdaMVS regularisers, one call per
               source view for AdaMVS's pair regulariser), the stage's depth / confidence (and AdaMVS's
               pair_confidence list as the next stage receives it)

so that `tests/test_networks_gpu.py` can replay the engine on the B200 (which has no reference checkout) between
the recorded CNN tensors: per stage with the reference's own inputs (volumes <= 1e-4, depth <= 1e-3), and
chained through all three stages with the stage glue restated in the test (final depth <= 1e-3).

    python -m oracle.make_golden_networks        # here, in the authoring container -> tests/golden/net_*.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from deep3d_aerial_b200 import synth  # noqa: E402
from oracle import ref_live  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
V, IMG_H, IMG_W, NUM_DEPTH = 3, 64, 64, 32
NDEPTHS, RATIOS = [8, 8, 8], [4, 2, 1]      # the 3-D U-Nets halve D, H and W three times
IN_STRIDE = 7      # regulariser INPUTS (the engine's volumes) are kept as every 7th element: they are 3 MB per network otherwise


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def scene():
    rig = synth.make_rig(num_views=V, width=IMG_W, height=IMG_H, focal=90.0, z_mean=10.0, baseline_frac=0.06,
                         range_frac=0.3)
    g = torch.Generator().manual_seed(2024)
    imgs = torch.randn(1, V, 3, IMG_H, IMG_W, generator=g)
    imgs = F.avg_pool2d(imgs.view(V, 3, IMG_H, IMG_W), 5, stride=1, padding=2).view(1, V, 3, IMG_H, IMG_W)
    imgs = imgs / imgs.std()
    proj = {"stage%d" % (i + 1): torch.from_numpy(rig.proj(s)).unsqueeze(0) for i, s in enumerate((4, 2, 1))}
    depth_values = torch.linspace(rig.dmin, rig.dmax, NUM_DEPTH).unsqueeze(0)
    return imgs, proj, depth_values


class Tape:
    """Records the inputs (first positional tensor) and the first output of every call of a regulariser."""

    def __init__(self, fn):
        self.fn, self.inputs, self.outputs = fn, [], []

    def __call__(self, x, *rest):
        out = self.fn(x, *rest)
        self.inputs.append(x.detach().clone())
        self.outputs.append((out[0] if isinstance(out, tuple) else out).detach().clone())
        return out


def record_stage(store, k, features, proj, dv, out, tapes):
    p = "s%d_" % k
    store[p + "feats"] = np.stack([_np(f[0]) for f in features])           # [V,C,h,w]
    store[p + "proj"] = _np(proj[0])                                       # [V,4,4]
    store[p + "hyps"] = _np(dv[0])                                         # [D,h,w]
    store[p + "depth"] = _np(out["depth"][0])
    store[p + "conf"] = _np(out["photometric_confidence"][0])
    for name, tape in tapes.items():
        if tape.inputs:
            store[p + name + "_in"] = np.stack([_np(t[0]) for t in tape.inputs]).reshape(-1)[::IN_STRIDE].copy()
            store[p + name + "_in_shape"] = np.array((len(tape.inputs),) + tuple(tape.inputs[0][0].shape))
            store[p + name + "_out"] = np.stack([_np(t[0]) for t in tape.outputs])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_live.load()
    torch.set_grad_enabled(False)
    imgs, proj, depth_values = scene()

    # ---------------------------------------------------------------- Cas-MVSNet and RED-Net: one DepthNet, called per stage
    for tag, make, depthnet_cls in (
            ("cas", lambda: ref.cas_mvsnet.Infer_CascadeMVSNet(num_depth=NUM_DEPTH, ndepths=NDEPTHS,
                                                               depth_intervals_ratio=RATIOS), ref.cas_mvsnet.DepthNet),
            ("red", lambda: ref.msrednet.Infer_CascadeREDNet(num_depth=NUM_DEPTH, ndepths=NDEPTHS,
                                                             depth_intervals_ratio=RATIOS), ref.msrednet.InferDepthNet)):
        torch.manual_seed(11)
        net = make().eval()
        store, stage = {}, [0]
        original = depthnet_cls.forward

        def spy(self, features, proj_matrices, depth_values, num_depth, cost_regularization, _orig=original,
                _store=store, _stage=stage):
            tape = Tape(cost_regularization)
            out = _orig(self, features, proj_matrices, depth_values, num_depth, tape)
            _stage[0] += 1
            record_stage(_store, _stage[0], features, proj_matrices, depth_values, out, {"reg": tape})
            return out

        depthnet_cls.forward = spy
        try:
            outputs = net(imgs, proj, depth_values)
        finally:
            depthnet_cls.forward = original
        store["depth_values"] = _np(depth_values[0])
        store["final_depth"] = _np(outputs["depth"][0])
        store["final_conf"] = _np(outputs["photometric_confidence"][0])
        store["ndepths"] = np.array(NDEPTHS)
        store["ratios"] = np.array(RATIOS)
        store["num_depth"] = np.array(NUM_DEPTH)
        store["img_hw"] = np.array([IMG_H, IMG_W])
        store["in_stride"] = np.array(IN_STRIDE)
        np.savez_compressed(os.path.join(OUT, "net_%s.npz" % tag), **store)
        print(tag, {k: v.shape for k, v in store.items() if k.startswith("s1_")})

    # ---------------------------------------------------------------- AdaMVS: one InferDepthNet per stage, own regularisers
    torch.manual_seed(11)
    net = ref.adamvs.Infer_AdaMVSNet(num_depth=NUM_DEPTH, ndepths=NDEPTHS, depth_intervals_ratio=RATIOS).eval()
    store = {}
    originals = []
    for k, dn in enumerate(net.DepthNet):
        tapes = {"reg": Tape(dn.reg.forward), "fuse": Tape(dn.reg_fuse.forward)}
        dn.reg.forward, dn.reg_fuse.forward = tapes["reg"], tapes["fuse"]
        orig = dn.forward

        def spy(features, proj_matrices, depth_values, num_depth, confidence_map=None, _orig=orig, _k=k + 1,
                _tapes=tapes, _dn=dn):
            given = None if confidence_map is None else [c.clone() for c in confidence_map]
            out = _orig(features, proj_matrices, depth_values=depth_values, num_depth=num_depth,
                        confidence_map=confidence_map)
            record_stage(store, _k, features, proj_matrices, depth_values, out, _tapes)
            n_src = len(features) - 1
            if given is not None:        # what the stage was handed: only the first V-1 maps are consumed (adamvs.py:498)
                store["s%d_conf_in" % _k] = np.stack([_np(c[0, 0]) for c in given[:n_src]])
            store["s%d_pair_conf_out_first" % _k] = np.stack([_np(c[0, 0]) for c in out["pair_confidence"][:n_src]])
            store["s%d_pair_conf_out_len" % _k] = np.array(len(out["pair_confidence"]))
            store["s%d_in_up" % _k] = np.array(int(_dn.in_up))
            return out

        dn.forward = spy
        originals.append((dn, orig))
    outputs = net(imgs, proj, depth_values)
    store["depth_values"] = _np(depth_values[0])
    store["final_depth"] = _np(outputs["depth"][0])
    store["final_conf"] = _np(outputs["photometric_confidence"][0])
    store["ndepths"] = np.array(NDEPTHS)
    store["ratios"] = np.array(RATIOS)
    store["num_depth"] = np.array(NUM_DEPTH)
    store["img_hw"] = np.array([IMG_H, IMG_W])
    store["in_stride"] = np.array(IN_STRIDE)
    np.savez_compressed(os.path.join(OUT, "net_ada.npz"), **store)
    print("ada", {k: v.shape for k, v in store.items() if k.startswith("s2_")})
    for f in ("net_cas.npz", "net_red.npz", "net_ada.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
