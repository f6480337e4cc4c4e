"""ORACLE — TEST INFRASTRUCTURE ONLY.

Generates the golden vectors under `tests/golden/` by running the *live* reference
(`/root/reference/mvs/mvs_cas/models/*.py`, imported, never copied) on small seeded inputs on CPU.
Run here, in the authoring container:  `python -m oracle.make_golden`.
The GPU box has no `/root/reference`; it only reads the committed `.npz` files.

Every case stores its inputs next to the reference outputs so the tests are self-contained.
Shapes are deliberately tiny (the whole directory is < 1 MB).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from deep3d_aerial_b200 import synth  # noqa: E402
from oracle import ref_live, standins  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _scene(v, c, h, w, seed, smooth=False, shifted=False):
    rig = synth.tiny_rig(num_views=v, width=w * 4, height=h * 4)
    proj = torch.from_numpy(rig.proj(4)).unsqueeze(0)            # [1,V,4,4]
    if shifted:  # push source views sideways so many samples fall outside the image
        proj = proj.clone()
        proj[:, 1:, 0, 3] += 9.0 * rig.z_mean
    feats = synth.make_features(v, c, h, w, seed=seed, smooth=smooth)
    return rig, proj, feats


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_live.load()
    torch.manual_seed(0)
    torch.set_grad_enabled(False)

    # ---- 1-3: homo_warping_float, uniform / per-pixel / mostly-out-of-bounds (module.py:516-557)
    for name, kw in (("warp_uniform", dict()), ("warp_perpixel", dict(perpixel=True)),
                     ("warp_oob", dict(shifted=True))):
        v, c, d, h, w = 3, 4, 5, 20, 24
        rig, proj, feats = _scene(v, c, h, w, seed=1, shifted=kw.get("shifted", False))
        if kw.get("perpixel"):
            cur = synth.smooth_depth_map(rig, h, w, seed=2)
            hyps = synth.per_pixel_hypotheses(cur, d, 0.35).unsqueeze(0)
        else:
            hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d).unsqueeze(0)
        outs = [ref.module.homo_warping_float(feats[i:i + 1], proj[:, i], proj[:, 0], hyps)
                for i in range(1, v)]
        np.savez(os.path.join(OUT, name + ".npz"), feats=_np(feats), proj=_np(proj), hyps=_np(hyps),
                 warped=_np(torch.stack(outs, 1)))

    # ---- 4: cas_mvsnet.DepthNet (variance + softmax + window-4 confidence), cas_mvsnet.py:35-78
    for name, perpixel in (("cas_depthnet_uniform", False), ("cas_depthnet_perpixel", True)):
        v, c, d, h, w = 4, 8, 12, 16, 24
        rig, proj, feats = _scene(v, c, h, w, seed=3, smooth=True)
        if perpixel:
            cur = synth.smooth_depth_map(rig, h, w, seed=1)
            hyps = synth.per_pixel_hypotheses(cur, d, 0.2).unsqueeze(0)
        else:
            hyps = ref.module.get_depth_range_samples(torch.tensor([[rig.dmin, rig.dmax]]), d, 0.0, "cpu",
                                                      torch.float32, [1, h, w])
        net = ref.cas_mvsnet.DepthNet().eval()
        cap = standins.Capture(standins.reg3d)
        out = net([feats[i:i + 1] for i in range(v)], proj, hyps, d, cap)
        np.savez(os.path.join(OUT, name + ".npz"), feats=_np(feats), proj=_np(proj), hyps=_np(hyps),
                 variance=_np(cap.seen[0]), depth=_np(out["depth"]),
                 conf=_np(out["photometric_confidence"]))

    # ---- 5: msrednet.InferDepthNet (slice-mode variance + streaming soft-argmax), msrednet.py:377-438
    v, c, d, h, w = 3, 8, 6, 16, 16
    rig, proj, feats = _scene(v, c, h, w, seed=4, smooth=True)
    hyps = ref.module.get_depth_range_samples(torch.tensor([[rig.dmin, rig.dmax]]), d, 0.0, "cpu",
                                              torch.float32, [1, h, w])
    cap = standins.Capture(standins.slice_reg_red)
    out = ref.msrednet.InferDepthNet().eval()([feats[i:i + 1] for i in range(v)], proj, hyps, d, cap)
    np.savez(os.path.join(OUT, "red_infer_depthnet.npz"), feats=_np(feats), proj=_np(proj), hyps=_np(hyps),
             variance_slices=_np(torch.stack(cap.seen, 2)), depth=_np(out["depth"]),
             conf=_np(out["photometric_confidence"]))

    # ---- 6: adamvs.InferDepthNet stage 1 (pair volumes -> view weights -> weighted product, in_up)
    #         and stage 2 fed with the stage-1 pair_confidence list (adamvs.py:436-531)
    v, c, d, h, w = 4, 8, 6, 12, 16
    rig, proj, feats = _scene(v, c, h, w, seed=5, smooth=True)
    hyps = ref.module.get_depth_range_samples(torch.tensor([[rig.dmin, rig.dmax]]), d, 0.0, "cpu",
                                              torch.float32, [1, h, w])
    net = ref.adamvs.InferDepthNet(in_depths=d, in_channels=c, in_up=True).eval()
    del net.reg, net.reg_fuse
    cap_pair = standins.Capture(standins.reg2d_pair)
    cap_fuse = standins.Capture(standins.slice_reg_up)
    net.reg, net.reg_fuse = cap_pair, cap_fuse
    out1 = net([feats[i:i + 1] for i in range(v)], proj, hyps, d, None)
    save = dict(feats=_np(feats), proj=_np(proj), hyps=_np(hyps),
                pair_volumes=_np(torch.stack(cap_pair.seen, 1)),
                similarity_slices=_np(torch.stack(cap_fuse.seen, 2)),
                depth=_np(out1["depth"]), conf=_np(out1["photometric_confidence"]),
                pair_result=_np(torch.stack(out1["pair_result"], 1)),
                n_pair_confidence=np.int64(len(out1["pair_confidence"])),
                pair_conf_head=_np(torch.stack(out1["pair_confidence"][:v - 1], 1)))
    # stage 2: features at 2x the resolution, per-pixel hypotheses around the stage-1 depth
    h2, w2, c2, d2 = 2 * h, 2 * w, 4, 4
    rig2 = synth.tiny_rig(num_views=v, width=w * 4, height=h * 4)
    proj2 = torch.from_numpy(rig2.proj(2)).unsqueeze(0)
    feats2 = synth.make_features(v, c2, h2, w2, seed=6, smooth=True)
    hyps2 = ref.module.get_depth_range_samples(out1["depth"], d2, 0.3, "cpu", torch.float32, [1, h2, w2])
    net2 = ref.adamvs.InferDepthNet(in_depths=d, in_channels=c2, in_up=False).eval()
    del net2.reg, net2.reg_fuse
    cap_fuse2 = standins.Capture(standins.slice_reg_same)
    net2.reg, net2.reg_fuse = standins.reg2d_pair, cap_fuse2
    out2 = net2([feats2[i:i + 1] for i in range(v)], proj2, hyps2, d2, out1["pair_confidence"])
    save.update(feats2=_np(feats2), proj2=_np(proj2), hyps2=_np(hyps2),
                similarity_slices2=_np(torch.stack(cap_fuse2.seen, 2)),
                depth2=_np(out2["depth"]), conf2=_np(out2["photometric_confidence"]),
                n_pair_confidence2=np.int64(len(out2["pair_confidence"])))
    np.savez(os.path.join(OUT, "ada_infer_depthnet.npz"), **save)

    # ---- 7: adamvs.DepthNet train-form weighted product (eps in numerator), adamvs.py:247-312
    net = ref.adamvs.DepthNet(in_depths=d, in_channels=c, in_up=False).eval()
    del net.reg, net.reg_fuse
    cap_fuse = standins.Capture(lambda vol: -3.0 * vol.mean(1))
    net.reg, net.reg_fuse = standins.reg2d_pair, cap_fuse
    out = net([feats[i:i + 1] for i in range(v)], proj, hyps, d, None)
    np.savez(os.path.join(OUT, "ada_train_depthnet.npz"), feats=_np(feats), proj=_np(proj), hyps=_np(hyps),
             fused=_np(cap_fuse.seen[0]), depth=_np(out["depth"]), conf=_np(out["photometric_confidence"]),
             pair_conf=_np(torch.stack(out["pair_confidence"], 1)))

    # ---- 8: depth_regression with 4-D hypotheses resized (module.py:605-613) + range samples
    g = torch.Generator().manual_seed(7)
    logits = synth.planted_logits(10, 12, 16, seed=7).unsqueeze(0)
    prob = torch.softmax(logits, 1)
    hy_small = 400 + 50 * torch.rand(1, 10, 6, 8, generator=g)
    cur = 450 + 30 * torch.rand(1, 12, 16, generator=g)
    np.savez(os.path.join(OUT, "regress_misc.npz"), logits=_np(logits), hy_small=_np(hy_small), cur=_np(cur),
             depth_resized=_np(ref.module.depth_regression(prob, hy_small)),
             samples_range=_np(ref.module.get_depth_range_samples(torch.tensor([[400.0, 600.0]]), 10, 0.0, "cpu",
                                                                   torch.float32, [1, 12, 16])),
             samples_cur=_np(ref.module.get_depth_range_samples(cur, 8, 0.52, "cpu", torch.float32, [1, 12, 16])))

    # ---- 9: cascade stage glue captured from Infer_CascadeMVSNet.forward (cas_mvsnet.py:183-241)
    model = ref.cas_mvsnet.Infer_CascadeMVSNet(num_depth=48, ndepths=[6, 4, 2], depth_intervals_ratio=[4, 2, 1]).eval()
    fh, fw = 32, 48
    del model.feature, model.DepthNet
    model.feature = lambda img: {"stage1": torch.zeros(1, 2, fh // 4, fw // 4), "stage2": torch.zeros(1, 2, fh // 2, fw // 2),
                                 "stage3": torch.zeros(1, 2, fh, fw)}
    seen = []
    gen = torch.Generator().manual_seed(8)

    def fake_depthnet(features, proj, depth_values, num_depth, cost_regularization):
        seen.append(depth_values.clone())
        hh, ww = depth_values.shape[2:]
        dep = 450 + 100 * torch.rand(1, hh, ww, generator=gen)
        seen.append(dep.clone())
        return {"depth": dep, "photometric_confidence": torch.zeros(1, hh, ww)}

    model.DepthNet = fake_depthnet
    pm = {"stage%d" % s: torch.eye(4).view(1, 1, 4, 4).repeat(1, 2, 1, 1) for s in (1, 2, 3)}
    model(torch.zeros(1, 2, 3, fh, fw), pm, torch.tensor([[400.0, 600.0]]))
    np.savez(os.path.join(OUT, "cas_stage_glue.npz"), dv1=_np(seen[0]), depth1=_np(seen[1]), dv2=_np(seen[2]),
             depth2=_np(seen[3]), dv3=_np(seen[4]), full_hw=np.array([fh, fw]), dmin=400.0, dmax=600.0,
             num_depth=48, ndepths=np.array([6, 4, 2]), ratios=np.array([4, 2, 1]))

    # ---- 10: ucsnet.compute_depth (variance + window-4 conf + exp_variance), ucsnet.py:99-151
    v, c, d, h, w = 3, 4, 8, 12, 12
    rig, proj, feats = _scene(v, c, h, w, seed=9, smooth=True)
    cur = synth.smooth_depth_map(rig, h, w, seed=3)
    hyps = ref.ucsnet.uncertainty_aware_samples(cur.view(1, 1, h, w), torch.full((1, 1, h, w), 1.3), d, "cpu",
                                                torch.float32, [1, h, w])
    cap = standins.Capture(standins.reg3d)
    out = ref.ucsnet.compute_depth([feats[i:i + 1] for i in range(v)], proj, hyps, cap, 1.5, False)
    np.savez(os.path.join(OUT, "ucs_compute_depth.npz"), feats=_np(feats), proj=_np(proj), hyps=_np(hyps),
             cur=_np(cur), variance=_np(cap.seen[0]), depth=_np(out["depth"]),
             conf=_np(out["photometric_confidence"]), exp_variance=_np(out["variance"]))

    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden vectors written to", OUT, "total bytes", total)


if __name__ == "__main__":
    main()
