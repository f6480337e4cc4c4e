"""Network-level parity on the B200, without the reference checkout: the engine is replayed between the CNN tensors
the LIVE reference produced when its three inference networks ran end to end (oracle/make_golden_networks.py ->
tests/golden/net_{cas,red,ada}.npz: Infer_CascadeMVSNet.forward cas_mvsnet.py:183-241, Infer_CascadeREDNet.forward
msrednet.py:473-528, Infer_AdaMVSNet.forward adamvs.py:565-617, seeded random weights, 64 x 64 images, V = 3).

  * per stage, with the reference's own stage inputs: every volume the engine hands a regulariser against the volume the
    reference handed it (<= 1e-4 norm-wise), depth (<= 1e-3 relative) and confidence of the stage;
  * chained through all three stages, the reference's stage glue restated below with the rebound module functions
    (what `install()` makes the reference's own forward call): final depth <= 1e-3, arg-max plane >= 99.9 %.

The regularisers are replayed from the record (their weights and classes live in the reference); everything else --
warp, aggregation, softmax / streaming regression, confidence, hypothesis resampling -- runs on the GPU through the C ABI.
"""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_norm_err
from deep3d_aerial_b200 import depthnets, module, sweep

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"
VOL_TOL, DEPTH_TOL = 1e-4, 1e-3


class Replay:
    """A regulariser that hands back the reference's recorded outputs, call by call, and keeps what it was given."""

    def __init__(self, outputs, n_states=0):
        self.outputs, self.n_states, self.seen = outputs, n_states, []

    def __call__(self, x, *states):
        k = len(self.seen)
        self.seen.append(x.detach().clone())
        out = self.outputs[k:k + 1].to(DEV)
        return (out,) + tuple(states) if self.n_states else out


def _depth_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())


def _volume_err(replay, g, key):
    """The volumes the engine handed the regulariser against the recorded ones (kept as every in_stride-th element)."""
    got = torch.stack([t[0] for t in replay.seen]).reshape(-1)[::int(g["in_stride"])].cpu()
    want = g[key]
    assert tuple(torch.stack([t[0] for t in replay.seen]).shape) == tuple(int(x) for x in g[key + "_shape"])
    return rel_norm_err(got, want)


def _stage_inputs(g, k):
    feats = [g["s%d_feats" % k][i:i + 1].to(DEV) for i in range(g["s%d_feats" % k].shape[0])]
    return feats, g["s%d_proj" % k].unsqueeze(0).to(DEV), g["s%d_hyps" % k].unsqueeze(0).to(DEV).contiguous()


def _run_stage(net, g, k, feats, proj, hyps, conf_in=None):
    d = hyps.shape[1]
    if net == "cas":
        rep = {"reg": Replay(g["s%d_reg_out" % k])}
        out = depthnets.cas_depthnet_forward(None, feats, proj, hyps, d, rep["reg"])
    elif net == "red":
        rep = {"reg": Replay(g["s%d_reg_out" % k], n_states=4)}
        out = depthnets.red_infer_forward(None, feats, proj, hyps, d, rep["reg"])
    else:
        rep = {"fuse": Replay(g["s%d_fuse_out" % k], n_states=2)}
        if conf_in is None:
            rep["reg"] = Replay(g["s%d_reg_out" % k])
        owner = types.SimpleNamespace(reg=rep.get("reg"), reg_fuse=rep["fuse"], in_up=bool(g["s%d_in_up" % k]))
        out = depthnets.ada_infer_forward(owner, feats, proj, hyps, d, confidence_map=conf_in)
    return out, rep


@pytest.fixture(params=[False, True], ids=["dense-block", "texel-pool"])
def texel_pool(request):
    """Both ways a DepthNet hands its views to the sweeps: a dense [V,H,W,C] block laid out per call, or slots of a per-image
    texel pool (depthnets.TEXEL_POOL, what predict.py --feature_cache installs)."""
    from deep3d_aerial_b200.texel_pool import TexelPool
    depthnets.TEXEL_POOL = TexelPool(8) if request.param else None
    yield depthnets.TEXEL_POOL
    depthnets.TEXEL_POOL = None


@pytest.mark.parametrize("net", ["cas", "red", "ada"])
def test_every_stage_on_the_reference_s_own_stage_inputs(net, texel_pool):
    g = load_golden("net_" + net)
    for k in (1, 2, 3):
        feats, proj, hyps = _stage_inputs(g, k)
        conf_in = None
        if net == "ada" and k > 1:
            conf_in = [c.unsqueeze(0).unsqueeze(0).to(DEV) for c in g["s%d_conf_in" % k]]
        out, rep = _run_stage(net, g, k, feats, proj, hyps, conf_in)
        for name, r in rep.items():
            err = _volume_err(r, g, "s%d_%s_in" % (k, name))
            print("%s stage %d: %s volume rel err %.2e" % (net, k, name, err))
            assert err < VOL_TOL
        derr = _depth_err(out["depth"][0], g["s%d_depth" % k])
        cerr = float((out["photometric_confidence"][0].cpu() - g["s%d_conf" % k]).abs().max())
        print("%s stage %d: depth rel err %.2e, confidence abs err %.2e" % (net, k, derr, cerr))
        assert derr < DEPTH_TOL and cerr < 1e-4
        if net == "ada":
            first = torch.stack([c[0, 0] for c in out["pair_confidence"][:len(feats) - 1]]).cpu()
            assert int(g["s%d_pair_conf_out_len" % k]) == len(out["pair_confidence"])
            assert float((first - g["s%d_pair_conf_out_first" % k]).abs().max()) < 1e-5
    if texel_pool is not None:
        assert texel_pool.stats()["misses"] > 0 and len(texel_pool.pools) == 3          # the sweeps did go through the pools


@pytest.mark.parametrize("net", ["cas", "red", "ada"])
def test_three_stages_chained_through_the_reference_s_stage_glue(net):
    """The loop of the reference's Infer*.forward, restated with the names install() rebinds (module.get_depth_range_samples
    and the DepthNet forwards): each stage's hypotheses come from the ENGINE's depth of the stage before."""
    g = load_golden("net_" + net)
    img_h, img_w = (int(x) for x in g["img_hw"])
    depth_values = g["depth_values"].unsqueeze(0).to(DEV)
    interval = (float(depth_values[0, -1]) - float(depth_values[0, 0])) / int(g["num_depth"])
    depth, pair_conf = None, None
    for k in (1, 2, 3):
        feats, proj, _ = _stage_inputs(g, k)
        nd, ratio = int(g["ndepths"][k - 1]), int(g["ratios"][k - 1])
        h, w = feats[0].shape[2:]
        if net == "ada":                                                                 # adamvs.py:590-608
            cur = depth if depth is not None else depth_values
            shape = [1, cur.shape[1], cur.shape[2]] if depth is not None else [1, h, w]
            hyps = module.get_depth_range_samples(cur_depth=cur, ndepth=nd, depth_inteval_pixel=ratio * interval,
                                                  device=DEV, dtype=torch.float32, shape=shape)
        else:                                                                            # cas_mvsnet.py:206-226
            cur = depth_values if depth is None else F.interpolate(depth.unsqueeze(1), [img_h, img_w], mode="bilinear",
                                                                   align_corners=False).squeeze(1)
            samples = module.get_depth_range_samples(cur_depth=cur, ndepth=nd, depth_inteval_pixel=ratio * interval,
                                                     device=DEV, dtype=torch.float32, shape=[1, img_h, img_w])
            hyps = F.interpolate(samples.unsqueeze(1), [nd, h, w], mode="trilinear", align_corners=False).squeeze(1)
        assert _depth_err(hyps[0], g["s%d_hyps" % k]) < DEPTH_TOL
        out, _ = _run_stage(net, g, k, feats, proj, hyps.contiguous(), pair_conf)
        depth = out["depth"]
        pair_conf = out.get("pair_confidence")
        assert _depth_err(depth[0], g["s%d_depth" % k]) < DEPTH_TOL
    derr = _depth_err(depth[0], g["final_depth"])
    cerr = float((out["photometric_confidence"][0].cpu() - g["final_conf"]).abs().max())
    print("%s chained: final depth rel err %.2e, confidence abs err %.2e" % (net, derr, cerr))
    assert derr < DEPTH_TOL and cerr < 1e-3


def test_argmax_plane_of_the_last_stage_agrees_with_the_reference():
    """Cas-MVSNet's last stage: the plane of maximum probability (softmax of the reference's regulariser output) as the
    regression kernel reports it (`index`) against torch's arg-max: >= 99.9 % of the pixels."""
    g = load_golden("net_cas")
    logits = g["s3_reg_out"][0, 0].to(DEV)
    hyps = g["s3_hyps"].to(DEV)
    got = sweep.depth_regress(logits, hyps, conf_mode=sweep.CONF_MAX_PROB)["index"].cpu().long()
    want = torch.softmax(g["s3_reg_out"][0, 0], 0).argmax(0)
    assert float((got == want).float().mean()) >= 0.999
