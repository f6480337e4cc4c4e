"""Pins the oracle: torch restatement == live-reference golden vectors (bit for bit, CPU ATen),
fp64 arbiter agrees within fp32 noise, and -- when /root/reference is present -- the
restatement equals the live reference on fresh inputs."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_norm_err
from deep3d_aerial_b200 import synth
from oracle import ref_live, standins, sweep_np64, sweep_torch

torch.set_grad_enabled(False)


def _views(feats):
    return [feats[i:i + 1] for i in range(feats.shape[0])]


@pytest.mark.parametrize("name", ["warp_uniform", "warp_perpixel", "warp_oob"])
def test_warp_matches_golden(name):
    g = load_golden(name)
    v = g["feats"].shape[0]
    for i in range(1, v):
        got = sweep_torch.warp_source(g["feats"][i:i + 1], g["proj"][:, i], g["proj"][:, 0], g["hyps"])
        assert torch.equal(got, g["warped"][:, i - 1])
        arb = sweep_np64.warp_source(g["feats"][i].numpy(), g["proj"][0, i].numpy(), g["proj"][0, 0].numpy(),
                                     g["hyps"][0].numpy())
        assert rel_norm_err(arb, g["warped"][0, i - 1]) < 2e-4


@pytest.mark.parametrize("name", ["warp_double_uniform", "warp_double_perpixel"])
def test_warp_double_matches_golden(name):
    """homo_warping_double (module.py:560-601, fp64 coordinates) restated: bit for bit against the live reference's output."""
    g = load_golden(name)
    for i in range(1, g["feats"].shape[0]):
        got = sweep_torch.warp_source_double(g["feats"][i:i + 1], g["proj"][:, i], g["proj"][:, 0], g["hyps"])
        assert torch.equal(got, g["warped"][:, i - 1])
    with pytest.raises(RuntimeError):       # fp32 projections: torch.matmul does not promote (upstream fails the same way)
        sweep_torch.warp_source_double(g["feats"][1:2], g["proj"][:, 1].float(), g["proj"][:, 0].float(), g["hyps"])


@pytest.mark.parametrize("name", ["cas_depthnet_uniform", "cas_depthnet_perpixel"])
def test_variance_and_window4_match_golden(name):
    g = load_golden(name)
    var = sweep_torch.variance_volume(_views(g["feats"]), g["proj"], g["hyps"])
    assert torch.equal(var, g["variance"])
    depth, conf, _ = sweep_torch.regress_window4(standins.reg3d(var).squeeze(1), g["hyps"])
    assert torch.equal(depth, g["depth"])
    assert torch.equal(conf, g["conf"])
    arb = sweep_np64.variance_volume(g["feats"].numpy(), g["proj"][0].numpy(), g["hyps"][0].numpy())
    assert rel_norm_err(arb, g["variance"][0]) < 1e-4


def test_red_slice_variance_and_streaming_match_golden():
    g = load_golden("red_infer_depthnet")
    d = g["hyps"].shape[1]
    slices = [sweep_torch.variance_volume(_views(g["feats"]), g["proj"], g["hyps"][:, k:k + 1]) for k in range(d)]
    var = torch.cat(slices, 2)
    assert torch.equal(var, g["variance_slices"])
    logits = [standins.slice_reg_red(s.squeeze(2), 0, 0, 0, 0)[0] for s in slices]
    depth, conf = sweep_torch.regress_streaming(logits, [g["hyps"][:, k:k + 1] for k in range(d)])
    assert torch.equal(depth, g["depth"])
    assert torch.equal(conf, g["conf"])


def test_adamvs_infer_matches_golden():
    g = load_golden("ada_infer_depthnet")
    feats, proj, hyps = _views(g["feats"]), g["proj"], g["hyps"]
    d = hyps.shape[1]
    pairs = sweep_torch.pair_mean_volumes(feats, proj, hyps)
    assert torch.equal(torch.stack(pairs, 1), g["pair_volumes"])
    weights, results = [], []
    for pv in pairs:
        dep, conf, _ = sweep_torch.regress_maxprob(standins.reg2d_pair(pv), hyps)
        weights.append(conf.unsqueeze(1))
        results.append(dep)
    assert torch.equal(torch.stack(results, 1), g["pair_result"])
    assert torch.equal(torch.stack(weights, 1), g["pair_conf_head"])
    sim = sweep_torch.weighted_product_volume(feats, proj, hyps, weights)
    assert torch.equal(sim, g["similarity_slices"])
    logits = [standins.slice_reg_up(sim[:, :, k], 0, 0)[0] for k in range(d)]
    depth, conf = sweep_torch.regress_streaming(logits, [hyps[:, k:k + 1] for k in range(d)], upsample2=True)
    assert torch.equal(depth, g["depth"])
    assert torch.equal(conf, g["conf"])
    # stage 2 consumes the first V-1 entries of the (quirkily over-long) pair_confidence list
    assert g["n_pair_confidence"] == (len(feats) - 1) * (1 + d)
    feats2 = _views(g["feats2"])
    sim2 = sweep_torch.weighted_product_volume(feats2, g["proj2"], g["hyps2"], weights)
    assert torch.equal(sim2, g["similarity_slices2"])
    d2 = g["hyps2"].shape[1]
    logits2 = [standins.slice_reg_same(sim2[:, :, k], 0, 0)[0] for k in range(d2)]
    depth2, conf2 = sweep_torch.regress_streaming(logits2, [g["hyps2"][:, k:k + 1] for k in range(d2)])
    assert torch.equal(depth2, g["depth2"])
    assert torch.equal(conf2, g["conf2"])
    arb = sweep_np64.pair_mean_volumes(g["feats"].numpy(), proj[0].numpy(), hyps[0].numpy())
    assert rel_norm_err(arb, g["pair_volumes"][0]) < 1e-4


def test_adamvs_train_form_matches_golden():
    g = load_golden("ada_train_depthnet")
    feats, proj, hyps = _views(g["feats"]), g["proj"], g["hyps"]
    weights = [g["pair_conf"][:, i] for i in range(len(feats) - 1)]
    fused = sweep_torch.weighted_product_volume(feats, proj, hyps, weights, eps_in_numerator=True)
    # the train form accumulates whole volumes rather than slices: same values up to summation order
    assert rel_norm_err(fused, g["fused"]) < 1e-6
    depth, conf, _ = sweep_torch.regress_maxprob(-3.0 * g["fused"].mean(1), hyps)
    assert torch.equal(depth, g["depth"])
    assert torch.equal(conf, g["conf"])


def test_regress_misc_matches_golden():
    g = load_golden("regress_misc")
    prob = torch.softmax(g["logits"], 1)
    assert torch.equal(sweep_torch.expectation(prob, g["hy_small"]), g["depth_resized"])
    assert torch.equal(sweep_torch.depth_range_samples(torch.tensor([[400.0, 600.0]]), 10, 0.0, [1, 12, 16]),
                       g["samples_range"])
    assert torch.equal(sweep_torch.depth_range_samples(g["cur"], 8, 0.52, [1, 12, 16]), g["samples_cur"])


def test_cascade_stage_glue_matches_golden():
    g = load_golden("cas_stage_glue")
    fh, fw = [int(x) for x in g["full_hw"]]
    interval = (g["dmax"] - g["dmin"]) / g["num_depth"]
    nd = [int(x) for x in g["ndepths"]]
    ratios = [int(x) for x in g["ratios"]]
    rng = torch.tensor([[g["dmin"], g["dmax"]]], dtype=torch.float32)
    assert torch.equal(sweep_torch.cascade_stage_hypotheses(rng, nd[0], ratios[0] * interval, (fh, fw), 4), g["dv1"])
    assert torch.equal(sweep_torch.cascade_stage_hypotheses(g["depth1"], nd[1], ratios[1] * interval, (fh, fw), 2), g["dv2"])
    assert torch.equal(sweep_torch.cascade_stage_hypotheses(g["depth2"], nd[2], ratios[2] * interval, (fh, fw), 1), g["dv3"])


def test_ucs_matches_golden():
    g = load_golden("ucs_compute_depth")
    h, w = g["cur"].shape
    hyps = sweep_torch.uncertainty_samples(g["cur"].view(1, 1, h, w), torch.full((1, 1, h, w), 1.3), 8)
    assert torch.equal(hyps, g["hyps"])
    var = sweep_torch.variance_volume(_views(g["feats"]), g["proj"], g["hyps"])
    assert torch.equal(var, g["variance"])
    logits = standins.reg3d(var).squeeze(1)
    depth, conf, _ = sweep_torch.regress_window4(logits, g["hyps"])
    assert torch.equal(depth, g["depth"]) and torch.equal(conf, g["conf"])
    ev = sweep_torch.exp_variance(torch.softmax(logits, 1), g["hyps"], depth, 1.5)
    assert torch.equal(ev, g["exp_variance"])


def test_gwc_g1_equals_pair_mean():
    """G=1 group-wise correlation is the reference's pair volume (adamvs.py:473) averaged over views."""
    g = load_golden("ada_infer_depthnet")
    feats, proj, hyps = _views(g["feats"]), g["proj"], g["hyps"]
    gwc = sweep_torch.groupwise_correlation_volume(feats, proj, hyps, 1)
    assert rel_norm_err(gwc[:, 0], g["pair_volumes"].mean(1)) < 1e-6
    arb = sweep_np64.groupwise_volume(g["feats"].numpy(), proj[0].numpy(), hyps[0].numpy(), 4)
    got = sweep_torch.groupwise_correlation_volume(feats, proj, hyps, 4)
    assert rel_norm_err(got[0], arb) < 1e-4


@pytest.mark.skipif(not ref_live.available(), reason="live reference only exists in the authoring container")
def test_restatement_equals_live_reference_fresh_inputs():
    ref = ref_live.load()
    rig = synth.tiny_rig(num_views=3, width=160, height=128)
    proj = torch.from_numpy(rig.proj(4)).unsqueeze(0)
    feats = synth.make_features(3, 8, 32, 40, seed=11)
    hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, 16).unsqueeze(0)
    for i in (1, 2):
        a = ref.module.homo_warping_float(feats[i:i + 1], proj[:, i], proj[:, 0], hyps)
        b = sweep_torch.warp_source(feats[i:i + 1], proj[:, i], proj[:, 0], hyps)
        assert torch.equal(a, b)
    cap = standins.Capture(standins.reg3d)
    hy4 = ref.module.get_depth_range_samples(torch.tensor([[rig.dmin, rig.dmax]]), 16, 0.0, "cpu", torch.float32, [1, 32, 40])
    out = ref.cas_mvsnet.DepthNet().eval()(_views(feats), proj, hy4, 16, cap)
    var = sweep_torch.variance_volume(_views(feats), proj, hy4)
    assert torch.equal(var, cap.seen[0])
    depth, conf, _ = sweep_torch.regress_window4(standins.reg3d(var).squeeze(1), hy4)
    assert torch.equal(depth, out["depth"]) and torch.equal(conf, out["photometric_confidence"])


def test_trilinear_downsample_is_centre_pair_average():
    """cas_mvsnet.py:224-226 at scale 4 averages the two *centre* samples per axis, not a 4x4 box
    (align_corners=False maps output o to source 4o+1.5)."""
    g = torch.Generator().manual_seed(0)
    full = torch.rand(1, 1, 3, 8, 12, generator=g)
    dv = torch.nn.functional.interpolate(full, [3, 2, 3], mode="trilinear", align_corners=False)
    centre = full[..., 1::4, :][..., 1::4] + full[..., 1::4, :][..., 2::4] + full[..., 2::4, :][..., 1::4] + full[..., 2::4, :][..., 2::4]
    assert torch.allclose(dv, centre / 4, atol=1e-6)
