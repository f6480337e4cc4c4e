"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libd3dsweep.so), against
  (1) the committed golden vectors produced by the live reference (tests/golden, oracle/make_golden.py),
  (2) the oracle (oracle/sweep_torch.py) on the same seeded synthetic inputs, on CPU-ATen and on
      CUDA-ATen (the reference's own path on this GPU),
  (3) size-independent properties at BASELINE.json's full sizes.

Tolerances are north_star's: cost volumes <= 1e-4 relative (norm-wise: max|a-b| / max|b|, SURVEY.md §7
hard part 2), regressed depth <= 1e-3 relative, arg-max plane agreement >= 99.9 %.
"""
import types

import pytest
import torch

from conftest import load_golden, rel_norm_err
from deep3d_aerial_b200 import depthnets, module, sweep, synth
from oracle import standins, sweep_torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

VOL_TOL = 1e-4     # cost volume, norm-wise relative
DEPTH_TOL = 1e-3   # regressed depth, relative
DEV = "cuda"


def _views(feats):
    return [feats[i:i + 1] for i in range(feats.shape[0])]


def _cuda_views(feats):
    return [f.to(DEV) for f in _views(feats)]


def _depth_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())


def _scene(v, c, d, h, w, seed=0, smooth=False, perpixel=False, rig=None, scale=4):
    rig = rig or synth.tiny_rig(num_views=v, width=w * scale, height=h * scale)
    proj = torch.from_numpy(rig.proj(scale)).unsqueeze(0)
    feats = synth.make_features(v, c, h, w, seed=seed, smooth=smooth)
    if perpixel:
        cur = synth.smooth_depth_map(rig, h, w, seed=seed)
        hyps = synth.per_pixel_hypotheses(cur, d, (rig.dmax - rig.dmin) / (4 * d)).unsqueeze(0)
    else:
        hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d).unsqueeze(0)
    return rig, proj, feats, hyps


def _ours_volume(feats, proj, hyps, mode, **kw):
    """feats [V,C,H,W] cpu, proj [1,V,4,4], hyps [1,D] or [1,D,H,W] -> volume on cpu."""
    tex = sweep.to_texels(feats.to(DEV))
    pose = sweep.relative_poses(proj[0].to(DEV))
    return sweep.cost_volume(tex, pose, hyps[0].to(DEV).contiguous(), mode, **kw).cpu()


# ------------------------------------------------------------------ (1) golden vectors of the live reference
@pytest.mark.parametrize("name", ["warp_uniform", "warp_perpixel", "warp_oob"])
def test_homo_warping_matches_golden(name):
    g = load_golden(name)
    v = g["feats"].shape[0]
    for i in range(1, v):
        got = module.homo_warping_float(g["feats"][i:i + 1].to(DEV), g["proj"][:, i].to(DEV), g["proj"][:, 0].to(DEV),
                                        g["hyps"].to(DEV))
        assert got.shape == g["warped"][:, i - 1].shape
        assert rel_norm_err(got, g["warped"][:, i - 1]) < VOL_TOL


@pytest.mark.parametrize("name", ["warp_double_uniform", "warp_double_perpixel"])
def test_homo_warping_double_matches_golden(name):
    """module.homo_warping_double -> d3d_homo_warp_f64 (fp64 coordinates, fp32 sampling) against the live reference's
    module.py:560-601; fp32 projection matrices are refused as upstream refuses them (dtype error of torch.matmul)."""
    g = load_golden(name)
    for i in range(1, g["feats"].shape[0]):
        got = module.homo_warping_double(g["feats"][i:i + 1].to(DEV), g["proj"][:, i].to(DEV), g["proj"][:, 0].to(DEV),
                                         g["hyps"].to(DEV))
        assert got.shape == g["warped"][:, i - 1].shape and got.dtype == torch.float32
        assert rel_norm_err(got, g["warped"][:, i - 1]) < 1e-6
    with pytest.raises(RuntimeError):
        module.homo_warping_double(g["feats"][1:2].to(DEV), g["proj"][:, 1].float().to(DEV), g["proj"][:, 0].float().to(DEV),
                                   g["hyps"].to(DEV))


@pytest.mark.parametrize("name", ["cas_depthnet_uniform", "cas_depthnet_perpixel"])
def test_cas_depthnet_matches_golden(name):
    g = load_golden(name)
    cap = standins.Capture(standins.reg3d)
    out = depthnets.DepthNet()(_cuda_views(g["feats"]), g["proj"].to(DEV), g["hyps"].to(DEV), g["hyps"].shape[1], cap)
    assert rel_norm_err(cap.seen[0], g["variance"]) < VOL_TOL
    assert _depth_err(out["depth"], g["depth"]) < DEPTH_TOL
    # same logits on both sides -> confidence must agree too (window index may flip on a knife edge)
    ref_logits = standins.reg3d(g["variance"]).squeeze(1)[0].to(DEV)
    r = sweep.depth_regress(ref_logits, g["hyps"][0].to(DEV).contiguous(), conf_mode=sweep.CONF_WINDOW4)
    assert _depth_err(r["depth"], g["depth"][0]) < DEPTH_TOL
    close = (r["conf"].cpu() - g["conf"][0]).abs() < 1e-4
    assert close.float().mean() >= 0.999


def test_red_infer_matches_golden():
    g = load_golden("red_infer_depthnet")
    cap = standins.Capture(standins.slice_reg_red)
    out = depthnets.REDInferDepthNet()(_cuda_views(g["feats"]), g["proj"].to(DEV), g["hyps"].to(DEV),
                                       g["hyps"].shape[1], cap)
    got = torch.stack(cap.seen, 2)            # [B,C,D,H,W]
    assert rel_norm_err(got, g["variance_slices"]) < VOL_TOL
    assert _depth_err(out["depth"], g["depth"]) < DEPTH_TOL
    assert rel_norm_err(out["photometric_confidence"], g["conf"]) < 1e-3


class _Ada:
    """Stands where the reference's adamvs.InferDepthNet instance would (self.reg / self.reg_fuse / in_up)."""

    def __init__(self, in_up):
        self.in_up = in_up
        self.reg = standins.reg2d_pair
        self.fuse = standins.Capture(standins.slice_reg_up if in_up else standins.slice_reg_same)
        self.reg_fuse = self.fuse


def test_adamvs_infer_matches_golden():
    g = load_golden("ada_infer_depthnet")
    d = g["hyps"].shape[1]
    net = _Ada(in_up=True)
    out = depthnets.ada_infer_forward(net, _cuda_views(g["feats"]), g["proj"].to(DEV), g["hyps"].to(DEV), d)
    assert rel_norm_err(torch.stack(out["pair_result"], 1), g["pair_result"]) < DEPTH_TOL
    assert rel_norm_err(torch.stack(out["pair_confidence"][:3], 1), g["pair_conf_head"]) < 1e-3
    assert len(out["pair_confidence"]) == g["n_pair_confidence"]          # the reference's list quirk
    assert rel_norm_err(torch.stack(net.fuse.seen, 2), g["similarity_slices"]) < VOL_TOL
    assert _depth_err(out["depth"], g["depth"]) < DEPTH_TOL
    assert rel_norm_err(out["photometric_confidence"], g["conf"]) < 1e-3
    pairs = _ours_volume(g["feats"], g["proj"], g["hyps"], sweep.AGG_PAIR_MEAN)
    assert rel_norm_err(pairs, g["pair_volumes"][0]) < VOL_TOL
    # stage 2 consumes the first V-1 maps of the over-long list, resized again
    d2 = g["hyps2"].shape[1]
    net2 = _Ada(in_up=False)
    out2 = depthnets.ada_infer_forward(net2, _cuda_views(g["feats2"]), g["proj2"].to(DEV), g["hyps2"].to(DEV), d2,
                                       confidence_map=out["pair_confidence"])
    assert rel_norm_err(torch.stack(net2.fuse.seen, 2), g["similarity_slices2"]) < VOL_TOL
    assert _depth_err(out2["depth"], g["depth2"]) < DEPTH_TOL
    assert rel_norm_err(out2["photometric_confidence"], g["conf2"]) < 1e-3
    assert len(out2["pair_confidence"]) == g["n_pair_confidence2"]


class _GruLike(torch.nn.Module):
    """A recurrent stand-in with real state (the golden stand-ins are stateless): enough launches per plane
    for a CUDA graph to matter, and a result that depends on the order of the planes."""

    def __init__(self, up):
        super().__init__()
        self.up = up

    def forward(self, x, s1, s2):
        s1 = torch.tanh(0.7 * s1 + 0.3 * x[:, :8])
        s2 = 0.5 * s2 + 0.5 * torch.nn.functional.avg_pool2d(torch.cat([s1, x[:, 8:16]], 1), 2)
        logit = 1.5 * s1.mean(1, keepdim=True) + torch.nn.functional.interpolate(s2.mean(1, keepdim=True), scale_factor=2)
        if self.up:
            logit = torch.nn.functional.interpolate(logit, scale_factor=2, mode="nearest")
        return logit, s1, s2


@pytest.mark.parametrize("in_up", [True, False])
def test_plane_loop_graph_matches_eager(in_up):
    """Row f1: the D-plane regulariser + streaming soft-argmax loop replayed from a CUDA graph gives bit for bit
    what the eager loop gives, on fresh inputs, twice (static buffers are refilled, states re-zeroed)."""
    v, c, d, h, w = 5, 16, 12, 32, 64

    class Net:
        pass

    net = Net()
    net.in_up, net.reg, net.reg_fuse = in_up, standins.reg2d_pair, _GruLike(in_up)
    outs = {}
    for graphs in (False, True):
        depthnets.PLANE_LOOP_GRAPHS = graphs
        try:
            for seed in (21, 22):
                _, proj, feats, hyps = _scene(v, c, d, h, w, seed=seed, perpixel=True)
                conf = [torch.rand(1, 1, h, w, generator=torch.Generator().manual_seed(seed)).to(DEV) for _ in range(v - 1)]
                out = depthnets.ada_infer_forward(net, _cuda_views(feats), proj.to(DEV), hyps.to(DEV), d,
                                                  confidence_map=conf)
                outs[(graphs, seed)] = (out["depth"].clone(), out["photometric_confidence"].clone(),
                                        len(out["pair_confidence"]))
        finally:
            depthnets.PLANE_LOOP_GRAPHS = False
    for seed in (21, 22):
        assert torch.equal(outs[(True, seed)][0], outs[(False, seed)][0])
        assert torch.equal(outs[(True, seed)][1], outs[(False, seed)][1])
        assert outs[(True, seed)][2] == outs[(False, seed)][2]
    assert not torch.equal(outs[(True, 21)][0], outs[(True, 22)][0])


class _RedLike(torch.nn.Module):
    """Four-state recurrent stand-in with the signature of msrednet's slice regulariser."""

    def forward(self, x, s1, s2, s3, s4):
        pool = torch.nn.functional.avg_pool2d
        s1 = torch.tanh(0.6 * s1 + 0.4 * x[:, :8])
        s2 = 0.5 * s2 + 0.5 * pool(torch.cat([s1, s1], 1), 2)
        s3 = 0.5 * s3 + 0.5 * pool(torch.cat([s2, s2], 1), 2)
        s4 = 0.5 * s4 + 0.5 * pool(torch.cat([s3, s3], 1), 2)
        up = torch.nn.functional.interpolate
        logit = -2.0 * x.mean(1, keepdim=True) + s1.mean(1, keepdim=True) + up(s2.mean(1, keepdim=True), scale_factor=2) \
            + up(s4.mean(1, keepdim=True), scale_factor=8)
        return logit, s1, s2, s3, s4


@pytest.mark.parametrize("in_up", [True, False])
def test_batched_stateless_convs_match_the_plane_at_a_time_loop(in_up, monkeypatch):
    """depthnets.BATCH_STATELESS_CONVS (row f1): a regulariser with the reference's layer list (SliceCostRegNetRED,
    adamvs.py:403-427; here synth.SliceRegulariser, seeded) run with its stateless convolutions batched over all planes
    and one soft-argmax launch over the whole logit volume, against the plane-at-a-time loop of adamvs.py:492-529."""
    v, c, d, h, w = 4, 16, 12, 32, 40
    _, proj, feats, hyps = _scene(v, c, d, h, w, seed=31, smooth=True)
    views = _cuda_views(feats)
    hy = hyps.view(1, d, 1, 1).expand(1, d, h, w).contiguous().to(DEV)
    conf = [torch.rand(1, 1, h, w, generator=torch.Generator().manual_seed(k)).to(DEV) for k in range(v - 1)]
    net = types.SimpleNamespace(in_up=in_up, reg=None, reg_fuse=synth.SliceRegulariser(c, up=in_up, seed=3).to(DEV).eval())
    monkeypatch.setattr(depthnets, "BATCH_STATELESS_CONVS", False)
    want = depthnets.ada_infer_forward(net, views, proj.to(DEV), hy, d, confidence_map=conf)
    monkeypatch.setattr(depthnets, "BATCH_STATELESS_CONVS", True)
    got = depthnets.ada_infer_forward(net, views, proj.to(DEV), hy, d, confidence_map=conf)
    assert _depth_err(got["depth"], want["depth"]) < 1e-5
    assert float((got["photometric_confidence"] - want["photometric_confidence"]).abs().max()) < 1e-5
    assert len(got["pair_confidence"]) == len(want["pair_confidence"])


def test_red_plane_loop_graph_matches_eager():
    v, c, d, h, w = 3, 8, 10, 32, 64
    net, reg = object.__new__(type("Net", (), {})), _RedLike()
    outs = {}
    for graphs in (False, True):
        depthnets.PLANE_LOOP_GRAPHS = graphs
        try:
            for seed in (31, 32):
                _, proj, feats, hyps = _scene(v, c, d, h, w, seed=seed, perpixel=True)
                out = depthnets.red_infer_forward(net, _cuda_views(feats), proj.to(DEV), hyps.to(DEV), d, reg)
                outs[(graphs, seed)] = (out["depth"].clone(), out["photometric_confidence"].clone())
        finally:
            depthnets.PLANE_LOOP_GRAPHS = False
    for seed in (31, 32):
        assert torch.equal(outs[(True, seed)][0], outs[(False, seed)][0])
        assert torch.equal(outs[(True, seed)][1], outs[(False, seed)][1])


def test_adamvs_train_form_matches_golden():
    g = load_golden("ada_train_depthnet")
    w = g["pair_conf"][0, :, 0].to(DEV).contiguous()             # [V-1,h,w]
    fused = _ours_volume(g["feats"], g["proj"], g["hyps"], sweep.AGG_WEIGHTED_PRODUCT, weights=w, eps_in_numerator=True)
    assert rel_norm_err(fused, g["fused"][0]) < VOL_TOL
    r = sweep.depth_regress((-3.0 * g["fused"].mean(1))[0].to(DEV), g["hyps"][0].to(DEV).contiguous())
    assert _depth_err(r["depth"], g["depth"][0]) < DEPTH_TOL
    assert rel_norm_err(r["conf"], g["conf"][0]) < 1e-3


def test_regression_and_samples_match_golden():
    g = load_golden("regress_misc")
    prob = torch.softmax(g["logits"], 1)
    got = module.depth_regression(prob.to(DEV), g["hy_small"].to(DEV))          # resized 4-D hypotheses
    assert _depth_err(got, g["depth_resized"]) < 1e-5
    rng = torch.tensor([[400.0, 600.0]], device=DEV)
    got = module.get_depth_range_samples(rng, 10, 0.0, DEV, torch.float32, [1, 12, 16])
    assert torch.equal(got.cpu(), g["samples_range"])
    got = module.get_depth_range_samples(g["cur"].to(DEV), 8, 0.52, DEV, torch.float32, [1, 12, 16])
    assert torch.equal(got.cpu(), g["samples_cur"])


def test_cascade_stage_glue_matches_golden():
    g = load_golden("cas_stage_glue")
    fh, fw = [int(x) for x in g["full_hw"]]
    interval = (g["dmax"] - g["dmin"]) / g["num_depth"]
    nd = [int(x) for x in g["ndepths"]]
    ratios = [int(x) for x in g["ratios"]]
    dv1 = sweep.depth_samples(sweep.SAMPLES_CASCADE, nd[0], (fh // 4, fw // 4), device=DEV, dmin=g["dmin"],
                              dmax=g["dmax"], full_hw=(fh, fw))
    assert rel_norm_err(dv1, g["dv1"][0]) < 1e-6
    dv2 = sweep.depth_samples(sweep.SAMPLES_CASCADE, nd[1], (fh // 2, fw // 2), cur=g["depth1"][0].to(DEV),
                              interval=ratios[1] * interval, full_hw=(fh, fw))
    assert rel_norm_err(dv2, g["dv2"][0]) < 1e-6
    dv3 = sweep.depth_samples(sweep.SAMPLES_CASCADE, nd[2], (fh, fw), cur=g["depth2"][0].to(DEV),
                              interval=ratios[2] * interval, full_hw=(fh, fw))
    assert rel_norm_err(dv3, g["dv3"][0]) < 1e-6


def test_ucs_compute_depth_matches_golden():
    g = load_golden("ucs_compute_depth")
    out = depthnets.ucs_compute_depth(_cuda_views(g["feats"]), g["proj"].to(DEV), g["hyps"].to(DEV), standins.reg3d, 1.5)
    assert _depth_err(out["depth"], g["depth"]) < DEPTH_TOL
    assert rel_norm_err(out["variance"], g["exp_variance"]) < 2e-3
    var = _ours_volume(g["feats"], g["proj"], g["hyps"], sweep.AGG_VARIANCE)
    assert rel_norm_err(var, g["variance"][0]) < VOL_TOL


# ------------------------------------------------------------------ (2) oracle on seeded synthetic inputs
CASES = [  # v, c, d, h, w, perpixel
    (3, 8, 48, 128, 160, False),     # BASELINE.json config 1
    (5, 32, 24, 86, 58, False),      # config-2 shaped (1/8 linear size), 4 lanes x 8 channels per pixel
    (5, 32, 16, 86, 58, True),
    (5, 16, 12, 77, 45, True),       # cascade stage 2 shape class, ragged tile tail
    (5, 8, 8, 90, 61, True),         # cascade stage 3 shape class
    (2, 4, 6, 33, 47, False),        # one source view, 4 channels
    (7, 32, 10, 40, 36, False),      # 6 source views -> 4 channels per lane
    (9, 16, 6, 24, 28, True),        # 8 source views (the maximum)
    (3, 64, 6, 20, 24, False),       # 64 channels
    # 32 channels, <= 4 source views, H*W a multiple of 32: the four-planes-per-pass kernel (sweep_quad.cuh)
    (5, 32, 24, 64, 48, False),
    (5, 32, 18, 64, 48, True),       # per-pixel hypotheses, partial last batch
    (2, 32, 6, 32, 40, False),       # one source view
    (3, 32, 9, 40, 48, True),        # two source views
    (4, 32, 8, 48, 40, False),       # three source views
    # ... and its 16- / 8-channel forms (cascade stages 2 and 3): H*W a multiple of 64 / 128
    (5, 16, 12, 64, 48, True),
    (3, 16, 10, 32, 64, False),
    (5, 8, 8, 64, 96, True),
    (4, 8, 6, 48, 64, False),
]


# kernel variants (D3dCostVolumeArgs.variant, csrc/abi.cu): 0 = production kernels, 1 = baseline kernel (sweep_base),
# 2 = production kernel with __fdiv_rn instead of the shared-reciprocal division, 6 = the two-planes-per-pass kernel
# (sweep_lean) where variant 0 picks sweep_quad, 7 = sweep_quad spelled out, 8 / 9 = the TMA-prefetch experiments (sweep_ws:
# warp-specialised; sweep_pre: sweep_quad with prefetched footprints) where they are instantiated (32-channel features),
# 32 = sweep_quad's one-block re-fetch whatever the sweep length: the form long sweeps (D > 128) run, whose variance volume
# keeps its footprints relative to the reference texel
VARIANTS = [0, 1, 2, 6, 7, 8, 9, 32]


@pytest.mark.parametrize("v,c,d,h,w,perpixel", CASES)
@pytest.mark.parametrize("oracle_dev", ["cpu", "cuda"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_variance_matches_oracle(v, c, d, h, w, perpixel, oracle_dev, variant):
    _, proj, feats, hyps = _scene(v, c, d, h, w, seed=3, perpixel=perpixel)
    want = sweep_torch.variance_volume([f.to(oracle_dev) for f in _views(feats)], proj.to(oracle_dev),
                                       hyps.to(oracle_dev))[0]
    got = _ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE, variant=variant)
    assert got.shape == want.shape
    assert rel_norm_err(got, want) < VOL_TOL


@pytest.mark.parametrize("d", [1, 2, 3, 7])
def test_short_sweeps_and_depth_chunks(d):
    """Few planes on a small image: the launch splits the sweep into depth chunks (blockIdx.y) to fill the
    SMs; every chunk re-warms its footprints and must agree with the oracle."""
    _, proj, feats, hyps = _scene(5, 32, d, 24, 20, seed=12)
    want = sweep_torch.variance_volume(_views(feats), proj, hyps)[0]
    for variant in VARIANTS:
        assert rel_norm_err(_ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE, variant=variant), want) < VOL_TOL


@pytest.mark.parametrize("smooth", [False, True])
@pytest.mark.parametrize("variant", VARIANTS)
def test_whu_shaped_rig_matches_oracle(smooth, variant):
    """The WHU-OMVS cross rig of SURVEY.md §8d (f=4000 at 1856x2752, 40 m baselines, 400-600 m) on a crop-sized
    feature map: coordinates of the production magnitude, so fp32 rounding of the projection matters."""
    rig = synth.make_rig(num_views=5)
    h, w, d, c = 172, 116, 12, 32          # 1/16 of full res: scale 16 keeps the full-size geometry
    _, proj, feats, hyps = _scene(5, c, d, h, w, seed=5, smooth=smooth, rig=rig, scale=16)
    want = sweep_torch.variance_volume(_cuda_views(feats), proj.to(DEV), hyps.to(DEV))[0]
    got = _ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE, variant=variant)
    assert rel_norm_err(got, want) < VOL_TOL


@pytest.mark.parametrize("groups", [1, 4, 8, 16, 32])
def test_group_corr_matches_oracle(groups):
    _, proj, feats, hyps = _scene(5, 32, 10, 40, 52, seed=4)
    want = sweep_torch.groupwise_correlation_volume(_views(feats), proj, hyps, groups)[0]
    got = _ours_volume(feats, proj, hyps, sweep.AGG_GROUP_CORR, groups=groups)
    assert got.shape == want.shape
    assert rel_norm_err(got, want) < VOL_TOL


@pytest.mark.parametrize("v,groups,perpixel,d", [(3, 8, True, 9), (4, 2, True, 6), (2, 4, False, 13), (5, 8, False, 24),
                                                 (5, 32, True, 7)])
@pytest.mark.parametrize("variant", [0, 48])
def test_group_corr_dot_cache_and_coefficient_cache_agree_with_the_oracle(v, groups, perpixel, d, variant):
    """Group-wise correlation of 32-channel features: variant 0 keeps cached corner DOT PRODUCTS where a group spans a
    lane's 4 channels or more (groups <= 8), variant 48 the 16 interpolation coefficients; per-pixel hypotheses, 1-4
    source views, ragged last batches.  (The oracle's group-wise volume is a restatement by analogy with adamvs.py:473:
    parity unpinned by the reference for this mode.)"""
    _, proj, feats, hyps = _scene(v, 32, d, 48, 40, seed=40 + v, perpixel=perpixel)
    want = sweep_torch.groupwise_correlation_volume(_views(feats), proj, hyps, groups)[0]
    got = _ours_volume(feats, proj, hyps, sweep.AGG_GROUP_CORR, groups=groups, variant=variant)
    assert got.shape == want.shape
    assert rel_norm_err(got, want) < VOL_TOL


def test_group_corr_wide_groups_lane_reduction():
    _, proj, feats, hyps = _scene(7, 32, 6, 20, 28, seed=6)      # 4 channels per lane, groups of 16 span 4 lanes
    want = sweep_torch.groupwise_correlation_volume(_views(feats), proj, hyps, 2)[0]
    got = _ours_volume(feats, proj, hyps, sweep.AGG_GROUP_CORR, groups=2)
    assert rel_norm_err(got, want) < VOL_TOL


@pytest.mark.parametrize("c,h,w", [(16, 36, 44), (32, 40, 48), (16, 32, 64), (8, 64, 64)])   # base / quad / acc kernels
@pytest.mark.parametrize("eps_num", [False, True])
@pytest.mark.parametrize("variant", [0, 10, 11, 12, 13, 14])   # production; sweep_acc / sweep_win A/B forms (abi.cu)
def test_weighted_product_matches_oracle(eps_num, c, h, w, variant):
    v, d = 5, 8
    _, proj, feats, hyps = _scene(v, c, d, h, w, seed=7, perpixel=True)
    g = torch.Generator().manual_seed(1)
    weights = [torch.rand(1, 1, h // 2, w // 2, generator=g) for _ in range(v - 1)]
    want = sweep_torch.weighted_product_volume(_views(feats), proj, hyps, weights, eps_in_numerator=eps_num)[0]
    wt = torch.cat([sweep_torch.resize_weight(x, h, w) for x in weights], 1)[0].to(DEV).contiguous()
    got = _ours_volume(feats, proj, hyps, sweep.AGG_WEIGHTED_PRODUCT, weights=wt, eps_in_numerator=eps_num, variant=variant)
    assert rel_norm_err(got, want) < VOL_TOL


@pytest.mark.parametrize("v,c,d,h,w,perpixel", [(5, 8, 8, 45, 67, True), (3, 8, 11, 33, 50, True), (4, 16, 32, 40, 56, True),
                                                (5, 16, 19, 31, 37, False), (2, 32, 13, 24, 40, True), (5, 8, 3, 130, 3, False)])
@pytest.mark.parametrize("variant", [10, 12, 13, 14])
def test_register_accumulated_sweep_chunks_slices_and_borders(v, c, d, h, w, perpixel, variant):
    """sweep_acc.cuh / sweep_win.cuh (variants 10, 14 / 12, 13 force them wherever they are instantiated): chunks of 4 or 8 planes with a ragged last
    chunk, ragged pixel tiles, uniform and per-pixel hypotheses, 1..4 source views; every plane-slice launch and the
    plane-major layout reproduce the whole-volume launch bit for bit (a plane's result does not depend on which
    footprint the lane happened to hold); the synthetic rig's oblique views put part of every source image out of
    bounds, so the zeros-padding side path runs."""
    _, proj, feats, hyps = _scene(v, c, d, h, w, seed=70 + c + d, perpixel=perpixel)
    g = torch.Generator().manual_seed(d)
    weights = [torch.rand(1, 1, h, w, generator=g) for _ in range(v - 1)]
    want = sweep_torch.weighted_product_volume(_views(feats), proj, hyps, weights)[0]
    wt = torch.cat(weights, 1)[0].to(DEV).contiguous()
    kw = dict(weights=wt, variant=variant)
    full = _ours_volume(feats, proj, hyps, sweep.AGG_WEIGHTED_PRODUCT, **kw)
    assert rel_norm_err(full, want) < VOL_TOL
    pm = _ours_volume(feats, proj, hyps, sweep.AGG_WEIGHTED_PRODUCT, plane_major=True, **kw)
    assert torch.equal(pm.permute(1, 0, 2, 3), full)
    for d0, dn in ((0, 1), (d // 2, 2), (1, d - 1), (max(d - 9, 0), min(9, d))):
        part = _ours_volume(feats, proj, hyps, sweep.AGG_WEIGHTED_PRODUCT, d_begin=d0, d_count=dn, **kw)
        assert torch.equal(part, full[:, d0:d0 + dn])
    old = _ours_volume(feats, proj, hyps, sweep.AGG_WEIGHTED_PRODUCT, weights=wt, variant=11)
    assert rel_norm_err(full, old) < 1e-5      # the kernels it replaces (differenced corners instead of corner weights)


@pytest.mark.parametrize("v,h,w", [(5, 43, 29), (5, 40, 48), (3, 32, 40)])      # base kernel / quad kernel
def test_pair_mean_matches_oracle(v, h, w):
    _, proj, feats, hyps = _scene(v, 32, 12, h, w, seed=8)
    hy4 = hyps.view(1, -1, 1, 1).repeat(1, 1, h, w)
    want = torch.stack(sweep_torch.pair_mean_volumes(_views(feats), proj, hy4), 1)[0]
    got = _ours_volume(feats, proj, hy4, sweep.AGG_PAIR_MEAN)
    assert rel_norm_err(got, want) < VOL_TOL


@pytest.mark.parametrize("v,c,h,w", [(3, 8, 40, 48), (5, 32, 40, 48), (4, 16, 32, 64), (5, 32, 43, 29)])
def test_plane_slices_and_plane_major_layout(v, c, h, w):
    """Slice launches (the plane-at-a-time callers) and the plane-major layout reproduce the whole-volume launch
    bit for bit on every kernel family (direct, quad with 8 / 4 lanes per pixel, lean)."""
    _, proj, feats, hyps = _scene(v, c, 12, h, w, seed=9)
    full = _ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE)
    part = _ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE, d_begin=5, d_count=4)
    assert torch.equal(part, full[:, 5:9])
    pm = _ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE, plane_major=True)
    assert torch.equal(pm.permute(1, 0, 2, 3), full)
    one = _ours_volume(feats, proj, hyps[:, 3:4], sweep.AGG_VARIANCE)           # D = 1
    assert torch.equal(one, full[:, 3:4])
    for d0 in range(0, 12, 3):                                                   # three-plane slices, any phase
        assert torch.equal(_ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE, d_begin=d0, d_count=3), full[:, d0:d0 + 3])


def test_points_behind_the_camera_contribute_zero():
    """z <= 0 is unguarded upstream (module.py:542 divides by it); here such samples are defined to be
    out of bounds, so the volume stays finite (SURVEY.md §7 hard part 6)."""
    _, proj, feats, hyps = _scene(3, 8, 6, 24, 32, seed=10)
    hyps = hyps.clone()
    hyps[0, 0] = -5.0
    hyps[0, 1] = 0.0
    got = _ours_volume(feats, proj, hyps, sweep.AGG_VARIANCE)
    assert torch.isfinite(got).all()
    want = sweep_torch.variance_volume(_views(feats), proj, hyps[:, 2:])[0]
    assert rel_norm_err(got[:, 2:], want) < VOL_TOL


# kernel 2 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,h,w", [(48, 64, 80), (384, 43, 29), (8, 30, 50), (5, 7, 9)])
@pytest.mark.parametrize("conf", ["max", "win4"])
def test_softmax_regression_matches_oracle(d, h, w, conf):
    logits = synth.planted_logits(d, h, w, seed=d)
    hyps = synth.uniform_hypotheses(400.0, 600.0, d)
    fn = sweep_torch.regress_maxprob if conf == "max" else sweep_torch.regress_window4
    depth, cf, idx = fn(logits.unsqueeze(0), hyps.unsqueeze(0))
    mode = sweep.CONF_MAX_PROB if conf == "max" else sweep.CONF_WINDOW4
    r = sweep.depth_regress(logits.to(DEV), hyps.to(DEV), conf_mode=mode)
    assert _depth_err(r["depth"], depth[0]) < DEPTH_TOL
    agree = (r["index"].cpu().long() == idx[0]).float().mean()
    assert agree >= 0.999
    same = r["index"].cpu().long() == idx[0]
    assert ((r["conf"].cpu() - cf[0]).abs()[same] < 1e-5).all()


def test_softmax_regression_per_pixel_hypotheses_and_next_stage():
    d, h, w = 32, 40, 56
    rig = synth.make_rig()
    logits = synth.planted_logits(d, h, w, seed=2)
    hyps = synth.per_pixel_hypotheses(synth.smooth_depth_map(rig, h, w), d, 1.04)
    depth, cf, idx = sweep_torch.regress_maxprob(logits.unsqueeze(0), hyps.unsqueeze(0))
    r = sweep.depth_regress(logits.to(DEV), hyps.to(DEV), next_num_depth=8, next_interval=0.52)
    assert _depth_err(r["depth"], depth[0]) < DEPTH_TOL
    want_next = sweep_torch.depth_range_samples(r["depth"].cpu().unsqueeze(0), 8, 0.52, [1, h, w])[0]
    assert torch.equal(r["next_hyps"].cpu(), want_next)


def test_streaming_raw_exp_matches_oracle_slice_by_slice():
    d, h, w = 12, 20, 28
    logits = 0.5 * synth.planted_logits(d, 2 * h, 2 * w, seed=4)           # adamvs stage 1: logits at 2x
    hyps = synth.per_pixel_hypotheses(synth.smooth_depth_map(synth.make_rig(), h, w), d, 4.0)
    want_d, want_c = sweep_torch.regress_streaming([logits[k].view(1, 1, 2 * h, 2 * w) for k in range(d)],
                                                   [hyps[k].view(1, 1, h, w) for k in range(d)], upsample2=True)
    state = torch.zeros(3, 2 * h, 2 * w, device=DEV)
    lg = logits.to(DEV)
    hy = hyps.to(DEV)
    for k in range(d):
        r = sweep.depth_regress(lg[k:k + 1], hy, softmax_mode=sweep.SOFTMAX_RAW_EXP, d_begin=k, state=state,
                                finalize=(k == d - 1))
    assert _depth_err(r["depth"], want_d[0]) < DEPTH_TOL
    assert rel_norm_err(r["conf"], want_c[0]) < 1e-4
    whole = sweep.depth_regress(lg, hy, softmax_mode=sweep.SOFTMAX_RAW_EXP)
    assert _depth_err(whole["depth"], want_d[0]) < DEPTH_TOL


@pytest.mark.parametrize("group", [1, 3, 5, 16])
def test_streaming_raw_exp_scattered_planes_are_bit_identical_to_plane_at_a_time(group):
    """D3dRegressArgs.logit_planes: K separately allocated planes folded into the accumulators by one launch give
    the very bytes K single-plane launches give (same fp32 operations in the same plane order)."""
    d, h, w = 37, 22, 30
    lg = (0.5 * synth.planted_logits(d, 2 * h, 2 * w, seed=9)).to(DEV)
    hy = synth.per_pixel_hypotheses(synth.smooth_depth_map(synth.make_rig(), h, w), d, 4.0).to(DEV)
    state1 = torch.zeros(3, 2 * h, 2 * w, device=DEV)
    for k in range(d):
        one = sweep.depth_regress(lg[k:k + 1], hy, softmax_mode=sweep.SOFTMAX_RAW_EXP, d_begin=k, state=state1,
                                  finalize=(k == d - 1))
    stateg = torch.zeros(3, 2 * h, 2 * w, device=DEV)
    for k0 in range(0, d, group):
        planes = [lg[k].clone().view(1, 2 * h, 2 * w) for k in range(k0, min(k0 + group, d))]   # scattered allocations
        many = sweep.depth_regress(planes, hy, softmax_mode=sweep.SOFTMAX_RAW_EXP, d_begin=k0, state=stateg,
                                   finalize=(k0 + group >= d))
    assert torch.equal(many["depth"], one["depth"]) and torch.equal(many["conf"], one["conf"])
    assert torch.equal(stateg, state1)
    with pytest.raises(ValueError, match="1..16"):
        sweep.depth_regress([lg[0]] * 17, hy, softmax_mode=sweep.SOFTMAX_RAW_EXP, d_begin=0, state=stateg)
    with pytest.raises(ValueError, match="RAW_EXP"):
        sweep.depth_regress([lg[0]], hy, softmax_mode=sweep.SOFTMAX_STABLE)


@pytest.mark.parametrize("batch_planes", [1, 4, 16])
def test_stream_batch_cadence_does_not_change_the_plane_at_a_time_models(batch_planes, monkeypatch):
    """depthnets.STREAM_BATCH_PLANES only changes how often the accumulators travel: AdaMVS and RED-Net inference
    forwards give identical bytes at every cadence (1 = upstream's per-plane update)."""
    def run(k):
        monkeypatch.setattr(depthnets, "STREAM_BATCH_PLANES", k)
        g = load_golden("ada_infer_depthnet")
        ada = depthnets.ada_infer_forward(_Ada(in_up=True), _cuda_views(g["feats"]), g["proj"].to(DEV), g["hyps"].to(DEV),
                                          g["hyps"].shape[1])
        g = load_golden("red_infer_depthnet")
        red = depthnets.REDInferDepthNet()(_cuda_views(g["feats"]), g["proj"].to(DEV), g["hyps"].to(DEV),
                                           g["hyps"].shape[1], standins.slice_reg_red)
        return ada["depth"], ada["photometric_confidence"], red["depth"], red["photometric_confidence"]
    base = run(1)
    got = run(batch_planes)
    assert all(torch.equal(a, b) for a, b in zip(base, got))


def test_ucs_uncertainty_samples_are_bit_identical_to_the_reference_s_elementwise_ops():
    """ucsnet.uncertainty_aware_samples (ucsnet.py:41-51) -> d3d_depth_samples SPREAD: cur -/+ exp_var in D steps + 1e-12,
    every fp32 rounding in torch's order."""
    rig = synth.tiny_rig()
    h, w = 37, 53
    cur = synth.smooth_depth_map(rig, h, w, seed=2).view(1, 1, h, w)
    spread = (0.05 + torch.rand(1, 1, h, w, generator=torch.Generator().manual_seed(5))) * 0.7
    for nd in (2, 8, 33):
        want = sweep_torch.uncertainty_samples(cur, spread, nd)
        got = depthnets.ucs_uncertainty_samples(cur.to(DEV), spread.to(DEV), nd, DEV, torch.float32, [1, h, w])
        assert got.shape == want.shape and torch.equal(got.cpu(), want)
    first = depthnets.ucs_uncertainty_samples(torch.tensor([[rig.dmin, rig.dmax]], device=DEV), None, 8, DEV, torch.float32,
                                              [1, h, w])
    assert torch.equal(first.cpu(), sweep_torch.depth_range_samples(torch.tensor([[rig.dmin, rig.dmax]]), 8, 0.0, [1, h, w]))


def test_exp_variance_matches_oracle():
    d, h, w = 8, 24, 24
    logits = synth.planted_logits(d, h, w, seed=6) * 0.3
    hyps = synth.per_pixel_hypotheses(synth.smooth_depth_map(synth.make_rig(), h, w), d, 2.0)
    prob = torch.softmax(logits, 0).unsqueeze(0)
    depth = sweep_torch.expectation(prob, hyps.unsqueeze(0))
    want = sweep_torch.exp_variance(prob, hyps.unsqueeze(0), depth, 1.5)
    r = sweep.depth_regress(logits.to(DEV), hyps.to(DEV), conf_mode=sweep.CONF_WINDOW4, lamb=1.5)
    assert rel_norm_err(r["exp_variance"], want[0]) < 1e-3


def test_texel_relayout_round_trip():
    f = synth.make_features(3, 16, 37, 53, seed=1)
    tex = sweep.to_texels(f.to(DEV))
    assert torch.equal(tex.cpu(), f.permute(0, 2, 3, 1).contiguous())


# ------------------------------------------------------------------ (3) properties at full size
@pytest.mark.parametrize("mode,kw", [(sweep.AGG_VARIANCE, {"variant": 0}), (sweep.AGG_VARIANCE, {"variant": 1}),
                                     (sweep.AGG_VARIANCE, {"variant": 2}), (sweep.AGG_VARIANCE, {"variant": 6}),
                                     (sweep.AGG_VARIANCE, {"variant": 7}),
                                     (sweep.AGG_VARIANCE, {"variant": 8}), (sweep.AGG_VARIANCE, {"variant": 9}),
                                     (sweep.AGG_GROUP_CORR, {"groups": 8})])
def test_full_size_config_against_cuda_aten_on_plane_subsets(mode, kw):
    """BASELINE.json configs 2 and 4 (V=5, C=32, D=384, 688x464): the whole volume is built in one launch;
    plane subsets are checked against the reference's ATen path on this GPU (the full ATen volume needs
    47 GB of temporaries), and a slice launch must reproduce the same planes bit for bit."""
    rig = synth.make_rig(num_views=5)
    h, w, c, d = 688, 464, 32, 384
    feats = synth.make_features(5, c, h, w, seed=0).to(DEV)
    proj = torch.from_numpy(rig.proj(4)).unsqueeze(0).to(DEV)
    hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d, device=DEV)
    tex = sweep.to_texels(feats)
    pose = sweep.relative_poses(proj[0])
    vol = sweep.cost_volume(tex, pose, hyps, mode, **kw)
    assert torch.isfinite(vol[:, ::37]).all()
    views = [feats[i:i + 1] for i in range(5)]
    for d0 in (0, 190, 380):
        sub = hyps[d0:d0 + 4].unsqueeze(0)
        if mode == sweep.AGG_VARIANCE:
            want = sweep_torch.variance_volume(views, proj, sub)[0]
        else:
            want = sweep_torch.groupwise_correlation_volume(views, proj, sub, kw["groups"])[0]
        err = rel_norm_err(vol[:, d0:d0 + 4], want)
        print("full-size planes %d..%d %s: rel err %.3e" % (d0, d0 + 3, kw, err))
        assert err < VOL_TOL
        part = sweep.cost_volume(tex, pose, hyps, mode, d_begin=d0, d_count=4, **kw)
        if mode == sweep.AGG_VARIANCE and kw.get("variant") in (0, 7):
            # the 4-plane slice runs sweep_quad's short-sweep form, the 384-plane launch the long-sweep form whose
            # footprints are kept relative to the reference texel: the same volume, rounded differently
            assert rel_norm_err(part, vol[:, d0:d0 + 4]) < 1e-5
        else:
            assert torch.equal(part, vol[:, d0:d0 + 4])
    # source views are interchangeable for the aggregate (up to summation order)
    perm = [0, 3, 1, 4, 2]
    vol_p = sweep.cost_volume(tex[perm].contiguous(), sweep.relative_poses(proj[0][perm]), hyps, mode, d_begin=100,
                              d_count=8, **kw)
    assert rel_norm_err(vol_p, vol[:, 100:108]) < 1e-5
    del vol
    torch.cuda.empty_cache()


def test_full_size_config_every_plane_against_cuda_aten():
    """BASELINE.json config 2, ALL 384 planes of the launch the benchmark times (the production kernel, variant 0)
    against the reference's ATen path on this GPU, 48 planes at a time (the ATen temporaries of a chunk fit)."""
    rig = synth.make_rig(num_views=5)
    h, w, c, d = 688, 464, 32, 384
    feats = synth.make_features(5, c, h, w, seed=1).to(DEV)
    proj = torch.from_numpy(rig.proj(4)).unsqueeze(0).to(DEV)
    hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, d, device=DEV)
    vol = sweep.cost_volume(sweep.to_texels(feats), sweep.relative_poses(proj[0]), hyps, sweep.AGG_VARIANCE)
    views = [feats[i:i + 1] for i in range(5)]
    worst = 0.0
    for d0 in range(0, d, 48):
        want = sweep_torch.variance_volume(views, proj, hyps[d0:d0 + 48].unsqueeze(0))[0]
        err = float((vol[:, d0:d0 + 48] - want).abs().max() / want.abs().max())      # on the device: 2 GB per chunk
        worst = max(worst, err)
        del want
        assert err < VOL_TOL, "planes %d..%d: %.3e" % (d0, d0 + 47, err)
    print("cfg2, all %d planes against CUDA-ATen: worst chunk rel err %.3e" % (d, worst))
    del vol
    torch.cuda.empty_cache()


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_full_size_cascade_stage_against_cuda_aten(stage):
    """BASELINE.json config 3 (AdaMVS at 1856x2752): every stage's weighted-product volume (and stage 1's pair
    volumes) at its production size and geometry -- 32 ch @ 688x464, 16 ch @ 1376x928, 8 ch @ 2752x1856, per-pixel
    hypotheses -- checked on plane subsets against the reference's ATen path on this GPU."""
    scale, c, d, ratio = {1: (4, 32, 48, 4), 2: (2, 16, 32, 2), 3: (1, 8, 8, 1)}[stage]
    v = 5
    rig = synth.make_rig(num_views=v)
    h, w = 2752 // scale, 1856 // scale
    g = torch.Generator().manual_seed(40 + stage)
    feats = torch.randn(v, c, h, w, generator=g).to(DEV)
    proj = torch.from_numpy(rig.proj(scale)).unsqueeze(0).to(DEV)
    interval = ratio * (rig.dmax - rig.dmin) / 384
    if stage == 1:
        hyps = sweep.depth_samples(sweep.SAMPLES_RANGE, d, (h, w), device=torch.device(DEV), dmin=rig.dmin, dmax=rig.dmax)
    else:
        cur = synth.smooth_depth_map(rig, h, w, seed=stage).to(DEV)
        hyps = sweep.depth_samples(sweep.SAMPLES_AROUND, d, (h, w), cur=cur, interval=interval)
    weights = [torch.rand(1, 1, h, w, generator=g).to(DEV) for _ in range(v - 1)]
    tex = sweep.to_texels(feats)
    pose = sweep.relative_poses(proj[0])
    wt = torch.cat(weights, 1)[0].contiguous()
    vol = sweep.cost_volume(tex, pose, hyps, sweep.AGG_WEIGHTED_PRODUCT, weights=wt)
    views = [feats[i:i + 1] for i in range(v)]
    for d0 in sorted({0, d // 2, d - 2}):
        sub = hyps[d0:d0 + 2].unsqueeze(0)
        want = sweep_torch.weighted_product_volume(views, proj, sub, weights)[0]
        err = rel_norm_err(vol[:, d0:d0 + 2], want)
        print("stage %d weighted product planes %d..%d: rel err %.3e" % (stage, d0, d0 + 1, err))
        assert err < VOL_TOL
        del want
    del vol
    if stage == 1:
        pairs = sweep.cost_volume(tex, pose, hyps, sweep.AGG_PAIR_MEAN)
        for d0 in (0, 23, 46):
            want = torch.stack(sweep_torch.pair_mean_volumes(views, proj, hyps[d0:d0 + 2].unsqueeze(0)), 1)[0]
            assert rel_norm_err(pairs[:, d0:d0 + 2], want) < VOL_TOL


def test_rays_from_the_reference_matmul_are_exact_at_every_size():
    """cuBLAS rounds rot @ [x,y,1] differently past column 2^20 - 32 when H*W = 1376*928 (tools/diag_rays_n.py): the
    kernel's own rays then differ from the reference's by an ulp in ~5 % of those columns (1.4e-4 on the variance
    volume); with the reference's own product handed in (the default) the volume agrees to rounding noise.  At
    sizes where the library keeps one order the two paths are bit-identical."""
    rig = synth.make_rig(num_views=5)
    h, w, c, d = 1376, 928, 16, 32
    g = torch.Generator().manual_seed(11)
    feats = torch.randn(5, c, h, w, generator=g).to(DEV)
    proj = torch.from_numpy(rig.proj(2)).unsqueeze(0).to(DEV)
    cur = synth.smooth_depth_map(rig, h, w, seed=2).to(DEV)
    hyps = sweep.depth_samples(sweep.SAMPLES_AROUND, d, (h, w), cur=cur, interval=2 * (rig.dmax - rig.dmin) / 384)
    tex, pose = sweep.to_texels(feats), sweep.relative_poses(proj[0])
    want = sweep_torch.variance_volume([feats[i:i + 1] for i in range(5)], proj, hyps[3:5].unsqueeze(0))[0]
    exact = sweep.cost_volume(tex, pose, hyps, sweep.AGG_VARIANCE, d_begin=3, d_count=2)
    own = sweep.cost_volume(tex, pose, hyps, sweep.AGG_VARIANCE, d_begin=3, d_count=2, exact_rays=False)
    err_exact, err_own = rel_norm_err(exact, want), rel_norm_err(own, want)
    print("1376x928 variance: reference rays %.3e, kernel rays %.3e" % (err_exact, err_own))
    assert err_exact < 1e-6
    assert err_own < 5e-4                      # an ulp of coordinate on noise features, documented in DESIGN.md
    _, proj2, feats2, hyps2 = _scene(5, 32, 8, 64, 48, seed=2)
    tex2, pose2 = sweep.to_texels(feats2.to(DEV)), sweep.relative_poses(proj2[0].to(DEV))
    a = sweep.cost_volume(tex2, pose2, hyps2[0].to(DEV), sweep.AGG_VARIANCE)
    b = sweep.cost_volume(tex2, pose2, hyps2[0].to(DEV), sweep.AGG_VARIANCE, exact_rays=False)
    assert rel_norm_err(a, b) < VOL_TOL


@pytest.mark.parametrize("h,w", [(128, 160), (344, 232), (688, 464), (1376, 928), (2752, 1856), (43, 29)])
def test_rays_for_skips_the_matmul_only_where_the_kernel_order_is_the_reference_order(h, w):
    """sweep.rays_for decides per image size whether the sweeps may form rot @ [x,y,1] themselves.  The verdict is
    taken on a probe rotation and the first pose; it must then hold for poses it has never seen (the rounding
    order is a property of the cuBLAS kernel chosen for the shape, not of the values)."""
    rig = synth.make_rig(num_views=5)
    pose_a = sweep.relative_poses(torch.from_numpy(rig.proj(4)).to(DEV))
    got = sweep.rays_for(pose_a, h, w)
    if got is not None:
        assert torch.equal(got, sweep.reference_rays(pose_a, h, w))
        assert sweep._RAY_ORDER[sweep._ray_key(h, w, pose_a.device)] is False
        return
    g = torch.Generator().manual_seed(h * 7 + w)
    for scale in (1.0, 2.0):
        fresh = pose_a.clone()
        fresh[:, :3, :3] += 0.03 * torch.randn(4, 3, 3, generator=g).to(DEV)
        fresh[:, :2, :] /= scale
        assert torch.equal(sweep.kernel_rays(fresh, h, w), sweep.reference_rays(fresh, h, w))
    assert sweep.rays_for(fresh, h, w) is None


def test_rays_verdict_is_retaken_when_the_matmul_settings_change():
    """The verdict is keyed on TF32 permission / float32 matmul precision too: with TF32 allowed the reference's own
    product is a different number, so the sweeps must be handed THAT product again, not the cached "same order"."""
    rig = synth.make_rig(num_views=5)
    pose = sweep.relative_poses(torch.from_numpy(rig.proj(4)).to(DEV))
    h, w = 96, 128
    base = sweep.rays_for(pose, h, w)
    before = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = not before
        assert sweep._ray_key(h, w, pose.device) not in sweep._RAY_ORDER
        other = sweep.rays_for(pose, h, w)
        assert sweep._ray_key(h, w, pose.device) in sweep._RAY_ORDER
        want = sweep.reference_rays(pose, h, w)
        assert (other is None and torch.equal(want, sweep.kernel_rays(pose, h, w))) or torch.equal(other, want)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = before
    again = sweep.rays_for(pose, h, w)
    assert (again is None) == (base is None)


def test_view_pipeline_matches_direct_calls():
    """pipeline.ViewPipeline (pinned host in, pinned host out, two slots, copy + compute streams): five views through
    two slots give what the synchronous calls give, bit for bit."""
    from deep3d_aerial_b200.pipeline import ViewPipeline

    v, c, d, h, w = 5, 32, 12, 40, 48
    pipe = ViewPipeline(v, c, h, w, d, DEV)
    views = []
    for seed in range(5):
        _, proj, feats, hyps = _scene(v, c, d, h, w, seed=50 + seed)
        logits = synth.planted_logits(d, h, w, seed=seed).to(DEV)
        views.append((feats.pin_memory(), proj[0].contiguous().pin_memory(), hyps[0].contiguous().pin_memory(), logits))
    got = []
    for i, (f, pr, hy, lg) in enumerate(views):
        pipe.submit(f, pr, hy, lambda vol, lg=lg: lg)
        if i:
            dep, conf = pipe.collect()
            got.append((dep.clone(), conf.clone()))
    dep, conf = pipe.collect()
    got.append((dep.clone(), conf.clone()))
    for (f, pr, hy, lg), (dep, conf) in zip(views, got):
        tex = sweep.to_texels(f.to(DEV))
        sweep.cost_volume(tex, sweep.relative_poses(pr.to(DEV)), hy.to(DEV), sweep.AGG_VARIANCE)
        r = sweep.depth_regress(lg, hy.to(DEV), want_index=False)
        assert torch.equal(dep, r["depth"].cpu()) and torch.equal(conf, r["conf"].cpu())
    assert pipe.h2d_bytes == 4 * (v * c * h * w + 16 * v + d) and pipe.d2h_bytes == 8 * h * w


@pytest.mark.parametrize("texels", [True, False])
@pytest.mark.parametrize("capacity", [15, 10])
def test_view_pipeline_image_residency_uploads_each_image_once_and_builds_the_same_volumes(capacity, texels):
    """ViewPipeline.submit(image_ids=...): the views of a scene block share their images; feature maps of resident
    images are not copied again, evicted buffers are not refilled under a sweep that still reads them, and every view
    gets the volume of ITS images (the regulariser here reads the volume, so a wrong image shows in the depth)."""
    from deep3d_aerial_b200.pipeline import ViewPipeline

    v, c, d, h, w, n_img = 5, 32, 8, 40, 48, 11
    rig, proj, _, hyps = _scene(v, c, d, h, w, seed=7)
    g = torch.Generator().manual_seed(99)
    images = [torch.randn(c, h, w, generator=g).pin_memory() for _ in range(n_img)]
    pr, hy = proj[0].contiguous().pin_memory(), hyps[0].contiguous().pin_memory()
    pipe = ViewPipeline(v, c, h, w, d, DEV, resident_images=capacity, resident_texels=texels)   # texel pool / per-image maps
    regulariser = lambda vol: (-4.0 * vol.mean(0)).contiguous()           # noqa: E731
    windows = [[i] + [j for j in range(min(max(i - 2, 0), n_img - v), min(max(i - 2, 0), n_img - v) + v) if j != i]
               for i in range(n_img)]
    order = list(range(n_img)) + [0, 1]                 # ... and back to images that were evicted meanwhile
    got = []
    for k, i in enumerate(order):
        pipe.submit([images[j] for j in windows[i]], pr, hy, regulariser, image_ids=windows[i])
        if k:
            dep, conf = pipe.collect()
            got.append((dep.clone(), conf.clone()))
    dep, conf = pipe.collect()
    got.append((dep.clone(), conf.clone()))
    for i, (dep, conf) in zip(order, got):
        feats = torch.stack([images[j] for j in windows[i]]).to(DEV)
        vol = sweep.cost_volume(sweep.to_texels(feats), sweep.relative_poses(pr.to(DEV)), hy.to(DEV), sweep.AGG_VARIANCE)
        r = sweep.depth_regress(regulariser(vol), hy.to(DEV), want_index=False)
        assert torch.equal(dep, r["depth"].cpu()) and torch.equal(conf, r["conf"].cpu()), "view %d" % i
    assert pipe.lru.hits + pipe.lru.misses == v * len(order)
    assert (pipe.lru.texels is not None) == texels
    if capacity >= n_img:   # every image is uploaded exactly once, the return visits find theirs resident
        assert pipe.lru.misses == n_img
    else:                   # the return visits re-upload what the walk evicted
        assert n_img < pipe.lru.misses <= n_img + v
    per_image = 4 * c * h * w
    assert pipe.h2d_bytes_total == pipe.lru.misses * per_image + len(order) * 4 * (16 * v + d)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_two_devices():
    """The shared-memory opt-in of the sweep kernels is a per-device attribute: a process that drives a second
    GPU after the first must get the same volumes there."""
    _, proj, feats, hyps = _scene(5, 32, 12, 64, 48, seed=3)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        with torch.cuda.device(dev):
            tex = sweep.to_texels(feats.to(dev))
            pose = sweep.relative_poses(proj[0].to(dev))
            outs.append(sweep.cost_volume(tex, pose, hyps[0].to(dev), sweep.AGG_VARIANCE).cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_full_size_regression_against_cuda_aten():
    d, h, w = 384, 688, 464
    logits = synth.planted_logits(d, h, w, seed=1).to(DEV)
    hyps = synth.uniform_hypotheses(400.0, 600.0, d, device=DEV)
    depth, cf, idx = sweep_torch.regress_maxprob(logits.unsqueeze(0), hyps.unsqueeze(0))
    r = sweep.depth_regress(logits, hyps)
    assert _depth_err(r["depth"], depth[0]) < DEPTH_TOL
    assert (r["index"].long() == idx[0]).float().mean() >= 0.999
    # shift invariance of the softmax: adding a constant to every logit changes nothing
    r2 = sweep.depth_regress(logits + 3.0, hyps)
    assert _depth_err(r2["depth"], r["depth"]) < 1e-5


# ------------------------------------------------------- pair-confidence resize between AdaMVS stages (adamvs.py:291-302)
@pytest.mark.parametrize("n,hi,wi,ho,wo", [(4, 58, 86, 116, 172), (4, 116, 172, 232, 344), (3, 17, 23, 40, 31), (2, 40, 31, 17, 23),
                                          (1, 8, 8, 8, 8), (4, 1, 5, 3, 10)])
def test_resize_bilinear_matches_aten(n, hi, wi, ho, wo):
    g = torch.Generator().manual_seed(n * 1000 + hi)
    maps = torch.rand(2, n, hi, wi, generator=g)
    want = torch.nn.functional.interpolate(maps, [ho, wo], mode="bilinear", align_corners=False)       # CPU ATen
    want_cuda = torch.nn.functional.interpolate(maps.to(DEV), [ho, wo], mode="bilinear", align_corners=False)
    got = sweep.resize_bilinear(maps.to(DEV), (ho, wo))
    assert got.shape == want.shape
    torch.testing.assert_close(got.cpu(), want, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(got, want_cuda, rtol=1e-6, atol=1e-6)


def test_resize_pair_weights_stacks_and_falls_through():
    g = torch.Generator().manual_seed(5)
    maps = [torch.rand(1, 1, 20, 30, generator=g).to(DEV) for _ in range(6)]         # an over-long list: V-1 = 4 are used
    resized, weights = depthnets._resize_pair_weights(maps, 4, 40, 60)
    want = [torch.nn.functional.interpolate(m, [40, 60], mode="bilinear", align_corners=False) for m in maps[:4]]
    assert weights.shape == (1, 4, 40, 60) and len(resized) == 4
    for r, w in zip(resized, want):
        torch.testing.assert_close(r, w, rtol=1e-6, atol=1e-6)
    same, wsame = depthnets._resize_pair_weights(maps, 4, 20, 30)                     # stage 1: same size, values untouched
    assert all(torch.equal(a, b) for a, b in zip(same, maps[:4])) and torch.equal(wsame, torch.cat(maps[:4], 1))


# ------------------------------------------------------- texel pool: views named by slot (D3dCostVolumeArgs.texel_slots)
_POOL_CASES = [
    # v, c, d, h, w, mode, kwargs, per-pixel hypotheses
    (5, 32, 136, 16, 32, sweep.AGG_VARIANCE, {}, False),                 # sweep_quad, one-block re-fetch, shifted footprints
    (5, 32, 12, 16, 32, sweep.AGG_VARIANCE, {}, True),                   # sweep_quad, two-phase re-fetch
    (5, 32, 136, 16, 32, sweep.AGG_GROUP_CORR, {"groups": 8}, False),    # corner dot products
    (5, 32, 12, 16, 32, sweep.AGG_GROUP_CORR, {"groups": 32}, False),    # coefficient cache
    (5, 32, 12, 16, 32, sweep.AGG_PAIR_MEAN, {}, True),
    (4, 32, 12, 16, 32, sweep.AGG_WEIGHTED_PRODUCT, {"weights": True}, True),
    (5, 16, 12, 16, 32, sweep.AGG_WEIGHTED_PRODUCT, {"weights": True}, True),   # four lanes per pixel
    (3, 16, 12, 16, 32, sweep.AGG_VARIANCE, {}, False),
    (5, 8, 8, 16, 32, sweep.AGG_WEIGHTED_PRODUCT, {"weights": True}, True),     # sweep_acc
    (3, 8, 12, 16, 32, sweep.AGG_VARIANCE, {}, False),                   # sweep_direct
    (7, 8, 6, 12, 20, sweep.AGG_VARIANCE, {}, False),                    # sweep_base (six source views)
    (2, 4, 6, 12, 20, sweep.AGG_WARP, {}, False),                        # sweep_base, the plain warp
    (3, 64, 6, 16, 32, sweep.AGG_VARIANCE, {}, False),                   # sweep_lean's shape: refuses the pool, sweep_base takes it
]


@pytest.mark.parametrize("v,c,d,h,w,mode,kw,perpix", _POOL_CASES)
def test_texel_pool_slots_match_dense_block(v, c, d, h, w, mode, kw, perpix):
    """The views of a sweep named as slots of a per-image texel pool: bit-identical to the dense [V,H,W,C] block wherever the
    same kernel serves both (every production kernel), and within the volume tolerance where the pool falls through to
    another kernel."""
    rig, proj, feats, hyps = _scene(v, c, d, h, w, seed=11, perpixel=perpix)
    tex = sweep.to_texels(feats.to(DEV))
    pose = sweep.relative_poses(proj[0].to(DEV))
    hy = hyps[0].to(DEV).contiguous()
    kw = dict(kw)
    if kw.pop("weights", False):
        kw["weights"] = torch.rand(v - 1, h, w, generator=torch.Generator().manual_seed(2)).to(DEV)
    dense = sweep.cost_volume(tex, pose, hy, mode, **kw)
    slots = [6, 2, 7, 0, 4, 9, 1][:v]                                    # scattered, out of order, the reference not first
    pool = torch.full((10, h, w, c), float("nan"), device=DEV)
    for i, s in enumerate(slots):
        pool[s].copy_(tex[i])
    pooled = sweep.cost_volume(pool, pose, hy, mode, view_slots=slots, **kw)
    assert not torch.isnan(pooled).any()
    if c == 64:
        assert rel_norm_err(pooled.cpu(), dense.cpu()) < VOL_TOL
    else:
        assert torch.equal(pooled, dense)
    # the identity naming of a pool that IS the dense block takes the dense path
    assert torch.equal(sweep.cost_volume(tex, pose, hy, mode, view_slots=list(range(v)), **kw), dense)


def test_texel_pool_slot_validation():
    rig, proj, feats, hyps = _scene(3, 8, 4, 12, 20)
    tex = sweep.to_texels(feats.to(DEV))
    pose = sweep.relative_poses(proj[0].to(DEV))
    with pytest.raises(ValueError, match="view_slots"):
        sweep.cost_volume(tex, pose, hyps[0].to(DEV), view_slots=[0, 1, 3])
    with pytest.raises(ValueError, match="pose"):
        sweep.cost_volume(tex, pose, hyps[0].to(DEV), view_slots=[0, 1])


def test_depthnet_texel_pool_lays_each_image_out_once():
    """depthnets.TEXEL_POOL (texel_pool.TexelPool): reference views that receive the SAME feature-map tensors (FeatureCache)
    lay an image out once; volumes are bit-identical to the dense-block path, a modified or replaced tensor is laid out again."""
    from deep3d_aerial_b200 import _lib
    from deep3d_aerial_b200.texel_pool import TexelPool

    v, c, d, h, w, n_img = 5, 32, 8, 16, 32, 8
    rig, proj, _, hyps = _scene(v, c, d, h, w, seed=5)
    g = torch.Generator().manual_seed(17)
    images = [torch.randn(1, c, h, w, generator=g).to(DEV) for _ in range(n_img)]
    pr = torch.unbind(proj.to(DEV), 1)
    hy = hyps.to(DEV)
    windows = [[i] + [j for j in range(i - 2, i + 3) if j != i] for i in range(2, n_img - 2)]
    dense = [depthnets._volume([images[j] for j in win], pr, hy, sweep.AGG_VARIANCE) for win in windows]
    depthnets.TEXEL_POOL = TexelPool(6)                                       # smaller than the walk: slots are recycled
    try:
        launches = []
        for win, want in zip(windows, dense):
            n0 = _lib.launch_count()
            got = depthnets._volume([images[j] for j in win], pr, hy, sweep.AGG_VARIANCE)
            launches.append(_lib.launch_count() - n0)
            assert torch.equal(got, want)
        base = launches[0] - v                                                 # everything but the relayouts
        assert launches == [base + v] + [base + 1] * (len(windows) - 1), launches
        images[4].mul_(2.0)                                                   # an in-place change bumps the version counter
        win = windows[-1]
        n0 = _lib.launch_count()
        got = depthnets._volume([images[j] for j in win], pr, hy, sweep.AGG_VARIANCE)
        assert _lib.launch_count() - n0 == base + 1
        assert torch.equal(got, depthnets._volume([images[j].clone() for j in win], pr, hy, sweep.AGG_VARIANCE))
        st = depthnets.TEXEL_POOL.stats()
        assert st["hits"] + st["misses"] == v * (len(windows) + 2)
        # a batch of two goes the dense way (its per-item slices are new tensors every call)
        two = [torch.cat([images[j], images[j]], 0) for j in win]
        pr2 = torch.unbind(proj.to(DEV).repeat(2, 1, 1, 1), 1)
        both = depthnets._volume(two, pr2, hy.repeat(2, 1), sweep.AGG_VARIANCE)
        assert torch.equal(both[0], got[0]) and torch.equal(both[1], got[0])
    finally:
        depthnets.TEXEL_POOL = None


@pytest.mark.parametrize("c,d,scale,mode", [(32, 16, 4, sweep.AGG_VARIANCE), (16, 8, 2, sweep.AGG_WEIGHTED_PRODUCT),
                                            (8, 8, 1, sweep.AGG_WEIGHTED_PRODUCT)])
def test_texel_pool_at_full_size_is_bit_identical_to_the_dense_block(c, d, scale, mode):
    """BASELINE.json's full image sizes (the three cascade resolutions of 1856x2752): views in scattered slots of a 7-slot
    pool -- texel offsets of up to 30 M texels / 1 GB -- against the dense block, bit for bit."""
    v = 5
    rig = synth.make_rig(num_views=v)
    h, w = 2752 // scale, 1856 // scale
    g = torch.Generator().manual_seed(23)
    pose = sweep.relative_poses(torch.from_numpy(rig.proj(scale)).to(DEV))
    if mode == sweep.AGG_VARIANCE:
        hyps = synth.uniform_hypotheses(rig.dmin, rig.dmax, 384)[100:100 + d].to(DEV).contiguous()
        kw = {}
    else:
        cur = synth.smooth_depth_map(rig, h, w, seed=2).to(DEV)
        hyps = synth.per_pixel_hypotheses(cur, d, scale * (rig.dmax - rig.dmin) / 384).contiguous()
        kw = {"weights": torch.rand(v - 1, h, w, generator=g).to(DEV), "plane_major": True}
    slots = [5, 0, 6, 2, 3]
    pool = torch.full((7, h, w, c), float("nan"), device=DEV)
    tex = torch.empty((v, h, w, c), device=DEV)
    for i, s in enumerate(slots):
        fmap = torch.randn(c, h, w, generator=g).to(DEV)
        sweep.to_texels([fmap], out=tex[i:i + 1])
        pool[s].copy_(tex[i])
    dense = sweep.cost_volume(tex, pose, hyps, mode, **kw)
    pooled = sweep.cost_volume(pool, pose, hyps, mode, view_slots=slots, **kw)
    assert torch.equal(pooled, dense) and bool(torch.isfinite(pooled).all())
