"""the input contract and the on-disk formats, against golden vectors made by the LIVE
reference (`oracle/make_golden_formats.py`; `datasets/data_io.py`, `preprocess.py`, `cas_normal_eval.py`).
Host-side code: everything here runs on CPU; all comparisons are bit-exact (`np.array_equal`)."""
import json
import os
import shutil
import types

import numpy as np
import pytest

from deep3d_aerial_b200 import dataset, formats

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WS = os.path.join(GOLDEN, "workspace_tiny")

ARGSETS = {   # oracle/make_golden_formats.py:ARGSETS
    "plain": dict(min_interval=0.1, interval_scale=1.0, numdepth=48, resize_scale=1.0, sample_scale=1.0,
                  max_h=64, max_w=96),
    "scaled": dict(min_interval=0.1, interval_scale=1.0, numdepth=32, resize_scale=0.5, sample_scale=0.25,
                   max_h=96, max_w=128),
}


@pytest.fixture
def workspace(tmp_path):
    dst = str(tmp_path / "ws")
    shutil.copytree(WS, dst)
    with open(os.path.join(WS, "image_path.template.txt")) as f:
        text = f.read().replace("{ROOT}", dst)
    with open(os.path.join(dst, "image_path.txt"), "w") as f:
        f.write(text)
    return dst


def _flatten(prefix, item, out):
    for k, v in item.items():
        if isinstance(v, dict):
            _flatten(prefix + k + ".", v, out)
        elif isinstance(v, np.ndarray):
            out[prefix + k] = v
        elif isinstance(v, list):
            out[prefix + k] = np.array([str(x) for x in v])


@pytest.mark.parametrize("tag", sorted(ARGSETS))
@pytest.mark.parametrize("view_num,norm", [(3, "mean"), (2, "standard")])
def test_dataset_items_match_the_live_reference(workspace, tag, view_num, norm):
    with np.load(os.path.join(GOLDEN, "formats_dataset.npz")) as z:
        gold = {k: z[k] for k in z.files}
    ds = dataset.MVSDataset(workspace, "val", view_num, norm, types.SimpleNamespace(**ARGSETS[tag]))
    pre = "%s.v%d." % (tag, view_num)
    assert len(ds) == int(gold[pre + "len"])
    assert ds.sample_list == json.loads(str(gold[pre + "samples"]))
    checked = 0
    for idx in range(len(ds)):
        got = {}
        _flatten(pre + "%d." % idx, ds[idx], got)
        want = {k: v for k, v in gold.items() if k.startswith(pre + "%d." % idx)}
        assert sorted(got) == sorted(want)
        for k, v in want.items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
            assert np.array_equal(got[k], v), k          # bit-exact, float32 projection pyramids included
            checked += 1
    assert checked == 11 * len(ds)


def test_view_pairs_pad_short_lists_and_drop_empty_ones(workspace):
    pairs = formats.read_view_pair_text(os.path.join(workspace, "viewpair.txt"), 3)
    assert pairs == [[0, 1, 2, 3], [1, 0, 2, 0], [3, 2, 2, 2]]       # view 2 has no sources: dropped
    cams = formats.read_cameras_text(os.path.join(workspace, "cameras.txt"))
    assert sorted(cams) == [1, 2] and cams[2].size == [128, 96] and cams[2].distortion[0] == 1e-3
    photos = formats.read_images_text(os.path.join(workspace, "images.txt"))
    assert photos[3].camera_id == 2 and photos[3].name == "b_003.png" and photos[3].rotation_matrix.shape == (3, 3)
    paths, names = formats.read_images_path_text(os.path.join(workspace, "image_path.txt"))
    assert names[1] == "a_001.png" and paths[1].endswith("/img/a_001.png")


def test_pfm_and_camera_files_are_byte_identical_to_the_reference(tmp_path):
    with np.load(os.path.join(GOLDEN, "formats_files.npz")) as z:
        g = {k: z[k] for k in z.files}
    for tag in ("grey", "colour", "grey1"):
        p = str(tmp_path / (tag + ".pfm"))
        formats.save_pfm_utf8(p, g["pfm.%s.in" % tag])
        assert np.array_equal(np.fromfile(p, dtype=np.uint8), g["pfm.%s.bytes" % tag]), tag
        if tag != "grey1":
            back, scale = formats.load_pfm_utf8(p)
            assert np.array_equal(back, g["pfm.%s.back" % tag]) and scale == float(g["pfm.%s.scale" % tag])
            assert np.array_equal(back, g["pfm.%s.in" % tag])            # round trip
    p = str(tmp_path / "cam.txt")
    formats.write_red_cam(p, g["cam.in"], list(g["cam.location"]), "/data/images/a_000.png")
    assert np.array_equal(np.fromfile(p, dtype=np.uint8), g["cam.bytes"])
    cam, location, ref_path = formats.read_red_cam(p)
    assert np.array_equal(cam, g["cam.in"]) and location == list(g["cam.location"])
    assert ref_path == "/data/images/a_000.png"
    for mode in ("standard", "mean", "vit"):
        assert np.array_equal(dataset.center_image(g["center.in"], mode=mode), g["center." + mode]), mode


def test_pfm_error_behaviour(tmp_path):
    with pytest.raises(Exception, match="float32"):
        formats.save_pfm_utf8(str(tmp_path / "x.pfm"), np.zeros((2, 2), dtype=np.float64))
    with pytest.raises(Exception, match="dimensions"):
        formats.save_pfm_utf8(str(tmp_path / "x.pfm"), np.zeros((2, 2, 2), dtype=np.float32))
    bad = tmp_path / "bad.pfm"
    bad.write_bytes(b"P6\n2 2\n-1.0\n")
    with pytest.raises(Exception, match="Not a PFM"):
        formats.load_pfm_utf8(str(bad))
    bad.write_bytes(b"Pf\n2x2\n-1.0\n")
    with pytest.raises(Exception, match="Malformed"):
        formats.load_pfm_utf8(str(bad))
    big = tmp_path / "big.pfm"                                   # big-endian files (positive scale) load too
    big.write_bytes(b"Pf\n2 1\n1.0\n" + np.array([1.5, -2.0], dtype=">f4").tobytes())
    data, scale = formats.load_pfm_utf8(str(big))
    assert data.tolist() == [[1.5, -2.0]] and scale == 1.0
    with pytest.raises(Exception, match="Not implemented"):
        dataset.center_image(np.zeros((2, 2, 3), np.uint8), mode="nope")


def test_collate_and_save_view_outputs(workspace, tmp_path):
    import torch
    ds = dataset.MVSDataset(workspace, "val", 3, "mean", types.SimpleNamespace(**ARGSETS["plain"]))
    sample = dataset.collate(ds[0])
    assert sample["imgs"].shape == (1, 3, 3, 64, 96) and sample["imgs"].dtype == torch.float32
    assert sample["proj_matrices"]["stage1"].shape == (1, 3, 4, 4)
    assert sample["depth_values"].shape == (1, 2) and sample["outlocation"][3] == ["a_000.png"]
    h, w = 64, 96
    depth = np.linspace(400, 640, h * w, dtype=np.float32).reshape(1, h, w)
    depth[0, 3, 5] = np.inf
    prob = np.full((1, h, w), 0.5, dtype=np.float32)
    location = [x[0] for x in sample["outlocation"]]
    out = dataset.save_view_outputs(str(tmp_path / "mvs"), depth, prob, sample["outcam"][0].numpy(), location,
                                    sample["ref_image_path"][0], display=True)
    back, _ = formats.load_pfm_utf8(out["depth"])
    assert np.array_equal(back, depth[0])
    assert np.array_equal(formats.load_pfm_utf8(out["prob"])[0], prob[0])
    cam, loc, ref_path = formats.read_red_cam(out["cam"])
    assert np.array_equal(cam, ds[0]["outcam"]) and loc == ["96", "64", "0", "a_000.png"]
    assert os.path.basename(out["depth"]) == "a_000_init.pfm" and os.path.isfile(out["depth_png"])


@pytest.mark.gpu
def test_predict_driver_writes_the_reference_outputs(workspace, tmp_path):
    """The b1 surface end to end on the B200: workspace text files -> dataset tensors -> a cascade stand-in built
    on this engine's DepthNet (the reference's CNNs are not on the GPU box) -> PFM / camera files on disk."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    from deep3d_aerial_b200 import depthnets, module as d3d_module, predict
    from oracle import standins, sweep_torch

    class TinyFeatureNet(nn.Module):                       # 8 fixed "feature" channels at 1/4 resolution
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(0)
            self.register_buffer("kernel", torch.randn(8, 3, 3, 3, generator=g) * 0.3)
            self.calls = 0

        def forward(self, img):
            self.calls += 1
            return F.avg_pool2d(F.conv2d(img, self.kernel, padding=1), 4)

    class TinyCascade(nn.Module):                          # one stage, the reference's walk over imgs[:, v]
        def __init__(self, num_depth):
            super().__init__()
            self.num_depth = num_depth
            self.depthnet = depthnets.DepthNet()
            self.feature = TinyFeatureNet()

        def features(self, imgs):
            return [self.feature(imgs[:, v]) for v in range(imgs.shape[1])]

        def forward(self, imgs, proj_matrices, depth_values):
            feats = self.features(imgs)
            proj = proj_matrices["stage1"]
            h, w = feats[0].shape[2:]
            hyps = d3d_module.get_depth_range_samples(depth_values, self.num_depth, 1.0, imgs.device, imgs.dtype,
                                                            [imgs.shape[0], h, w])
            out = self.depthnet(feats, proj, hyps, self.num_depth, standins.reg3d)
            return {"depth": out["depth"], "photometric_confidence": out["photometric_confidence"], "hyps": hyps,
                    "feats": feats}

    out_dir = str(tmp_path / "dense")
    args = predict.build_parser().parse_args(["--data_folder", workspace, "--output_folder", out_dir, "--view_num", "3",
                                              "--numdepth", "16", "--max_h", "64", "--max_w", "96",
                                              "--feature_cache", "8"])
    model = TinyCascade(16)
    written = predict.predict_depth(args, model=model)
    assert len(written) == 3
    assert model.feature.calls == 4          # 3 reference views x 3 images, 4 distinct images: each encoded once (row f4)
    model.feature.calls = 0
    ds = dataset.MVSDataset(workspace, "val", 3, "mean", args)
    for idx, paths in enumerate(written):
        depth, _ = formats.load_pfm_utf8(paths["depth"])
        prob, _ = formats.load_pfm_utf8(paths["prob"])
        assert depth.shape == (16, 24) and prob.shape == (16, 24) and np.isfinite(depth).all()
        item = ds[idx]
        assert depth.min() >= item["depth_values"][0] - 1e-3 and depth.max() <= item["depth_values"][1] + 1e-3
        cam, loc, _ = formats.read_red_cam(paths["cam"])
        assert np.array_equal(cam, item["outcam"]) and loc == item["outlocation"]
        # the same view through the oracle's restatement of DepthNet.forward (CPU ATen) on the same tensors
        sample = dataset.collate(item)
        with torch.no_grad():
            feats = [f.cpu() for f in model.features(sample["imgs"].cuda())]
            hyps = sweep_torch.depth_range_samples(sample["depth_values"], 16, 1.0, (1, 16, 24))
            volume = sweep_torch.variance_volume(feats, sample["proj_matrices"]["stage1"], hyps)
            want_depth, want_conf, _ = sweep_torch.regress_window4(standins.reg3d(volume).squeeze(1), hyps)
        want_depth, want_conf = want_depth[0].numpy(), want_conf[0].numpy()
        assert np.abs(depth - want_depth).max() / np.abs(want_depth).max() < 1e-3          # north_star: depth 1e-3
        assert np.abs(prob - want_conf).max() < 1e-3


def test_mvs_inference_drop_in_builds_the_predict_command(tmp_path, capsys, monkeypatch):
    """`mvs/mvs_dl.py:MVS_Inference` with the same signature: default checkpoint lookup, the command line of
    predict.py (same option names), launched through os.system with its status ignored."""
    from deep3d_aerial_b200 import mvs_dl, predict
    ref = tmp_path / "mvs_cas"
    ck = ref / "checkpoints" / "adamvs" / "whu_omvs"
    ck.mkdir(parents=True)
    (ck / "notes.txt").write_text("x")
    (ck / "model_000019.ckpt").write_text("x")
    mi = mvs_dl.MVS_Inference(2752, 1856, 5, 384, 0.1, 'AdaMVS', None, False, reference_root=str(ref))
    cmd = mi.command("/ws/export", "/ws/dense/MVS")
    assert "-m deep3d_aerial_b200.predict" in cmd and "--loadckpt=" + str(ck / "model_000019.ckpt") in cmd
    argv = cmd.split(" -m deep3d_aerial_b200.predict ")[1].split()
    args = predict.build_parser().parse_args(argv)                       # every option is one predict accepts
    assert (args.model, args.view_num, args.numdepth, args.max_w, args.max_h) == ("adamvs", 5, 384, 2752, 1856)
    assert args.data_folder == "/ws/export" and args.output_folder == "/ws/dense/MVS" and args.display == "False"
    multi = mvs_dl.MVS_Inference(2752, 1856, pretrain_weight="w.ckpt", gpus=8, feature_cache=64).command("a", "b")
    assert "--nproc-per-node 8" in multi and "--master-addr 127.0.0.1" in multi and "--partition=contiguous" in multi
    with pytest.raises(Exception, match="Not implemented"):
        mvs_dl.MVS_Inference(1, 1, model_type="colmap").command("a", "b")
    ran = []
    monkeypatch.setattr(mvs_dl.os, "system", lambda c: ran.append(c) or 1)   # a failing launch is ignored, as upstream
    mi.run("/ws/export", str(tmp_path / "dense" / "MVS"))
    assert ran == [mi.command("/ws/export", str(tmp_path / "dense" / "MVS"))] and (tmp_path / "dense").is_dir()
