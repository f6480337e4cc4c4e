"""the depth-map fusion consistency check (`fuse/consistency_check_n.py:29-138`,
`fuse/fusion_3d_normal.py:436-541`).

CPU: the numpy oracle against golden vectors made by the LIVE reference (`oracle/make_golden_fuse.py`), the
geometry block, argument validation.  GPU (`-m gpu`): `d3d_consistency_fuse` through the C ABI against the golden
vectors and against the oracle on seeded scenes (invalid depths, projections that leave the source maps, source
maps of another size, 1..16 source views), the `ConsistencyChecker.check` drop-in, and size-independent
properties at the full 1856 x 2752 depth-map size.

Tolerances: masks, counts and zeroed pixels are compared exactly; float32 outputs to 1 ulp-level relative error
(2e-6 of the scene scale) because the order of the three products inside numpy's fp32 matmul (the normals'
rotation) is BLAS's business.  A mask may only differ where the oracle itself sits within 1e-6 of a threshold.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from deep3d_aerial_b200 import _lib, fusion, synth
from oracle import fuse_np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD_SCENE = dict(num_src=3, height=40, width=56, seed=5, src_pad=12, invalid=0.0)      # oracle/make_golden_fuse.py
TH = dict(position_threshold=1.0, depth_threshold=0.01, normal_cos=fuse_np.normal_threshold_cos(10.0),
          confidence_threshold=0.2)


def _gold(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_matches_the_live_reference_bit_for_bit():
    sc = synth.fusion_scene(**GOLD_SCENE)
    d, n, k, e, prob = sc["ref"]
    pair, view = _gold("fuse_pair"), _gold("fuse_view")
    for s, (ds, ns, ks, es) in enumerate(sc["src"]):
        keep = ds.copy()
        got = fuse_np.check(d, n, k, e, ds, ns, ks, es, prob, **TH)
        assert np.array_equal(ds, keep)                      # the caller's map is not modified
        for name, arr in zip(("mask", "depth_reprojected", "depth_src_out", "xyz", "angle"), got):
            want = pair["%d.%s" % (s, name)]
            assert arr.dtype == want.dtype and np.array_equal(arr, want), (s, name)
    r = fuse_np.fuse_view(d, n, k, e, prob, sc["src"], min_consistent=3, **TH)
    assert np.array_equal(r["count"], view["count"]) and np.array_equal(r["xyz"], view["xyz"])
    assert np.array_equal(r["final_mask"], view["final_mask.3"])
    assert 0.2 < r["final_mask"].mean() < 0.9               # the scene exercises both outcomes


def test_oracle_wraps_out_of_bounds_indices_like_cupy():
    sc = synth.fusion_scene(num_src=2, height=24, width=32, seed=2, shift=0.7)
    d, n, k, e, prob = sc["ref"]
    ds, ns, ks, es = sc["src"][0]
    mask, rep, removed, xyz, angle = fuse_np.check(d, n, k, e, ds, ns, ks, es, prob, **TH)   # numpy alone would raise
    assert mask.shape == d.shape and np.isfinite(rep).all()
    assert ((removed == 0) | (removed == ds)).all()


def test_pair_geometry_layout():
    sc = synth.fusion_scene(num_src=2, height=8, width=8, seed=0)
    _, _, k, e, _ = sc["ref"]
    g = fusion.pair_geometry(k, e, [v[2] for v in sc["src"]], [v[3] for v in sc["src"]])
    assert g.shape == (3, 64) and g.dtype == np.float64
    pad = lambda m: np.pad(np.asarray(m, dtype=np.float64), ((0, 0), (0, 4 - m.shape[1]))).reshape(-1)   # noqa: E731
    assert np.array_equal(g[0, 24:36], pad(k))
    assert np.array_equal(g[0, 36:52], pad(np.linalg.inv(e)))                                  # float32 inverse, widened
    assert np.array_equal(g[2, 0:12], pad(np.matmul(sc["src"][1][3], np.linalg.inv(e))[:3]))
    assert np.array_equal(g[1, 52:64], pad(np.linalg.inv(sc["src"][0][3][:3, :3])))
    assert (g[0, 3:12:4] == 0).all() and (g[:, 27:36:4] == 0).all()       # the padding of the 3x3 matrices


def test_fuse_validation_needs_no_gpu():
    lib = _lib.load()
    assert lib.d3d_consistency_fuse(None, None) == _lib.ERR_BAD_ARGUMENT
    a = _lib.FuseArgs()
    a.struct_size = 12
    assert lib.d3d_consistency_fuse(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT and b"struct_size" in lib.d3d_last_error()
    a.struct_size = C.sizeof(a)
    a.num_src, a.height, a.width, a.src_height, a.src_width = 17, 4, 4, 4, 4
    assert lib.d3d_consistency_fuse(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT and b"num_src" in lib.d3d_last_error()
    a.num_src = 1
    assert lib.d3d_consistency_fuse(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT       # NULL maps
    a.depth_ref = a.normal_ref = a.prob_ref = a.geometry = 256
    a.depth_src[0] = a.normal_src[0] = 512
    a.depth_src_out[0] = 512
    assert lib.d3d_consistency_fuse(C.byref(a), None) == _lib.ERR_BAD_ARGUMENT and b"aliases" in lib.d3d_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fusion.fuse_view(torch.zeros(4, 4), torch.zeros(4, 4, 3), torch.zeros(4, 4), torch.zeros(2, 64, dtype=torch.float64),
                         [torch.zeros(4, 4)], [torch.zeros(4, 4, 3)])
    with pytest.raises(ValueError, match="source views"):
        fusion.fuse_view(torch.zeros(4, 4), torch.zeros(4, 4, 3), torch.zeros(4, 4), torch.zeros(2, 64, dtype=torch.float64),
                         [], [])


# ------------------------------------------------------------------------------------------------ GPU
def _run_gpu(sc, min_consistent=3, per_source=True):
    dev = torch.device("cuda", 0)
    d, n, k, e, prob = sc["ref"]
    geom = torch.from_numpy(fusion.pair_geometry(k, e, [v[2] for v in sc["src"]], [v[3] for v in sc["src"]])).to(dev)
    up = lambda x: torch.from_numpy(x).to(dev)                       # noqa: E731
    return fusion.fuse_view(up(d), up(n), up(prob), geom, [up(v[0]) for v in sc["src"]], [up(v[1]) for v in sc["src"]],
                            position_threshold=TH["position_threshold"], depth_threshold=TH["depth_threshold"],
                            normal_threshold_cos=TH["normal_cos"], confidence_threshold=TH["confidence_threshold"],
                            min_consistent=min_consistent, per_source=per_source)


def _close(got, want, scale):
    return np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= 2e-6 * scale


def _compare_with_oracle(sc, r, min_consistent):
    d, n, k, e, prob = sc["ref"]
    want = fuse_np.fuse_view(d, n, k, e, prob, sc["src"], min_consistent=min_consistent, **TH)
    masks = r["masks"].cpu().numpy()
    assert np.array_equal(masks, want["masks"])
    assert np.array_equal(r["count"].cpu().numpy(), want["count"])
    assert np.array_equal(r["final_mask"].cpu().numpy(), want["final_mask"])
    assert np.array_equal(r["depth_ref_filtered"].cpu().numpy(), want["depth_ref_filtered"])
    scale = float(np.abs(want["xyz"]).max())
    assert _close(r["xyz"].cpu().numpy(), want["xyz"], scale)
    for s, (ds, ns, ks, es) in enumerate(sc["src"]):
        assert np.array_equal(r["depth_src_out"][s].cpu().numpy(), want["depth_src_out"][s]), s
        m, rep, _, xyz, angle = fuse_np.check(d, n, k, e, ds, ns, ks, es, prob, **TH)
        assert _close(r["depth_reprojected"][s].cpu().numpy(), rep, float(d.max())), s
        assert _close(r["xyz_world_src"][s].cpu().numpy(), xyz, scale), s
        assert _close(r["angle_conf"][s].cpu().numpy(), angle[0], 1.0), s
    return want


@pytest.mark.gpu
def test_kernel_matches_the_live_reference_golden():
    sc = synth.fusion_scene(**GOLD_SCENE)
    r = _run_gpu(sc, min_consistent=3)
    pair, view = _gold("fuse_pair"), _gold("fuse_view")
    assert np.array_equal(r["count"].cpu().numpy(), view["count"])
    assert np.array_equal(r["final_mask"].cpu().numpy(), view["final_mask.3"])
    assert _close(r["xyz"].cpu().numpy(), view["xyz"], float(np.abs(view["xyz"]).max()))
    for s in range(3):
        assert np.array_equal(r["masks"][s].cpu().numpy(), pair["%d.mask" % s])
        assert np.array_equal(r["depth_src_out"][s].cpu().numpy(), pair["%d.depth_src_out" % s])
        assert _close(r["depth_reprojected"][s].cpu().numpy(), pair["%d.depth_reprojected" % s], 600.0)
        assert _close(r["xyz_world_src"][s].cpu().numpy(), pair["%d.xyz" % s], 600.0)
        assert _close(r["angle_conf"][s].cpu().numpy(), pair["%d.angle" % s][0], 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("kw,min_consistent", [
    (dict(num_src=4, height=96, width=128, seed=1), 3),                      # invalid depths, outliers
    (dict(num_src=1, height=33, width=47, seed=2), 2),                       # one source, ragged tail of the grid
    (dict(num_src=10, height=64, width=80, seed=3), 4),                      # the reference's fusion_num
    (dict(num_src=16, height=40, width=40, seed=4), 5),                      # the ABI's maximum
    (dict(num_src=3, height=48, width=64, seed=5, shift=0.7), 2),            # projections leave the maps: wrap-around
    (dict(num_src=3, height=48, width=64, seed=6, src_pad=9), 3),            # source maps larger than the reference's
    (dict(num_src=2, height=48, width=64, seed=7, invalid=1.0), 1),          # nothing valid
])
def test_kernel_matches_oracle(kw, min_consistent):
    sc = synth.fusion_scene(**kw)
    want = _compare_with_oracle(sc, _run_gpu(sc, min_consistent=min_consistent), min_consistent)
    if kw.get("invalid") == 1.0:
        assert not want["masks"].any()


@pytest.mark.gpu
def test_consistency_checker_drop_in_matches_the_live_reference_golden(capsys):
    sc = synth.fusion_scene(**GOLD_SCENE)
    d, n, k, e, prob = sc["ref"]
    chk = fusion.ConsistencyChecker(1.0, 0.01, 10.0, 0.2, implement="cupy")
    assert "normal_th:0.98480" in capsys.readouterr().out
    pair = _gold("fuse_pair")
    ds, ns, ks, es = sc["src"][1]
    keep = ds.copy()
    mask, rep, removed, xyz, angle = chk.check(d, n, k, e, ds, ns, ks, es, prob)
    assert np.array_equal(ds, keep)
    assert mask.dtype == np.bool_ and np.array_equal(mask, pair["1.mask"])
    assert np.array_equal(removed, pair["1.depth_src_out"])
    assert angle.shape == (3,) + d.shape and _close(angle, pair["1.angle"], 1.0)
    assert _close(rep, pair["1.depth_reprojected"], 600.0) and _close(xyz, pair["1.xyz"], 600.0)
    with pytest.raises(AssertionError):
        fusion.ConsistencyChecker(1.0, 0.01, 10.0, 0.2, implement="jax")


@pytest.mark.gpu
def test_full_size_properties():
    """1856 x 2752 depth maps, 10 source views (fusion_3d_normal.py's --fusion_num): properties that need no oracle."""
    sc = synth.fusion_scene(num_src=10, height=2752, width=1856, seed=11, focal=4000.0)
    r = _run_gpu(sc, min_consistent=4, per_source=False)
    masks, count = r["masks"], r["count"]
    assert torch.equal(count, 1 + masks.sum(0, dtype=torch.int32))
    assert torch.equal(r["final_mask"], count >= 4)
    d = torch.from_numpy(sc["ref"][0]).cuda()
    assert torch.equal(r["depth_ref_filtered"], torch.where(r["final_mask"], d, torch.zeros_like(d)))
    assert not masks[:, d <= 0].any() and not masks[:, torch.from_numpy(sc["ref"][4]).cuda() <= 0.2].any()
    assert 0.2 < float(r["final_mask"].float().mean()) < 0.9
    for s in range(10):
        src = torch.from_numpy(sc["src"][s][0]).cuda()
        out = r["depth_src_out"][s]
        changed = out != src
        assert (out[changed] == 0).all()                                       # pixels are only ever zeroed ...
        assert int(changed.sum()) <= int(masks[s].sum())                       # ... at most one per consistent pixel
        assert int(changed.sum()) > 0.5 * int(masks[s].sum())
    # a consistent pixel's fused point stays within the depth tolerance of the reference pixel's own world point
    xyz = r["xyz"]
    assert torch.isfinite(xyz).all()
    # deterministic: a second launch gives the same bytes
    r2 = _run_gpu(sc, min_consistent=4, per_source=False)
    assert torch.equal(r2["xyz"], xyz) and torch.equal(r2["masks"], masks)
    # the row-subset oracle check: the first 64 rows of the reference view against the numpy oracle
    rows = 64
    dref, nref, k, e, prob = sc["ref"]
    sub = fuse_np.fuse_view(dref[:rows], nref[:rows], k, e, prob[:rows], sc["src"][:3], min_consistent=2, **TH)
    assert np.array_equal(masks[:3, :rows].cpu().numpy(), sub["masks"])


def _block_inputs(seed=3, h=40, w=48):
    sc = synth.fusion_scene(num_src=3, height=h, width=w, seed=seed)
    views = [sc["ref"][:4]] + sc["src"]
    rng = np.random.default_rng(seed)
    depths = {i: v[0].copy() for i, v in enumerate(views)}
    normals = {i: v[1] for i, v in enumerate(views)}
    conf = {i: rng.random((h, w)).astype(np.float32) for i in range(4)}
    return depths, normals, conf, {i: v[2] for i, v in enumerate(views)}, {i: v[3] for i, v in enumerate(views)}


VIEW_LIST = [{"ref": 0, "src": [1, 2, 3, 1, 1]},      # a padded list: source 1 three times (fusion_3d_normal.py:243-246)
             {"ref": 1, "src": [0, 2]},               # sees what view 0 left of maps 0 and 2
             {"ref": 7, "src": [0]},                  # no depth map for this view: skipped
             {"ref": 2, "src": [9]},                  # no usable source: skipped
             {"ref": 3, "src": [0, 1, 2]}]


def test_oracle_block_loop_carries_the_consumed_maps():
    depths, normals, conf, K, E = _block_inputs()
    before = {k: v.copy() for k, v in depths.items()}
    r = fuse_np.fuse_block(VIEW_LIST, depths, normals, conf, K, E, min_consistent=2, **TH)
    assert sorted(r) == [0, 1, 3] and r[0]["masks"].shape[0] == 5 and r[0]["sources"] == [1, 2, 3, 1, 1]
    assert not (r[0]["masks"][0] & r[0]["masks"][3] & (before[1] == 0)).any()
    # a second visit of source 1 cannot re-use pixels the first visit consumed
    first, second = r[0]["masks"][0], r[0]["masks"][3]
    assert second.sum() < first.sum()
    for k in depths:
        changed = depths[k] != before[k]
        assert (depths[k][changed] == 0).all()
    assert (depths[0] == 0).sum() > (before[0] == 0).sum()


@pytest.mark.gpu
def test_fuse_block_matches_the_oracle_loop_including_repeated_sources():
    depths, normals, conf, K, E = _block_inputs()
    want_depths = {k: v.copy() for k, v in depths.items()}
    want = fuse_np.fuse_block(VIEW_LIST, want_depths, normals, conf, K, E, min_consistent=2, **TH)
    dev = torch.device("cuda", 0)
    up = lambda d: {k: torch.from_numpy(v).to(dev) for k, v in d.items()}          # noqa: E731
    got_depths = up(depths)
    seen = []
    got = fusion.fuse_block(VIEW_LIST, got_depths, up(normals), up(conf), K, E, min_consistent=2,
                            position_threshold=TH["position_threshold"], depth_threshold=TH["depth_threshold"],
                            normal_threshold=10.0, confidence_threshold=TH["confidence_threshold"],
                            on_view=lambda ref, res: seen.append(ref))
    assert seen == [0, 1, 3] and sorted(got) == sorted(want)
    for ref in want:
        assert got[ref]["sources"] == want[ref]["sources"]
        assert np.array_equal(got[ref]["masks"].cpu().numpy(), want[ref]["masks"]), ref
        assert np.array_equal(got[ref]["count"].cpu().numpy(), want[ref]["count"]), ref
        assert np.array_equal(got[ref]["final_mask"].cpu().numpy(), want[ref]["final_mask"]), ref
        assert _close(got[ref]["xyz"].cpu().numpy(), want[ref]["xyz"], float(np.abs(want[ref]["xyz"]).max())), ref
    for k in want_depths:                                   # the state the next scene block would start from
        assert np.array_equal(got_depths[k].cpu().numpy(), want_depths[k]), k


def test_camera_files_are_parsed_as_the_fusion_stage_parses_them(tmp_path):
    """`<name>.txt` written by the MVS stage (formats.write_red_cam) read back the way fusion_3d_normal.py:112-172 does:
    float32, and the default 'Tcw' orientation runs the extrinsics through two float32 inversions."""
    from deep3d_aerial_b200 import formats
    sc = synth.fusion_scene(num_src=1, height=8, width=8, seed=0)
    _, _, k, e, _ = sc["ref"]
    cam = np.zeros((2, 4, 4), dtype=np.float32)
    cam[0], cam[1, :3, :3], cam[1, 3] = e, k, [400.0, 0.5, 384, 640.0]
    p = str(tmp_path / "v.txt")
    formats.write_red_cam(p, cam, ["8", "8", "3", "v.png"], "/data/images/v.png")
    intr, extr, path = fusion.read_camera_parameters(p, scale=0.5)
    want_k = k.copy()
    want_k[:2] *= 0.5
    assert intr.dtype == np.float32 and np.array_equal(intr, want_k) and path == "/data/images/v.png"
    want_e = np.linalg.inv(np.linalg.inv(e))                       # float32 both times (create_extrinsics_matrix, 'Tcw')
    assert extr.dtype == np.float32 and np.array_equal(extr, want_e)
    up = fusion.read_camera_parameters(p, cams_ori="XrightYup")[1]
    flip = np.diag([1.0, -1.0, -1.0, 1.0]).astype(np.float32)             # Tcw of a y-up camera: rows 1, 2 change sign
    assert np.allclose(up, flip @ e, rtol=1e-4, atol=1e-2)
    twc = fusion.read_camera_parameters(p, images_ori="Twc")[1]
    assert np.allclose(twc, np.linalg.inv(e), atol=1e-3)


@pytest.mark.gpu
def test_fusion_pipeline_matches_direct_calls():
    """Overlapped uploads / kernel / downloads give the bytes the direct call gives, view after view, slot reuse included."""
    dev = torch.device("cuda", 0)
    scenes = [synth.fusion_scene(num_src=3, height=48, width=64, seed=20 + i) for i in range(5)]
    src_d = [torch.from_numpy(v[0]).to(dev) for v in scenes[0]["src"]]
    src_n = [torch.from_numpy(v[1]).to(dev) for v in scenes[0]["src"]]
    th = dict(position_threshold=1.0, depth_threshold=0.01, normal_threshold_cos=TH["normal_cos"],
              confidence_threshold=0.2, min_consistent=2)
    pipe = fusion.FusionPipeline(48, 64, 3, dev, **th)
    got = []
    for i, sc in enumerate(scenes):
        d, n, k, e, prob = sc["ref"]
        geom = fusion.pair_geometry(k, e, [v[2] for v in scenes[0]["src"]], [v[3] for v in scenes[0]["src"]])
        pipe.submit(torch.from_numpy(d).pin_memory(), torch.from_numpy(n).pin_memory(), torch.from_numpy(prob).pin_memory(),
                    geom, src_d, src_n)
        if i >= 1:
            got.append({k_: v.clone() for k_, v in pipe.collect().items()})
    got += [{k_: v.clone() for k_, v in r.items()} for r in pipe.drain()]
    assert len(got) == 5
    for _ in range(2):
        pipe.submit(torch.from_numpy(d).pin_memory(), torch.from_numpy(n).pin_memory(), torch.from_numpy(prob).pin_memory(),
                    geom, src_d, src_n)
    with pytest.raises(RuntimeError, match="collect"):          # both slots hold uncollected views
        pipe.submit(torch.from_numpy(d).pin_memory(), torch.from_numpy(n).pin_memory(), torch.from_numpy(prob).pin_memory(),
                    geom, src_d, src_n)
    pipe.drain()
    for sc, g in zip(scenes, got):
        d, n, k, e, prob = sc["ref"]
        geom = torch.from_numpy(fusion.pair_geometry(k, e, [v[2] for v in scenes[0]["src"]],
                                                     [v[3] for v in scenes[0]["src"]])).to(dev)
        want = fusion.fuse_view(torch.from_numpy(d).to(dev), torch.from_numpy(n).to(dev), torch.from_numpy(prob).to(dev),
                                geom, src_d, src_n, **th)
        for key in fusion.FusionPipeline.KEYS:
            assert torch.equal(g[key], want[key].cpu()), key
