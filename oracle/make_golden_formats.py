"""ORACLE — TEST INFRASTRUCTURE ONLY.

Golden vectors for SURVEY.md §8 row f2 (input contract / on-disk formats), made by running the LIVE reference
(`/root/reference/mvs/mvs_cas/datasets/{data_io,preprocess,cas_normal_eval}.py`, imported, never copied) on a tiny
synthetic workspace.  Run here, in the authoring container:  `python -m oracle.make_golden_formats`.

The reference's `datasets` package imports `gdal`, `matplotlib` and `imageio` at module scope (none of them is in
this image and none is used by the functions pinned here) and calls `np.float`, removed in numpy 1.24: the three
modules are stubbed in `sys.modules` and `np.float = float` is restored for the duration of the run.

Writes
    tests/golden/workspace_tiny/{cameras,images,viewpair}.txt, image_path.template.txt, img/*.png   (the fixture)
    tests/golden/formats_dataset.npz   MVSDataset items of the live reference for two argument sets
    tests/golden/formats_files.npz     bytes of the PFM / camera files the live reference writes, and what it reads back
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
WS = os.path.join(OUT, "workspace_tiny")
REF_ROOT = "/root/reference/mvs/mvs_cas"

ARGSETS = {
    "plain": dict(min_interval=0.1, interval_scale=1.0, numdepth=48, resize_scale=1.0, sample_scale=1.0,
                  max_h=64, max_w=96),
    "scaled": dict(min_interval=0.1, interval_scale=1.0, numdepth=32, resize_scale=0.5, sample_scale=0.25,
                   max_h=96, max_w=128),
}


def load_live_datasets():
    for name in ("gdal", "matplotlib", "matplotlib.pyplot", "imageio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for fn in ("imread", "imsave", "imwrite"):
        setattr(sys.modules["imageio"], fn, None)
    if not hasattr(np, "float"):
        np.float = float        # cas_normal_eval.py:69
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    return types.SimpleNamespace(data_io=importlib.import_module("datasets.data_io"),
                                 preprocess=importlib.import_module("datasets.preprocess"),
                                 eval=importlib.import_module("datasets.cas_normal_eval"))


def rot(axis, deg):
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    m = {"x": [[1, 0, 0], [0, c, -s], [0, s, c]], "y": [[c, 0, s], [0, 1, 0], [-s, 0, c]],
         "z": [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[axis]
    return np.array(m)


def make_workspace():
    """4 images of 96 x 128 (PNG, lossless), two camera models, oblique poses, a view list with a short entry
    (padded by the reader) and an entry without sources (dropped by the reader)."""
    from PIL import Image
    os.makedirs(os.path.join(WS, "img"), exist_ok=True)
    rng = np.random.default_rng(7)
    names = ["a_000.png", "a_001.png", "b_002.png", "b_003.png"]
    for i, n in enumerate(names):
        yy, xx = np.mgrid[0:96, 0:128]
        base = 96 + 60 * np.sin(xx / (7.0 + i)) * np.cos(yy / (5.0 + i))
        img = np.clip(base[..., None] + rng.normal(0, 25, (96, 128, 3)), 0, 255).astype(np.uint8)
        Image.fromarray(img).save(os.path.join(WS, "img", n))
    with open(os.path.join(WS, "cameras.txt"), "w") as f:
        f.write("# camera_id width height pixelsize fx fy x0 y0 k1 k2 k3 p1 p2\n")
        f.write("1 128 96 0.0046 410.25 409.75 63.5 47.25 0 0 0 0 0\n\n")
        f.write("2 128 96 0.0046 398.5 398.5 64.125 48.5 1e-3 0 0 0 0\n")
    with open(os.path.join(WS, "images.txt"), "w") as f:
        f.write("# image_id camera_id R(9) C(3) dmin dmax name\n")
        for i, n in enumerate(names):
            r = rot("z", 3.0 * i) @ rot("x", 180 + 2.0 * i) @ rot("y", -4.0 + 2.5 * i)
            c = np.array([12.5 * i, -3.25 * i, 520.0 + 1.5 * i])
            vals = list(r.reshape(-1)) + list(c) + [400.0 + i, 640.0 - 2 * i]
            f.write("%d %d %s %s\n" % (i, 1 + i // 2, " ".join(repr(float(v)) for v in vals), n))
    with open(os.path.join(WS, "image_path.template.txt"), "w") as f:
        f.write("4\n")
        for i, n in enumerate(names):
            f.write("%d %s {ROOT}/img/%s\n" % (i, n, n))
    with open(os.path.join(WS, "viewpair.txt"), "w") as f:
        f.write("4\n0\n3 1 0.9 2 0.8 3 0.7\n1\n2 0 0.9 2 0.5\n2\n0\n3\n1 2 0.4\n")


def materialise(dst):
    """Copy the fixture to `dst` with absolute image paths filled in (image_path.txt holds absolute paths)."""
    import shutil
    shutil.copytree(WS, dst, dirs_exist_ok=True)
    with open(os.path.join(WS, "image_path.template.txt")) as f:
        text = f.read().replace("{ROOT}", dst)
    with open(os.path.join(dst, "image_path.txt"), "w") as f:
        f.write(text)
    return dst


def flatten_item(prefix, item, out):
    for k, v in item.items():
        if isinstance(v, dict):
            flatten_item(prefix + k + ".", v, out)
        elif isinstance(v, np.ndarray):
            out[prefix + k] = v
        elif isinstance(v, list):
            out[prefix + k] = np.array([str(x) for x in v])
        # ref_image_path is an absolute path of the run: not stored


def main():
    make_workspace()
    ref = load_live_datasets()
    with tempfile.TemporaryDirectory() as tmp:
        ws = materialise(os.path.join(tmp, "ws"))
        out = {}
        for tag, kw in ARGSETS.items():
            args = types.SimpleNamespace(**kw)
            for view_num, norm in ((3, "mean"), (2, "standard")):
                ds = ref.eval.MVSDataset(ws, "val", view_num, norm, args)
                out["%s.v%d.len" % (tag, view_num)] = np.array(len(ds))
                out["%s.v%d.samples" % (tag, view_num)] = np.array(json.dumps(ds.sample_list))   # ragged
                for idx in range(len(ds)):
                    flatten_item("%s.v%d.%d." % (tag, view_num, idx), ds[idx], out)
        np.savez_compressed(os.path.join(OUT, "formats_dataset.npz"), **out)

        files = {}
        rng = np.random.default_rng(3)
        depth = rng.uniform(400, 640, (5, 7)).astype(np.float32)
        colour = rng.normal(0, 1, (4, 6, 3)).astype(np.float32)
        for tag, arr in (("grey", depth), ("colour", colour), ("grey1", depth[:, :, None])):
            p = os.path.join(tmp, tag + ".pfm")
            ref.data_io.save_pfm_utf8(p, arr)
            files["pfm." + tag + ".in"] = arr
            files["pfm." + tag + ".bytes"] = np.fromfile(p, dtype=np.uint8)
            if tag != "grey1":
                back, scale = ref.data_io.load_pfm_utf8(p)
                files["pfm." + tag + ".back"] = np.ascontiguousarray(back)
                files["pfm." + tag + ".scale"] = np.array(scale)
        item = ref.eval.MVSDataset(ws, "val", 3, "mean", types.SimpleNamespace(**ARGSETS["plain"]))[0]
        p = os.path.join(tmp, "cam.txt")
        ref.data_io.write_red_cam(p, item["outcam"], item["outlocation"], "/data/images/a_000.png")
        files["cam.in"] = item["outcam"]
        files["cam.location"] = np.array(item["outlocation"])
        files["cam.bytes"] = np.fromfile(p, dtype=np.uint8)
        for mode in ("standard", "mean", "vit"):
            files["center." + mode] = ref.preprocess.center_image(item["outimage"], mode=mode)
        files["center.in"] = item["outimage"]
        np.savez_compressed(os.path.join(OUT, "formats_files.npz"), **files)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
