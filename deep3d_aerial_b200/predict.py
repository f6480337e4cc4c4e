"""`python -m deep3d_aerial_b200.predict` -- the reference's `mvs/mvs_cas/predict.py` command line (SURVEY.md §8
level b1) on this engine: same options, same input files, same output files, no gdal / matplotlib.

    python -m deep3d_aerial_b200.predict --reference_root <checkout>/mvs/mvs_cas \
        --data_folder <ws>/export --output_folder <ws>/dense/MVS --model adamvs --loadckpt model.ckpt \
        --view_num 5 --numdepth 384 --max_w 2752 --max_h 1856

What is whose: the networks (`Infer_CascadeMVSNet`, `Infer_AdaMVSNet`, `Infer_CascadeREDNet`; FeatureNet and the
regularisers with their `state_dict` keys) are imported from the reference checkout named by `--reference_root`
and stay PyTorch; `deep3d_aerial_b200.install()` rebinds their hot path to libd3dsweep; the dataset tensors and
the PFM / camera writers are this package's (`dataset.py`, `formats.py`).  Under torchrun every rank takes its
share of the reference views (`shard.partition`), one process per GPU, nothing exchanged.
Differences from upstream, all deliberate: no `nn.DataParallel` wrapper (checkpoint keys lose their `module.`
prefix on load), no DataLoader worker processes, `--loadckpt ''` runs with random weights (smoke tests).
"""
from __future__ import annotations

import argparse
import os
import sys
import time


def build_parser():
    p = argparse.ArgumentParser(description="Predict depth")      # predict.py:32-62, same names and defaults
    p.add_argument("--model", default="adamvs", help="casmvsnet, ucsnet, msrednet or adamvs")
    p.add_argument("--dataset", default="cas_normal_eval")
    p.add_argument("--data_folder", required=True)
    p.add_argument("--output_folder", required=True)
    p.add_argument("--loadckpt", default="")
    p.add_argument("--view_num", type=int, default=5)
    p.add_argument("--numdepth", type=int, default=192)
    p.add_argument("--max_w", type=int, default=3584)
    p.add_argument("--max_h", type=int, default=4096)
    p.add_argument("--min_interval", type=float, default=0.1)
    p.add_argument("--fext", type=str, default=".jpg")
    p.add_argument("--normalize", type=str, default="mean")
    p.add_argument("--resize_scale", type=float, default=1.0)
    p.add_argument("--sample_scale", type=float, default=1)
    p.add_argument("--interval_scale", type=float, default=1)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--display", default=True)
    p.add_argument("--share_cr", action="store_true")
    p.add_argument("--ndepths", type=str, default="48,32,8")
    p.add_argument("--depth_inter_r", type=str, default="4,2,1")
    p.add_argument("--cr_base_chs", type=str, default="8,8,8")
    p.add_argument("--reference_root", default=os.environ.get("DEEP3D_REFERENCE_ROOT", ""),
                   help="the reference checkout's mvs/mvs_cas directory (its `models` package holds the networks)")
    p.add_argument("--plane_loop_graphs", action="store_true", help="CUDA-graph the per-plane loops (row f1)")
    p.add_argument("--feature_cache", type=int, default=0,
                   help="keep the FeatureNet pyramids of this many images resident and reuse them across reference "
                        "views (row f4); 0 = recompute per view, as upstream")
    p.add_argument("--partition", default="round_robin", choices=["round_robin", "contiguous"],
                   help="how reference views are dealt to ranks; contiguous keeps neighbours (shared sources) together")
    return p


def build_model(args):
    """predict.py:75-101 -- the reference's own classes, hot path rebound by install()."""
    if args.reference_root and args.reference_root not in sys.path:
        sys.path.insert(0, args.reference_root)
    import deep3d_aerial_b200 as d3d
    d3d.install()
    ints = lambda s: [int(x) for x in s.split(",") if x]                      # noqa: E731
    kw = dict(num_depth=args.numdepth, ndepths=ints(args.ndepths),
              depth_intervals_ratio=[float(x) for x in args.depth_inter_r.split(",") if x],
              share_cr=args.share_cr, cr_base_chs=ints(args.cr_base_chs))
    if args.model == "casmvsnet":
        from models.cas_mvsnet import Infer_CascadeMVSNet as Net
    elif args.model == "msrednet":
        from models.msrednet import Infer_CascadeREDNet as Net
    elif args.model == "adamvs":
        from models.adamvs import Infer_AdaMVSNet as Net
    elif args.model == "ucsnet":
        from models.ucsnet import Infer_UCSNet as Net
        kw = dict(lamb=1.5, num_depth=args.numdepth, ndepths=ints(args.ndepths))
    else:
        raise Exception("{}? Not implemented yet!".format(args.model))
    return Net(**kw)


def load_checkpoint(model, path):
    import torch
    state = torch.load(path, map_location="cpu")["model"]                    # predict.py:108-109
    model.load_state_dict({(k[7:] if k.startswith("module.") else k): v for k, v in state.items()})


def predict_depth(args, model=None):
    import numpy as np
    import torch

    import contextlib

    from . import dataset, depthnets, shard
    from .feature_cache import FeatureCache

    if str(args.display).lower() in ("false", "0", "no"):
        args.display = False
    rank, world, local_rank = shard.world()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if model is None:
        model = build_model(args)
        if args.loadckpt:
            print("loading model {}".format(args.loadckpt))
            load_checkpoint(model, args.loadckpt)
    if args.plane_loop_graphs:
        depthnets.PLANE_LOOP_GRAPHS = True
    model = model.to(device).eval()
    cache = None
    if args.feature_cache > 0 and hasattr(model, "feature"):
        cache = FeatureCache(args.feature_cache).attach(model.feature)
        # the cached maps of an image are the same tensors for every reference view: lay each out once (texel_pool.py)
        from .texel_pool import TexelPool
        depthnets.TEXEL_POOL = TexelPool(max(args.feature_cache, 2 * args.view_num))
    views = dataset.MVSDataset(args.data_folder, "val", args.view_num, args.normalize, args)
    os.makedirs(args.output_folder, exist_ok=True)
    written = []
    t_first = time.time()
    with torch.no_grad():
        for idx in shard.partition(range(len(views)), world, rank, mode=args.partition):
            t0 = time.time()
            sample = dataset.collate(views[idx])
            imgs = sample["imgs"].to(device, non_blocking=True)
            proj = {k: v.to(device) for k, v in sample["proj_matrices"].items()}
            ids = views.sample_list[idx][:args.view_num]
            with (cache.views(ids) if cache else contextlib.nullcontext()):
                outputs = model(imgs, proj, sample["depth_values"].to(device))
            depth = outputs["depth"].float().cpu().numpy()
            prob = outputs["photometric_confidence"].float().cpu().numpy()
            t1 = time.time()
            location = [x[0] for x in sample["outlocation"]]
            print(np.array(location))
            paths = dataset.save_view_outputs(args.output_folder, depth, prob, sample["outcam"][0].numpy(), location,
                                              sample["ref_image_path"][0], display=bool(args.display))
            written.append(paths)
            print("depth inference {} finished, image {} finished, ({:3f}s and {:3f} sec/step)".format(
                len(written), os.path.splitext(location[3])[0], t1 - t0, time.time() - t1))
    print("final, total_cnt = {}, total_time = {:3f}".format(len(written), time.time() - t_first))
    if cache:
        print("feature cache:", cache.stats())
        print("texel pool:", depthnets.TEXEL_POOL.stats())
        depthnets.TEXEL_POOL = None
        cache.detach()
    return written


def main(argv=None):
    args = build_parser().parse_args(argv)
    print("argv:", sys.argv[1:] if argv is None else argv)
    predict_depth(args)


if __name__ == "__main__":
    main()
