"""deep3d_aerial_b200 -- a B200-native (sm_100a) plane-sweep cost-volume engine for the cascade MVS
networks of gpcv-liujin/Deep3D_Aerial (`mvs/mvs_cas`).

Layers (bottom up):
    libd3dsweep.so   hand-written CUDA behind the C ABI of include/d3d_sweep.h   (csrc/)
    _lib             ctypes binding (fails loudly when the library is missing)
    sweep            tensor-level calls: to_texels, cost_volume, depth_regress, depth_samples
    module           the reference's `models/module.py` hot-path names
    depthnets        the reference's DepthNet / InferDepthNet forward passes
    shard            reference-view sharding across the GPUs of a node (no collective on the path)
    install()        rebinds the names above inside the reference's own modules (drop-in)

Either side of the path:
    graphs           the per-plane regulariser + streaming soft-argmax loops as CUDA graphs
    formats, dataset the workspace text files, PFM / camera files and the MVSDataset tensors (host-side, as upstream)
    predict, mvs_dl  predict.py's command line and mvs_dl.MVS_Inference on this engine
    fusion           ConsistencyChecker / fuse_view: the depth-map fusion consistency check (d3d_consistency_fuse)
    feature_cache    FeatureNet pyramids shared across reference views
"""
from __future__ import annotations

__version__ = "0.1.0"


def install(models_pkg=None):
    """Rebind the hot path of an importable reference checkout to this engine.

    Call once, before the model is built, from a process whose `sys.path` contains the reference's
    `mvs/mvs_cas` directory (as `predict.py` has):

        import deep3d_aerial_b200; deep3d_aerial_b200.install()

    Only the parameter-free hot-path code is replaced; the CNNs (FeatureNet, CostRegNet, GRU
    regularisers) and every `state_dict` key stay the reference's own.  Returns the list of
    rebound names.  Raises if the CUDA library is missing (there is no fallback).
    """
    import importlib

    from . import _lib, depthnets, module

    _lib.load()
    pkg = models_pkg or "models"
    done = []

    def imp(name):
        try:
            return importlib.import_module(pkg + "." + name)
        except ImportError:
            return None

    hot = {n: getattr(module, n) for n in module.__all__}
    for modname in ("module", "cas_mvsnet", "adamvs", "msrednet", "ucsnet"):
        m = imp(modname)
        if m is None:
            continue
        for n, fn in hot.items():
            if hasattr(m, n):
                setattr(m, n, fn)
                done.append("%s.%s" % (modname, n))
    binds = (("cas_mvsnet", "DepthNet", depthnets.cas_depthnet_forward),
             ("msrednet", "DepthNet", depthnets.red_depthnet_forward),
             ("msrednet", "InferDepthNet", depthnets.red_infer_forward),
             ("adamvs", "InferDepthNet", depthnets.ada_infer_forward),
             ("adamvs", "DepthNet", depthnets.ada_depthnet_forward))
    for modname, cls, fwd in binds:
        m = imp(modname)
        if m is not None and hasattr(m, cls):
            getattr(m, cls).forward = fwd
            done.append("%s.%s.forward" % (modname, cls))
    m = imp("ucsnet")
    if m is not None:
        m.compute_depth = depthnets.ucs_compute_depth
        m.uncertainty_aware_samples = depthnets.ucs_uncertainty_samples
        done += ["ucsnet.compute_depth", "ucsnet.uncertainty_aware_samples"]
    return done
