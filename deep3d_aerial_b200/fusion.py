"""Depth-map fusion consistency check on the B200: the reference's
`fuse/consistency_check_n.py` call surface over `d3d_consistency_fuse` (include/d3d_sweep.h).

    ConsistencyChecker(position_threshold, depth_threshold, normal_threshold, confidence_threshold, implement)
        .check(depth_ref, normal_ref, intrinsics_ref, extrinsics_ref, depth_src, normal_src, intrinsics_src,
               extrinsics_src, prob_map_ref) -> (mask, depth_reprojected, depth_src, xyz_world_src, angle_confidence)
                                                 numpy arrays, as upstream (consistency_check_n.py:17-148)
    fuse_view(...)   one launch for a reference view and ALL its source views, device tensors in and out: what
                     the loop of Fuse_Depth_Map.fuse_depths accumulates (fuse/fusion_3d_normal.py:436-541)

The small matrices (inverses and products of K and Tcw) are formed on the host with the reference's own numpy
calls in the matrices' own dtype and handed to the kernel as fp64 (`pair_geometry`); everything per pixel runs in
one CUDA kernel.  No CPU fallback: a missing library or CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

MAX_SRC = _lib.FUSE_MAX_SRC


def pair_geometry(intrinsics_ref, extrinsics_ref, intrinsics_src: Sequence, extrinsics_src: Sequence) -> np.ndarray:
    """[(1+S), 64] float64: the matrices `d3d_consistency_fuse` needs (layout in include/d3d_sweep.h), each
    computed as the reference computes it (consistency_check_n.py:52, 56, 76, 80, 103, 107;
    fusion_3d_normal.py:449-451)."""
    k_ref, e_ref = np.asarray(intrinsics_ref), np.asarray(extrinsics_ref)
    g = np.zeros((1 + len(intrinsics_src), 64), dtype=np.float64)

    def put(block, at, m):                                   # row-major, rows padded to 4 entries
        m = np.asarray(m)
        view = g[block, at:at + 4 * m.shape[0]].reshape(m.shape[0], 4)
        view[:, :m.shape[1]] = m

    e_ref_inv = np.linalg.inv(e_ref)
    put(0, 0, np.linalg.inv(k_ref))
    put(0, 12, e_ref[:3, :4])
    put(0, 24, k_ref)
    put(0, 36, e_ref_inv)
    put(0, 52, np.linalg.inv(e_ref[:3, :3]))
    for s, (k, e) in enumerate(zip(intrinsics_src, extrinsics_src)):
        k, e = np.asarray(k), np.asarray(e)
        put(1 + s, 0, np.matmul(e, e_ref_inv)[:3, :4])
        put(1 + s, 12, k)
        put(1 + s, 24, np.linalg.inv(k))
        put(1 + s, 36, np.linalg.inv(e))
        put(1 + s, 52, np.linalg.inv(e[:3, :3]))
    return g


def _dev(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libd3dsweep has no CPU fallback)" % name)
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError("%s must be contiguous %s, got %s" % (name, dtype, t.dtype))
    return t


def fuse_view(depth_ref, normal_ref, prob_ref, geometry, depth_src, normal_src, *, position_threshold=1.0,
              depth_threshold=0.01, normal_threshold_cos=0.0, confidence_threshold=0.2, min_consistent=4,
              per_source=False, update_sources=True, accumulate=False, out: Optional[dict] = None) -> dict:
    """One reference view against S <= 16 source views, one kernel launch on the current stream.

    depth_ref [H,W], normal_ref [H,W,3], prob_ref [H,W]; depth_src / normal_src: lists of S tensors [Hs,Ws] /
    [Hs,Ws,3]; geometry: `pair_geometry(...)` as a CUDA float64 tensor.  Returns device tensors:
      count [H,W] int32, xyz [3,H,W], final_mask [H,W] bool, depth_ref_filtered [H,W], masks [S,H,W] bool,
      depth_src_out (list of S maps with the consumed pixels zeroed; `update_sources=False` skips them) and, with
      `per_source`, depth_reprojected [S,H,W], xyz_world_src [S,3,H,W], angle_conf [S,H,W].
    `out` may carry preallocated tensors under the same keys.  `accumulate=True` continues the accumulators
    (`out["accum"]` [4,H,W] and `out["count"]`) an earlier call over OTHER sources of this reference view left."""
    lib = _lib.load()
    s_n = len(depth_src)
    if not 1 <= s_n <= MAX_SRC or len(normal_src) != s_n:
        raise ValueError("need 1..%d source views with a normal map each, got %d / %d" % (MAX_SRC, s_n, len(normal_src)))
    _dev(depth_ref, "depth_ref"), _dev(normal_ref, "normal_ref"), _dev(prob_ref, "prob_ref")
    _dev(geometry, "geometry", torch.float64)
    h, w = depth_ref.shape
    hs, ws = depth_src[0].shape
    if tuple(normal_ref.shape) != (h, w, 3) or tuple(prob_ref.shape) != (h, w):
        raise ValueError("normal_ref / prob_ref do not match depth_ref %dx%d" % (h, w))
    if geometry.numel() != (1 + s_n) * _lib.FUSE_GEOM_DOUBLES:
        raise ValueError("geometry holds %d doubles, expected %d" % (geometry.numel(), (1 + s_n) * _lib.FUSE_GEOM_DOUBLES))
    dev = depth_ref.device
    out = dict(out or {})

    def buf(key, shape, dtype):
        t = out.get(key)
        if t is None:
            t = out[key] = torch.empty(shape, device=dev, dtype=dtype)
        if tuple(t.shape) != tuple(shape):
            raise ValueError("out[%r] has shape %s, this call needs %s" % (key, tuple(t.shape), tuple(shape)))
        return _dev(t, key, dtype)

    a = _lib.FuseArgs()
    a.struct_size = C.sizeof(a)
    a.num_src, a.height, a.width, a.src_height, a.src_width = s_n, h, w, hs, ws
    a.min_consistent = int(min_consistent)
    a.position_threshold = float(position_threshold)
    a.depth_threshold, a.confidence_threshold = float(depth_threshold), float(confidence_threshold)
    a.normal_threshold_cos = float(normal_threshold_cos)
    a.depth_ref, a.normal_ref, a.prob_ref = depth_ref.data_ptr(), normal_ref.data_ptr(), prob_ref.data_ptr()
    a.geometry = geometry.data_ptr()
    if update_sources and "depth_src_out" not in out:
        out["depth_src_out"] = [torch.empty_like(d) for d in depth_src]
    for s in range(s_n):
        if tuple(depth_src[s].shape) != (hs, ws) or tuple(normal_src[s].shape) != (hs, ws, 3):
            raise ValueError("source %d: maps must all be %dx%d" % (s, hs, ws))
        a.depth_src[s] = _dev(depth_src[s], "depth_src[%d]" % s).data_ptr()
        a.normal_src[s] = _dev(normal_src[s], "normal_src[%d]" % s).data_ptr()
        if update_sources:
            if len(out["depth_src_out"]) != s_n or tuple(out["depth_src_out"][s].shape) != (hs, ws):
                raise ValueError("out['depth_src_out'] must hold %d maps of %dx%d" % (s_n, hs, ws))
            a.depth_src_out[s] = _dev(out["depth_src_out"][s], "depth_src_out[%d]" % s).data_ptr()
    a.mask = buf("masks", (s_n, h, w), torch.bool).data_ptr()
    a.consistent_count = buf("count", (h, w), torch.int32).data_ptr()
    a.xyz_fused = buf("xyz", (3, h, w), torch.float32).data_ptr()
    a.final_mask = buf("final_mask", (h, w), torch.bool).data_ptr()
    a.depth_ref_filtered = buf("depth_ref_filtered", (h, w), torch.float32).data_ptr()
    if accumulate or "accum" in out:
        if accumulate and ("accum" not in out or "count" not in (out or {})):
            raise ValueError("accumulate=True needs out['accum'] and out['count'] from the earlier call")
        a.accum = buf("accum", (4, h, w), torch.float32).data_ptr()
        a.accumulate = int(bool(accumulate))
    if per_source:
        a.depth_reprojected = buf("depth_reprojected", (s_n, h, w), torch.float32).data_ptr()
        a.xyz_world_src = buf("xyz_world_src", (s_n, 3, h, w), torch.float32).data_ptr()
        a.angle_conf = buf("angle_conf", (s_n, h, w), torch.float32).data_ptr()
    _lib.check(lib.d3d_consistency_fuse(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def read_camera_parameters(filename, scale=1.0, cams_ori="XrightYdown", images_ori="Tcw"):
    """`<name>.txt` as `predict.py` writes it -> (K [3,3] float32, Tcw [4,4] float32, image path), parsed as
    `Fuse_Depth_Map.read_camera_parameters` / `create_extrinsics_matrix` parse it (fusion_3d_normal.py:112-172):
    float32 throughout, rows 0-1 of K times `scale`, and for the default 'Tcw' orientation the extrinsics go
    through inverse, axis flip, inverse -- in float32, which changes their low bits, so it is restated as is."""
    with open(filename) as f:
        lines = [line.rstrip() for line in f.readlines()]
    extr = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape((4, 4))
    flip = np.eye(3, dtype=np.float32)
    if cams_ori == "XrightYup":
        flip[1, 1] = flip[2, 2] = -1
    if images_ori == "Twc":
        extr[0:3, 0:3] = np.matmul(extr[0:3, 0:3], flip)
        extr = np.linalg.inv(extr)
    elif images_ori == "Rcw":
        extr[0:3, 0:3] = np.linalg.inv(np.matmul(flip, extr[0:3, 0:3]))
        extr = np.linalg.inv(extr)
    elif images_ori == "Tcw":
        extr = np.linalg.inv(extr)
        extr[0:3, 0:3] = np.matmul(extr[0:3, 0:3], flip)
        extr = np.linalg.inv(extr)
    intr = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape((3, 3))
    intr[:2, :] *= scale
    image_path = lines[13].split(" ")[4]
    return intr, extr, image_path


def load_mvs_outputs(depth_path, names, device, camera_scale=1.0):
    """What `fuse_depths` reads per view from the MVS output folder (fusion_3d_normal.py:405-440, 470-489), as the
    dicts `fuse_block` takes: `<name>_init.pfm`, `<name>_prob.pfm` (ones if missing), `<name>_normal.pfm` mapped from
    [0,1] to [-1,1] (`read_normal`; missing: the default normal (0,0,-1)), `<name>.txt`.  Views without a depth map
    are left out (upstream warns and skips them)."""
    import os

    from .formats import load_pfm_utf8

    depths, normals, confs, intr, extr = {}, {}, {}, {}, {}
    for name in names:
        base = os.path.join(depth_path, str(name))
        if not os.path.exists(base + "_init.pfm"):
            continue
        depth = np.ascontiguousarray(load_pfm_utf8(base + "_init.pfm")[0], dtype=np.float32)
        h, w = depth.shape
        if os.path.exists(base + "_normal.pfm"):
            normal = np.ascontiguousarray(load_pfm_utf8(base + "_normal.pfm")[0] * 2.0 - 1.0, dtype=np.float32)
        else:
            normal = np.zeros([h, w, 3], dtype=np.float32)
            normal[:, :, 2] = -1.0
        if os.path.exists(base + "_prob.pfm"):
            conf = np.ascontiguousarray(load_pfm_utf8(base + "_prob.pfm")[0], dtype=np.float32)
        else:
            conf = np.ones([h, w], dtype=np.float32)
        intr[name], extr[name], _ = read_camera_parameters(base + ".txt", camera_scale)
        depths[name] = torch.from_numpy(depth).to(device)
        normals[name] = torch.from_numpy(normal).to(device)
        confs[name] = torch.from_numpy(conf).to(device)
    return depths, normals, confs, intr, extr


def fuse_block(view_list, depths, normals, confidences, intrinsics, extrinsics, *, fusion_num=10,
               position_threshold=1.0, depth_threshold=0.01, normal_threshold=10.0, confidence_threshold=0.2,
               min_consistent=4, on_view=None):
    """The reference-view loop of `Fuse_Depth_Map.fuse_depths` (fuse/fusion_3d_normal.py:405-543) with every map
    resident in HBM instead of travelling through `tmp/*_init.pfm` between reference views.

    view_list: [{"ref": id, "src": [ids...]}, ...] in fusion order (blocks.txt / viewpair.txt);
    depths, normals, confidences: dict id -> CUDA tensor [H,W] / [H,W,3] / [H,W]; intrinsics, extrinsics: dict id ->
    numpy K [3,3] / Tcw [4,4].  `depths` IS MODIFIED as upstream modifies its tmp files: after a reference view, each of
    its sources keeps only the pixels that view did not consume (:519-523) and the reference view itself only its
    final-mask pixels (:539-543) -- so later views see earlier views' deletions.  A source named twice in one list
    (the reader pads short lists with the first source, :243-246) is visited again against the map its first visit
    left, in a follow-up launch that continues the accumulators.
    Returns {ref id: dict(count, xyz, final_mask, masks [S,H,W], sources [ids])}; `on_view(ref, result)` is called
    per view if given (results are fresh tensors)."""
    cos = math.cos(math.radians(normal_threshold))
    results = {}
    for pair in view_list:
        ref = pair["ref"]
        if ref not in depths:
            continue                                         # upstream warns and skips (:413-415)
        srcs = [s for s in pair["src"][:fusion_num] if s in depths]
        dev = depths[ref].device
        out, masks, first = {}, [], True
        at = 0
        while at < len(srcs) or first:
            run = []
            while at < len(srcs) and srcs[at] not in run and len(run) < MAX_SRC:
                run.append(srcs[at])
                at += 1
            if not run:                                      # no source at all: the reference pixel alone
                break
            geom = torch.from_numpy(pair_geometry(intrinsics[ref], extrinsics[ref], [intrinsics[s] for s in run],
                                                  [extrinsics[s] for s in run])).to(dev)
            keep = {k: out[k] for k in ("accum", "count") if k in out}
            if first:
                keep["accum"] = torch.empty((4,) + tuple(depths[ref].shape), device=dev, dtype=torch.float32)
            out = fuse_view(depths[ref], normals[ref], confidences[ref], geom, [depths[s] for s in run],
                            [normals[s] for s in run], position_threshold=position_threshold,
                            depth_threshold=depth_threshold, normal_threshold_cos=cos,
                            confidence_threshold=confidence_threshold, min_consistent=min_consistent,
                            accumulate=not first, out=keep)
            for s, new in zip(run, out["depth_src_out"]):
                depths[s] = new                              # tmp/<src>_init.pfm
            masks.append(out["masks"])
            first = False
        if first:
            # no usable source: upstream still fuses the view (fusion_3d_normal.py:436-541 with an empty loop) -- every
            # pixel counts itself only, so with min_consistent > 1 nothing survives and tmp/<ref>_init.pfm is all zeros:
            # views fused later see an emptied map, not the original one
            if min_consistent > 1:
                depths[ref] = torch.zeros_like(depths[ref])
            continue
        depths[ref] = out["depth_ref_filtered"]              # tmp/<ref>_init.pfm
        res = {"count": out["count"], "xyz": out["xyz"], "final_mask": out["final_mask"],
               "masks": torch.cat(masks, 0), "sources": srcs}
        results[ref] = res
        if on_view is not None:
            on_view(ref, res)
    return results


class FusionPipeline:
    """Reference views streamed from pinned host memory against source maps that stay in HBM: the upload of view
    n+1 and the download of view n-1 overlap the kernel of view n (three streams, two slots; PCIe carries both
    directions at once).  `submit` takes pinned CPU tensors depth [H,W], normal [H,W,3], prob [H,W] and the
    `pair_geometry` block (numpy or pinned tensor); `collect` returns the oldest submitted view's
    {"count", "xyz", "final_mask", "depth_ref_filtered"} as pinned CPU tensors (valid until two more submits)."""

    KEYS = ("count", "xyz", "final_mask", "depth_ref_filtered")

    def __init__(self, height, width, num_src, device, slots=2, **thresholds):
        self.dev = torch.device(device)
        self.th = thresholds
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(3)]            # up, compute, down
        self.slots = []
        shapes = {"depth": (height, width), "normal": (height, width, 3), "prob": (height, width)}
        for _ in range(slots):
            s = {k: torch.empty(v, device=self.dev) for k, v in shapes.items()}
            s["geom"] = torch.empty((1 + num_src, _lib.FUSE_GEOM_DOUBLES), device=self.dev, dtype=torch.float64)
            s["geom_host"] = torch.empty((1 + num_src, _lib.FUSE_GEOM_DOUBLES), dtype=torch.float64).pin_memory()
            s["out"] = {}
            s["host"] = None
            for name in ("copied", "done", "downloaded"):
                s[name] = torch.cuda.Event()
                s[name].record(torch.cuda.current_stream(self.dev))
            self.slots.append(s)
        self.pending = []
        self._next = 0

    def submit(self, depth_ref, normal_ref, prob_ref, geometry, depth_src, normal_src):
        s = self.slots[self._next]
        if any(s is q for q in self.pending):
            raise RuntimeError("FusionPipeline: collect() the view submitted %d calls ago before submitting another "
                               "(its slot would be overwritten)" % len(self.slots))
        self._next = (self._next + 1) % len(self.slots)
        up, compute, down = self.streams
        # the caller's device tensors (source maps, or reference maps handed over on the device) were produced on ITS
        # stream: neither the upload nor the kernel may run ahead of that
        producer = torch.cuda.current_stream(self.dev)
        up.wait_stream(producer)
        compute.wait_stream(producer)
        if isinstance(geometry, np.ndarray):
            s["copied"].synchronize()                         # the slot's previous upload has left its pinned block
            s["geom_host"].copy_(torch.from_numpy(geometry))
            geometry = s["geom_host"]
        with torch.cuda.stream(up):
            up.wait_event(s["done"])                          # the slot's previous kernel has read its inputs
            s["depth"].copy_(depth_ref, non_blocking=True)
            s["normal"].copy_(normal_ref, non_blocking=True)
            s["prob"].copy_(prob_ref, non_blocking=True)
            s["geom"].copy_(geometry, non_blocking=True)
            s["copied"].record()
        with torch.cuda.stream(compute):
            compute.wait_event(s["copied"])
            compute.wait_event(s["downloaded"])               # the slot's previous results have left
            # (the pipeline treats the source maps as read-only: the per-source "consumed pixels" copies of fuse_view
            # would be S device-to-device copies per view that nobody reads here -- fuse_block is the stateful loop)
            s["out"] = fuse_view(s["depth"], s["normal"], s["prob"], s["geom"], depth_src, normal_src, out=s["out"],
                                 update_sources=False, **self.th)
            s["done"].record()
        with torch.cuda.stream(down):
            down.wait_event(s["done"])
            if s["host"] is None:
                s["host"] = {k: torch.empty(s["out"][k].shape, dtype=s["out"][k].dtype).pin_memory() for k in self.KEYS}
            for k in self.KEYS:
                s["host"][k].copy_(s["out"][k], non_blocking=True)
            s["downloaded"].record()
        self.pending.append(s)

    def collect(self):
        s = self.pending.pop(0)
        s["downloaded"].synchronize()
        return s["host"]

    def drain(self):
        return [self.collect() for _ in range(len(self.pending))]


class ConsistencyChecker(object):
    """Drop-in for `fuse/consistency_check_n.py:ConsistencyChecker`: same constructor, same `check` signature,
    numpy arrays in and out.  `implement` is accepted for compatibility (upstream asserts it is one of
    numpy / cupy / torch and always runs CuPy); here every call runs the CUDA kernel on the current device."""

    def __init__(self, position_threshold, depth_threshold, normal_threshold, confidence_threshold, implement="cupy"):
        self.position_threshold = position_threshold
        self.depth_threshold = depth_threshold
        self.normal_threshold = math.cos(math.radians(normal_threshold))
        self.confidence_threshold = confidence_threshold
        print("normal_th:" + str(self.normal_threshold))
        self.implement = implement.lower()
        assert self.implement in ["numpy", "cupy", "torch"]
        _lib.load()

    def check(self, depth_ref, normal_ref, intrinsics_ref, extrinsics_ref, depth_src, normal_src, intrinsics_src,
              extrinsics_src, prob_map_ref):
        dev = torch.device("cuda", torch.cuda.current_device())

        def up(x):
            return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)

        geom = torch.from_numpy(pair_geometry(intrinsics_ref, extrinsics_ref, [intrinsics_src], [extrinsics_src])).to(dev)
        r = fuse_view(up(depth_ref), up(normal_ref), up(prob_map_ref), geom, [up(depth_src)], [up(normal_src)],
                      position_threshold=self.position_threshold, depth_threshold=self.depth_threshold,
                      normal_threshold_cos=self.normal_threshold, confidence_threshold=self.confidence_threshold,
                      min_consistent=1, per_source=True)
        angle = r["angle_conf"][0].cpu().numpy()
        return (r["masks"][0].cpu().numpy(), r["depth_reprojected"][0].cpu().numpy(), r["depth_src_out"][0].cpu().numpy(),
                r["xyz_world_src"][0].cpu().numpy(), np.repeat(angle[None], 3, axis=0))
