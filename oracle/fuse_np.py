"""ORACLE — TEST INFRASTRUCTURE ONLY (imported by tests/, smoke() and bench.py's cpu_baseline leg; never by the
product path).

CPU restatement in numpy of the depth-map fusion consistency check:

    check()        ConsistencyChecker.check_cupy           fuse/consistency_check_n.py:29-138
    fuse_view()    the per-reference-view accumulation      fuse/fusion_3d_normal.py:436-541

The reference runs this on CuPy, which is not in this image; numpy follows the same promotion rules
(int64 grid x float32 depth -> float64 points; float32 rotation x float32 normals -> float32), with ONE
documented difference restated here: CuPy wraps out-of-bounds integer-array indices around the axis where numpy
raises (CuPy docs, "Differences between CuPy and NumPy": out-of-bounds indices), so every fancy index below goes
through `np.mod`.  Pinning: `tests/golden/fuse_*.npz` hold the outputs of the LIVE reference
(`oracle/make_golden_fuse.py` imports it with `cupy` bound to a numpy shim) on scenes whose projections all land
inside the source maps; the wrap-around cases have no reference run behind them (parity unpinned for those).
"""
from __future__ import annotations

import math

import numpy as np


def normal_threshold_cos(degrees):
    return math.cos(math.radians(degrees))                  # consistency_check_n.py:22


def check(depth_ref, normal_ref, intrinsics_ref, extrinsics_ref, depth_src, normal_src, intrinsics_src,
          extrinsics_src, prob_map_ref, position_threshold=1.0, depth_threshold=0.01, normal_cos=0.0,
          confidence_threshold=0.2):
    """-> mask [H,W] bool, depth_reprojected [H,W] f32, depth_src with consumed pixels zeroed (a copy),
    xyz_world_src [3,H,W] f32, angle_confidence [3,H,W] (cos, three identical planes)."""
    depth_src = np.array(depth_src)                          # cp.array copies: the caller's map is untouched
    height, width = depth_ref.shape
    hs, ws = depth_src.shape
    valid = depth_ref > 0
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    xr, yr = x_ref.reshape([-1]), y_ref.reshape([-1])
    ones = np.ones_like(xr)

    with np.errstate(all="ignore"):
        xyz_ref = np.matmul(np.linalg.inv(intrinsics_ref), np.vstack((xr, yr, ones)) * depth_ref.reshape([-1]))
        xyz_src = np.matmul(np.matmul(extrinsics_src, np.linalg.inv(extrinsics_ref)), np.vstack((xyz_ref, ones)))[:3]
        k_xyz = np.matmul(intrinsics_src, xyz_src)
        xy_src = k_xyz[:2] / k_xyz[2:3]
        x_src = (xy_src[0].reshape([height, width]) + 0.5).astype(np.int64)
        y_src = (xy_src[1].reshape([height, width]) + 0.5).astype(np.int64)
        yi, xi = np.mod(y_src, hs), np.mod(x_src, ws)       # CuPy's wrap-around
        sampled_depth = depth_src[yi, xi]
        sampled_normal = normal_src[yi, xi, :]

        xyz_src = np.matmul(np.linalg.inv(intrinsics_src),
                            np.vstack((x_src.reshape([-1]), y_src.reshape([-1]), ones)) * sampled_depth.reshape([-1]))
        src_world = np.matmul(np.linalg.inv(extrinsics_src), np.vstack((xyz_src, ones)))
        xyz_rep = np.matmul(extrinsics_ref, src_world)[:3]
        depth_rep = xyz_rep[2].reshape([height, width]).astype(np.float32)
        k_rep = np.matmul(intrinsics_ref, xyz_rep)
        xy_rep = k_rep[:2] / k_rep[2:3]
        x_rep = xy_rep[0].reshape([height, width]).astype(np.float32)
        y_rep = xy_rep[1].reshape([height, width]).astype(np.float32)

        dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
        rel = np.abs(depth_rep - depth_ref) / depth_ref

        ns = sampled_normal.transpose([2, 0, 1])
        ns_world = np.matmul(np.linalg.inv(extrinsics_src[:3, :3]), ns.reshape([3, -1]))
        ns_world = ns_world.reshape([3, height, width]).transpose([1, 2, 0])
        nr = normal_ref.transpose([2, 0, 1])
        nr_world = np.matmul(np.linalg.inv(extrinsics_ref[:3, :3]), nr.reshape([3, -1]))
        nr_world = nr_world.reshape([3, height, width]).transpose([1, 2, 0])
        cos = np.sum(nr_world * ns_world, axis=-1)
        cos /= (np.linalg.norm(nr_world, axis=-1) * np.linalg.norm(ns_world, axis=-1))

        mask = np.logical_and(dist < position_threshold, rel < depth_threshold)
        mask = np.logical_and(mask, prob_map_ref > confidence_threshold)
        mask = np.logical_and(mask, cos > normal_cos)
        mask = np.logical_and(mask, valid)

    angle = np.repeat(np.expand_dims(cos, axis=0), 3, axis=0)
    depth_rep[~mask] = 0
    # consistency_check_n.py:123-126: the consumed source pixel is (x_src[mask] + 0.5).astype(int) -- on the INTEGER x_src,
    # i.e. x_src itself where it is >= 0 and x_src + 1 where it is negative (truncation towards zero; negative
    # coordinates only occur through the wrap-around above), then CuPy's index rule again
    x_inv = (x_src[mask] + 0.5).astype(np.int64)
    y_inv = (y_src[mask] + 0.5).astype(np.int64)
    depth_src[np.mod(y_inv, hs), np.mod(x_inv, ws)] = 0
    xyz_world = src_world[0:3].reshape([3, height, width]).astype(np.float32)
    xyz_world[:, ~mask] = 0
    angle[:, ~mask] = 0
    angle[angle < 0] = 0
    return mask, depth_rep, depth_src, xyz_world, angle


def fuse_view(depth_ref, normal_ref, intrinsics_ref, extrinsics_ref, prob_map_ref, sources, min_consistent=4, **th):
    """`sources` = [(depth_src, normal_src, intrinsics_src, extrinsics_src), ...] in fusion order.
    -> dict(count [H,W] i32, xyz [3,H,W] f32, final_mask, depth_ref_filtered, masks [S,H,W], depth_src_out [S,...])."""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    xr, yr = x_ref.reshape([-1]), y_ref.reshape([-1])
    cam = np.matmul(np.linalg.inv(intrinsics_ref), np.vstack((xr, yr, np.ones_like(xr))) * depth_ref.reshape([-1]))
    world = np.matmul(np.linalg.inv(extrinsics_ref), np.vstack((cam, np.ones_like(xr))))[:3]
    all_xyz = world.reshape([-1, height, width]).astype(np.float32)             # fusion_3d_normal.py:449-454
    conf_sum = 0 + np.ones_like(all_xyz)                                        # :455
    count = 0 + np.ones([height, width], dtype=np.int32)                        # :466
    masks, outs = [], []
    for depth_src, normal_src, k_src, e_src in sources:
        mask, _, src_removed, xyz_src, angle = check(depth_ref, normal_ref, intrinsics_ref, extrinsics_ref, depth_src,
                                                     normal_src, k_src, e_src, prob_map_ref, **th)
        count = count + mask.astype(np.int32)                                   # :525
        all_xyz += (angle * xyz_src).astype(np.float32)                         # :526
        conf_sum = conf_sum + angle                                             # :527
        masks.append(mask)
        outs.append(src_removed)
    xyz = (all_xyz / conf_sum).astype(np.float32)                               # :533-534
    final = np.array(count >= min_consistent)                                   # :537
    filtered = np.array(depth_ref)
    filtered[~final] = 0                                                        # :541-543
    return {"count": count, "xyz": xyz, "final_mask": final, "depth_ref_filtered": filtered,
            "masks": np.stack(masks), "depth_src_out": outs}


def fuse_block(view_list, depths, normals, confidences, intrinsics, extrinsics, fusion_num=10, min_consistent=4, **th):
    """The reference-view loop of `fuse_depths` (fusion_3d_normal.py:405-543) with the tmp/*_init.pfm files kept in a
    dict: a source's map is replaced by what `check` returns (:519-523), the reference view's by its final-mask
    pixels (:539-543), and every later read takes the replaced map (:409-411, 475-477).  `depths` is modified.
    No live run of the reference stands behind this loop (that file needs IO.*, tools.*, matplotlib and a parsed
    command line to import): parity unpinned for the loop, pinned for `check`."""
    results = {}
    for pair in view_list:
        ref = pair["ref"]
        if ref not in depths:
            continue
        d_ref, n_ref, prob = depths[ref], normals[ref], confidences[ref]
        k_ref, e_ref = intrinsics[ref], extrinsics[ref]
        height, width = d_ref.shape
        x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
        xr, yr = x_ref.reshape([-1]), y_ref.reshape([-1])
        cam = np.matmul(np.linalg.inv(k_ref), np.vstack((xr, yr, np.ones_like(xr))) * d_ref.reshape([-1]))
        world = np.matmul(np.linalg.inv(e_ref), np.vstack((cam, np.ones_like(xr))))[:3]
        all_xyz = world.reshape([-1, height, width]).astype(np.float32)
        conf_sum = 0 + np.ones_like(all_xyz)
        count = 0 + np.ones([height, width], dtype=np.int32)
        masks, used = [], []
        for src in pair["src"][:fusion_num]:
            if src not in depths:
                continue
            mask, _, removed, xyz_src, angle = check(d_ref, n_ref, k_ref, e_ref, depths[src], normals[src],
                                                     intrinsics[src], extrinsics[src], prob, **th)
            depths[src] = removed
            count = count + mask.astype(np.int32)
            all_xyz += (angle * xyz_src).astype(np.float32)
            conf_sum = conf_sum + angle
            masks.append(mask)
            used.append(src)
        if not used:
            # upstream runs the rest of the body with an empty source loop (:525-543): geo_mask_sum is 1 everywhere, so with
            # min_consistent > 1 the final mask is empty and the view's tmp map is written as all zeros; no points are
            # saved for it (:549-551)
            if min_consistent > 1:
                depths[ref] = np.zeros_like(d_ref)
            continue
        final = np.array(count >= min_consistent)
        filtered = np.array(d_ref)
        filtered[~final] = 0
        depths[ref] = filtered
        results[ref] = {"count": count, "xyz": (all_xyz / conf_sum).astype(np.float32), "final_mask": final,
                        "masks": np.stack(masks), "sources": used}
    return results
