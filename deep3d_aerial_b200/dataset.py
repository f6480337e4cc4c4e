"""Input contract of the cascade MVS networks: workspace text files + images -> the tensors
`model(imgs, proj_matrices, depth_values)` takes, and the files `predict.py` leaves behind.

Host-side Python/numpy as upstream; same names, arguments and arithmetic (dtype for dtype, so the float32
projection pyramids come out bit-identical) as

    scale_camera / scale_image / scale_input / crop_input / center_image    datasets/preprocess.py:19-115
    MVSDataset (create_cams, __getitem__)                                     datasets/cas_normal_eval.py:12-182
    the result-saving block of predict_depth()                                predict.py:136-183  -> save_view_outputs

No gdal / matplotlib / imageio imports (the reference's `datasets` package needs all three at import time).
"""
from __future__ import annotations

import math
import os

import numpy as np

from .formats import (read_cameras_text, read_images_path_text, read_images_text, read_view_pair_text, save_pfm_utf8,
                      write_red_cam)

__all__ = ["scale_camera", "scale_image", "scale_input", "crop_input", "center_image", "MVSDataset", "collate",
           "save_view_outputs"]


def scale_camera(cam, scale=1):
    """cam [2,4,4] (cam[1][:3,:3] = K): focal lengths and principal point times `scale`; returns a copy."""
    out = np.copy(cam)
    for r, c in ((0, 0), (1, 1), (0, 2), (1, 2)):
        out[1][r][c] = cam[1][r][c] * scale
    return out


def scale_image(image, scale=1, interpolation="linear"):
    """cv2.resize by `scale`; upstream maps 'biculic' to nearest-neighbour (`preprocess.py:45-46`) -- kept."""
    import cv2
    flag = {"linear": cv2.INTER_LINEAR, "biculic": cv2.INTER_NEAREST}.get(interpolation)
    if flag is None:
        return None
    return cv2.resize(image, None, fx=scale, fy=scale, interpolation=flag)


def scale_input(image, cam, depth_image=None, scale=1):
    image = scale_image(image, scale=scale)
    cam = scale_camera(cam, scale=scale)
    if depth_image is None:
        return image, cam
    return image, cam, scale_image(depth_image, scale=scale, interpolation="linear")


def crop_input(image, cam, depth_image=None, max_h=384, max_w=768, resize_scale=1, base_image_size=32):
    """Centre crop to at most max_h x max_w (a side below the limit is 'rounded up' to a multiple of
    `base_image_size`, which a slice cannot grow -- as upstream); the principal point moves with the crop.
    Modifies `cam` in place and returns it, as upstream."""
    max_h, max_w = int(max_h * resize_scale), int(max_w * resize_scale)
    h, w = image.shape[0:2]
    new_h = max_h if h > max_h else int(math.ceil(h / base_image_size) * base_image_size)
    new_w = max_w if w > max_w else int(math.ceil(w / base_image_size) * base_image_size)
    top = int(math.ceil((h - new_h) / 2))
    left = int(math.ceil((w - new_w) / 2))
    image = image[top:top + new_h, left:left + new_w]
    cam[1][0][2] = cam[1][0][2] - left
    cam[1][1][2] = cam[1][1][2] - top
    if depth_image is not None:
        return image, cam, depth_image[top:top + new_h, left:left + new_w]
    return image, cam


def center_image(img, mode="mean"):
    """'standard': /255; 'mean': per-channel (x - mean) / (std + 1e-8); 'vit': ImageNet statistics."""
    if mode == "standard":
        return np.array(img, dtype=np.float32) / 255.
    if mode == "mean":
        x = np.array(img).astype(np.float32)
        var = np.var(x, axis=(0, 1), keepdims=True)
        mean = np.mean(x, axis=(0, 1), keepdims=True)
        return (x - mean) / (np.sqrt(var) + 0.00000001)
    if mode == "vit":
        x = np.array(img).astype(np.float32)
        mean = np.array([123.675, 116.28, 103.53]).astype(np.float32)
        std = np.array([58.395, 57.12, 57.375]).astype(np.float32)
        return (x - mean) / (std + 0.00000001)
    raise Exception("{}? Not implemented yet!".format(mode))


class MVSDataset:
    """`datasets/cas_normal_eval.py:MVSDataset` -- item = one reference view with its `view_num - 1` sources.

    `args` needs `min_interval, interval_scale, numdepth, resize_scale, sample_scale, max_h, max_w`
    (predict.py's argparse namespace).  Indexable and sized, so `torch.utils.data.DataLoader` takes it as is.
    """

    def __init__(self, data_folder, mode, view_num, normalize, args, **kwargs):
        assert mode in ["train", "val", "test"]
        self.data_folder = data_folder
        self.mode, self.args, self.view_num, self.normalize = mode, args, view_num, normalize
        self.min_interval = args.min_interval
        self.interval_scale = args.interval_scale
        self.num_depth = args.numdepth
        self.cam_params_dict = read_cameras_text(data_folder + "/cameras.txt")
        self.image_params_dict = read_images_text(data_folder + "/images.txt")
        self.image_paths, _ = read_images_path_text(data_folder + "/image_path.txt")
        self.sample_list = read_view_pair_text(data_folder + "/viewpair.txt", view_num)
        self.sample_num = len(self.sample_list)

    def __len__(self):
        return len(self.sample_list)

    def read_img(self, filename):
        from PIL import Image
        return Image.open(filename)

    def create_cams(self, image_params, cam_params_dict, num_depth=384, min_interval=0.1):
        """images.txt pose (XrightYup, [Rwc|twc]) -> cam [2,4,4] float32: cam[0] = Tcw (XrightYdown),
        cam[1][:3,:3] = K, cam[1][3] = [dmin, (dmax-dmin)/num_depth, num_depth, dmax]."""
        cam = np.zeros((2, 4, 4), dtype=np.float32)
        twc = np.zeros((4, 4), dtype=np.float32)
        flip_yz = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=float)
        twc[0:3, 0:3] = np.matmul(image_params.rotation_matrix, flip_yz)
        twc[0:3, 3] = image_params.project_center
        twc[3, 3] = 1.0
        cam[0, :, :] = np.linalg.inv(twc)             # float32 inverse, as upstream
        k = cam_params_dict[image_params.camera_id]
        cam[1][0][0], cam[1][1][1] = k.focallength[0], k.focallength[1]
        cam[1][0][2], cam[1][1][2] = k.x0y0[0], k.x0y0[1]
        cam[1][2][2] = 1
        dmin, dmax = image_params.depth[0], image_params.depth[1]
        cam[1][3][0] = dmin
        cam[1][3][1] = (dmax - dmin) / num_depth
        cam[1][3][2] = num_depth
        cam[1][3][3] = dmax
        return cam

    def __getitem__(self, idx):
        ids = self.sample_list[idx]
        a = self.args
        images, projs, intrs, location = [], [], [], []
        outimage = outcam = ref_path = None
        depth_min = depth_max = None
        for view in range(self.view_num):
            image_idx = ids[view]
            image = np.array(self.read_img(self.image_paths[image_idx]))
            params = self.image_params_dict[image_idx]
            cam = self.create_cams(params, self.cam_params_dict, self.num_depth, self.min_interval * self.interval_scale)
            image, cam = scale_input(image, cam, scale=a.resize_scale)
            image, cam = crop_input(image, cam, max_h=a.max_h, max_w=a.max_w, resize_scale=a.resize_scale)
            if view == 0:
                ref_path = self.image_paths[image_idx]
                outimage, outcam = image, cam
                depth_min, depth_max = cam[1][3][0], cam[1][3][3]
                h, w = image.shape[0:2]
                location = [str(w), str(h), str(params.image_id), str(params.name)]
            cam = scale_camera(cam, scale=a.sample_scale)
            k = cam[1, 0:3, 0:3]
            proj = cam[0, :, :].copy()
            proj[:3, :4] = np.matmul(k, proj[:3, :4])
            projs.append(proj)
            intrs.append(k)
            images.append(center_image(image, mode=self.normalize))

        def pyramid(full):
            half, quarter = full.copy(), full.copy()
            half[:, :2, :] = full[:, :2, :] / 2
            quarter[:, :2, :] = full[:, :2, :] / 4
            return {"stage1": quarter, "stage2": half, "stage3": full}

        return {"imgs": np.stack(images).transpose([0, 3, 1, 2]),
                "proj_matrices": pyramid(np.stack(projs)),
                "intri_matrices": pyramid(np.stack(intrs)),
                "depth_values": np.array([depth_min, depth_max], dtype=np.float32),
                "outimage": outimage,
                "outcam": outcam,
                "ref_image_path": ref_path,
                "outlocation": location}


def collate(item):
    """What `DataLoader(batch_size=1)` hands `predict.py`: arrays become tensors with a leading batch axis,
    strings become 1-element lists."""
    import torch

    def go(v):
        if isinstance(v, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(v)).unsqueeze(0)
        if isinstance(v, dict):
            return {k: go(x) for k, x in v.items()}
        if isinstance(v, list):
            return [go(x) for x in v]
        if isinstance(v, str):
            return [v]
        return v
    return {k: go(v) for k, v in item.items()}


def _colour_png(path, image):
    """`plt.imsave(path, image, format='png')` when matplotlib is there (viridis, min-max normalised); the same
    normalisation as an 8-bit grey PNG through PIL otherwise -- display only, nothing downstream reads it."""
    try:
        import matplotlib.pyplot as plt
        plt.imsave(path, image, format="png")
        return
    except ImportError:
        pass
    from PIL import Image
    lo, hi = float(np.nanmin(image)), float(np.nanmax(image))
    grey = (image - lo) / (hi - lo) if hi > lo else np.zeros_like(image)
    Image.fromarray(np.uint8(np.clip(grey, 0, 1) * 255 + 0.5)).save(path, format="PNG")


def save_view_outputs(output_folder, depth_est, photometric_confidence, ref_cam, out_location, ref_path,
                      display=True):
    """predict.py:136-183: <name>_init.pfm, <name>_prob.pfm, <name>.txt (+ color/<name>_{init,prob}.png).
    `out_location` = [w, h, view id, image name]; returns the four / six paths written."""
    depth_est = np.float32(np.squeeze(depth_est))
    prob = np.float32(np.squeeze(photometric_confidence))
    name = os.path.splitext(str(out_location[3]))[0]
    paths = {"depth": output_folder + ("/%s_init.pfm" % name), "prob": output_folder + ("/%s_prob.pfm" % name),
             "cam": output_folder + ("/%s.txt" % name)}
    os.makedirs(os.path.dirname(paths["depth"]), exist_ok=True)
    if display:
        shown = np.float32(36000) - depth_est                     # predict.py:158-159
        for i in range(shown.shape[1]):                           # per column: inf/nan -> (column minimum) - 1
            col = shown[:, i]
            col[np.isinf(col)] = np.nan
            col[np.isnan(col)] = np.nanmin(col) - 1
        paths["depth_png"] = output_folder + ("/color/%s_init.png" % name)
        paths["prob_png"] = output_folder + ("/color/%s_prob.png" % name)
        os.makedirs(os.path.dirname(paths["depth_png"]), exist_ok=True)
        _colour_png(paths["depth_png"], shown)
        _colour_png(paths["prob_png"], np.nan_to_num(prob).clip(0, 1))
    save_pfm_utf8(paths["depth"], depth_est)
    save_pfm_utf8(paths["prob"], prob)
    write_red_cam(paths["cam"], ref_cam, out_location, ref_path)
    return paths
