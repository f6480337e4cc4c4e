"""On-disk formats either side of the hot path -- host-side Python, as upstream.

Same names, argument order, return values and error behaviour as the reference's
`mvs/mvs_cas/datasets/data_io.py`, minus its hard imports of `gdal` and `matplotlib` (`data_io.py:12,14`), which is
what keeps `predict.py` from running on a stock PyTorch image:

    Camera, Photo                      data_io.py:17-45
    read_cameras_text                  data_io.py:48-69      cameras.txt
    read_images_text                   data_io.py:72-93      images.txt
    read_images_path_text              data_io.py:96-111     image_path.txt
    read_view_pair_text                data_io.py:114-130    viewpair.txt
    load_pfm_utf8 / save_pfm_utf8      data_io.py:164-223    <name>_init.pfm, <name>_prob.pfm
    write_red_cam                      data_io.py:291-314    <name>.txt
    read_red_cam                       the inverse (what fuse/fusion_3d_normal.py reads back)
"""
from __future__ import annotations

import re
import sys

import numpy as np

__all__ = ["Camera", "Photo", "read_cameras_text", "read_images_text", "read_images_path_text",
           "read_view_pair_text", "load_pfm_utf8", "save_pfm_utf8", "load_pfm", "save_pfm", "write_red_cam",
           "read_red_cam"]


class Camera:
    """One line of cameras.txt: `id width height pixelsize fx fy x0 y0 [distortion ...]`."""

    def __init__(self, camera_id=None, size=None, pixelsize=None, focallength=None, x0y0=None, distortion=None):
        self.camera_id = camera_id
        self.size = size                  # [width, height]
        self.pixelsize = pixelsize
        self.focallength = focallength    # [fx, fy]
        self.x0y0 = x0y0                  # [x0, y0]
        self.distortion = distortion      # [k1, k2, k3, p1, p2]


class Photo:
    """One line of images.txt: `id camera_id R(9, row major) C(3) dmin dmax name` (XrightYup, [Rwc|twc])."""

    def __init__(self, image_id=None, camera_id=None, rotation_matrix=None, project_center=None, depth=None,
                 name=None, camera_coordinate_type="XrightYup", rotation_type="Rwc", translation_type="twc"):
        self.image_id = image_id
        self.camera_id = camera_id
        self.name = name
        self.rotation_matrix = rotation_matrix
        self.project_center = project_center
        self.depth = depth
        self.camera_coordinate_type = camera_coordinate_type
        self.rotation_type = rotation_type
        self.translation_type = translation_type


def _records(path):
    """Whitespace-split fields of every line that is neither blank nor a `#` comment."""
    with open(path, "r") as fid:
        for raw in fid:
            line = raw.strip()
            if line and not line.startswith("#"):
                yield line.split()


def _floats(tokens):
    return np.array(tuple(float(t) for t in tokens))


def read_cameras_text(path):
    cams = {}
    for f in _records(path):
        cid = int(f[0])
        k = _floats(f[4:8])
        cams[cid] = Camera(camera_id=cid, size=[int(f[1]), int(f[2])], pixelsize=float(f[3]),
                           focallength=[k[0], k[1]], x0y0=[k[2], k[3]], distortion=_floats(f[8:]))
    return cams


def read_images_text(path):
    images = {}
    for f in _records(path):
        iid = int(f[0])
        images[iid] = Photo(image_id=iid, camera_id=int(f[1]), rotation_matrix=_floats(f[2:11]).reshape(3, 3),
                            project_center=_floats(f[11:14]), depth=_floats(f[14:16]), name=f[16])
    return images


def read_images_path_text(path):
    """image_path.txt: a count, then `index name path` triples, all whitespace separated."""
    with open(path) as fid:
        tok = fid.read().split()
    paths, names = {}, {}
    for i in range(int(tok[0])):
        index = int(tok[3 * i + 1])
        names[index] = tok[3 * i + 2]
        paths[index] = tok[3 * i + 3]
    return paths, names


def read_view_pair_text(pair_path, view_num):
    """viewpair.txt: a count; per reference view its id on one line and `n id score id score ...` on the next.
    Views without sources are dropped; short source lists are padded with their first entry up to `view_num`
    entries (upstream pads to view_num, not view_num - 1)."""
    metas = []
    with open(pair_path) as f:
        for _ in range(int(f.readline())):
            ref = int(f.readline().rstrip())
            srcs = [int(x) for x in f.readline().rstrip().split()[1::2]]
            if not srcs:
                continue
            if len(srcs) < view_num:
                print("{}< num_views:{}".format(len(srcs), view_num))
                srcs = srcs + [srcs[0]] * (view_num - len(srcs))
            metas.append([ref] + srcs)
    return metas


_DIMS = re.compile(r"^(\d+)\s(\d+)\s$")


def load_pfm_utf8(filename):
    """-> (array [H,W] or [H,W,3] float32, top row first; scale).  Raises on a bad magic or header."""
    with open(filename, "rb") as fid:
        magic = fid.readline().decode("utf-8").rstrip()
        if magic not in ("PF", "Pf"):
            raise Exception("Not a PFM file.")
        dims = _DIMS.match(fid.readline().decode("utf-8"))
        if not dims:
            raise Exception("Malformed PFM header.")
        width, height = int(dims.group(1)), int(dims.group(2))
        scale = float(fid.readline().rstrip())
        order = "<" if scale < 0 else ">"
        data = np.fromfile(fid, order + "f")
    shape = (height, width, 3) if magic == "PF" else (height, width)
    return np.flipud(data.reshape(shape)), abs(scale)


def save_pfm_utf8(filename, image, scale=1):
    """float32 [H,W], [H,W,1] or [H,W,3]; rows are stored bottom-up, the sign of `scale` carries the byte order."""
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        magic = "PF"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        magic = "Pf"
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    order = image.dtype.byteorder
    if order == "<" or (order == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as fid:
        fid.write(("%s\n%d %d\n" % (magic, image.shape[1], image.shape[0])).encode("utf-8"))
        fid.write(("%f\n" % scale).encode("utf-8"))
        np.flipud(image).tofile(fid)


load_pfm = load_pfm_utf8
save_pfm = save_pfm_utf8


def write_red_cam(file, cam, location, ref_path):
    """cam [2,4,4]: cam[0] = Tcw, cam[1][:3,:3] = K, cam[1][3] = [dmin, interval, num_depth, dmax]."""
    out = ["extrinsic: XrightYdown, [Rcw|tcw]\n"]
    for i in range(4):
        out.append("".join(str(cam[0][i][j]) + " " for j in range(4)) + "\n")
    out.append("\nintrinsic\n")
    for i in range(3):
        out.append("".join(str(cam[1][i][j]) + " " for j in range(3)) + "\n")
    out.append("\n" + " ".join(str(cam[1][3][j]) for j in range(4)) + "\n\n")
    out.append("".join(str(word) + " " for word in location) + str(ref_path) + "\n")
    with open(file, "w") as f:
        f.write("".join(out))


def read_red_cam(file):
    """-> (cam [2,4,4] float32, location [w, h, view id, name], ref_path): what `write_red_cam` wrote."""
    with open(file) as f:
        tok = f.read().split()
    at = tok.index("[Rcw|tcw]") + 1
    cam = np.zeros((2, 4, 4), dtype=np.float32)
    cam[0] = np.array(tok[at:at + 16], dtype=np.float64).reshape(4, 4)
    at = tok.index("intrinsic", at) + 1
    cam[1, :3, :3] = np.array(tok[at:at + 9], dtype=np.float64).reshape(3, 3)
    cam[1, 3] = np.array(tok[at + 9:at + 13], dtype=np.float64)
    rest = tok[at + 13:]
    return cam, rest[:4], (rest[4] if len(rest) > 4 else "")
