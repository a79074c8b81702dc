"""COLMAP sparse-model binary I/O (cameras.bin / images.bin / points3D.bin).

Host-side scene loading for `Splatter(colmap_path, image_path, ...)`, the constructor form the
reference's train.py uses (reference splatter.py:363-383 via utils.py:111-141, :181-224,
:259-294).  Written against the published COLMAP binary layout (little endian); it also WRITES
models so that tests and benchmarks can build synthetic datasets without COLMAP.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict

import numpy as np

# model id -> (name, number of intrinsic parameters)
CAMERA_MODELS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5),
                 4: ("OPENCV", 8), 5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5),
                 8: ("SIMPLE_RADIAL_FISHEYE", 4), 9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12)}
MODEL_IDS = {v[0]: k for k, v in CAMERA_MODELS.items()}


@dataclass(frozen=True)
class Camera:
    id: int
    model: str
    width: int
    height: int
    params: np.ndarray


@dataclass(frozen=True)
class Image:
    id: int
    qvec: np.ndarray      # world->camera rotation, wxyz
    tvec: np.ndarray      # world->camera translation
    camera_id: int
    name: str


@dataclass(frozen=True)
class Point3D:
    id: int
    xyz: np.ndarray
    rgb: np.ndarray       # uint8
    error: float


def _read(f, fmt):
    size = struct.calcsize("<" + fmt)
    data = f.read(size)
    if len(data) != size:
        raise EOFError("truncated COLMAP file")
    return struct.unpack("<" + fmt, data)


def read_cameras_binary(path) -> Dict[int, Camera]:
    cams = {}
    with open(path, "rb") as f:
        (n,) = _read(f, "Q")
        for _ in range(n):
            cid, model_id, w, h = _read(f, "iiQQ")
            name, npar = CAMERA_MODELS[model_id]
            params = np.array(_read(f, "d" * npar), dtype=np.float64)
            cams[cid] = Camera(cid, name, int(w), int(h), params)
    return cams


def read_images_binary(path) -> Dict[int, Image]:
    imgs = {}
    with open(path, "rb") as f:
        (n,) = _read(f, "Q")
        for _ in range(n):
            vals = _read(f, "idddddddi")
            iid, q, t, cid = vals[0], np.array(vals[1:5]), np.array(vals[5:8]), vals[8]
            name = b""
            while True:
                ch = f.read(1)
                if ch in (b"\x00", b""):
                    break
                name += ch
            (n2d,) = _read(f, "Q")
            f.seek(24 * n2d, os.SEEK_CUR)                      # (x, y, point3D_id) per observation
            imgs[iid] = Image(iid, q, t, cid, name.decode("utf-8"))
    return imgs


def read_points3d_binary(path) -> Dict[int, Point3D]:
    pts = {}
    with open(path, "rb") as f:
        (n,) = _read(f, "Q")
        for _ in range(n):
            vals = _read(f, "QdddBBBd")
            pid, xyz, rgb, err = vals[0], np.array(vals[1:4]), np.array(vals[4:7], dtype=np.uint8), vals[7]
            (tl,) = _read(f, "Q")
            f.seek(8 * tl, os.SEEK_CUR)                        # (image_id, point2D_idx) track entries
            pts[pid] = Point3D(pid, xyz, rgb, err)
    return pts


# ---- writers (tests / synthetic datasets) ----------------------------------------------------
def write_cameras_binary(path, cams: Dict[int, Camera]):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(cams)))
        for c in cams.values():
            f.write(struct.pack("<iiQQ", c.id, MODEL_IDS[c.model], c.width, c.height))
            f.write(struct.pack("<" + "d" * len(c.params), *[float(x) for x in c.params]))


def write_images_binary(path, imgs: Dict[int, Image]):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(imgs)))
        for im in imgs.values():
            f.write(struct.pack("<idddddddi", im.id, *[float(x) for x in im.qvec], *[float(x) for x in im.tvec],
                                im.camera_id))
            f.write(im.name.encode("utf-8") + b"\x00")
            f.write(struct.pack("<Q", 0))


def write_points3d_binary(path, pts: Dict[int, Point3D]):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(pts)))
        for p in pts.values():
            f.write(struct.pack("<QdddBBBd", p.id, *[float(x) for x in p.xyz], *[int(x) for x in p.rgb],
                                float(p.error)))
            f.write(struct.pack("<Q", 0))


def qvec_to_rotmat(q):
    """wxyz -> 3x3 (same convention as reference utils.py:318-333 / gaussian.cu:1231-1245)."""
    w, x, y, z = [float(v) for v in q]
    return np.array([
        [1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w],
        [2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w],
        [2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y]], dtype=np.float64)


def rotmat_to_qvec(R):
    """3x3 -> wxyz (w >= 0)."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q if q[0] >= 0 else -q
