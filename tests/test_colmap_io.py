"""CPU tests of the COLMAP binary reader/writer behind Splatter(colmap_path, image_path)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting_b200"))
import colmap_io as C  # noqa: E402


def _model():
    cams = {1: C.Camera(1, "PINHOLE", 640, 360, np.array([500.0, 505.0, 320.0, 180.0])),
            2: C.Camera(2, "SIMPLE_PINHOLE", 320, 240, np.array([300.0, 160.0, 120.0]))}
    rng = np.random.default_rng(0)
    imgs = {}
    for i in range(1, 5):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if q[0] < 0:
            q = -q
        imgs[i] = C.Image(i, q, rng.normal(size=3), 1 + i % 2, f"frame_{i:03d}.png")
    pts = {j + 10: C.Point3D(j + 10, rng.normal(size=3), rng.integers(0, 256, 3).astype(np.uint8), 0.5) for j in range(50)}
    return cams, imgs, pts


def test_round_trip(tmp_path):
    cams, imgs, pts = _model()
    C.write_cameras_binary(tmp_path / "cameras.bin", cams)
    C.write_images_binary(tmp_path / "images.bin", imgs)
    C.write_points3d_binary(tmp_path / "points3D.bin", pts)
    c2 = C.read_cameras_binary(tmp_path / "cameras.bin")
    i2 = C.read_images_binary(tmp_path / "images.bin")
    p2 = C.read_points3d_binary(tmp_path / "points3D.bin")
    assert sorted(c2) == [1, 2] and c2[1].model == "PINHOLE" and c2[2].model == "SIMPLE_PINHOLE"
    assert (c2[1].width, c2[1].height) == (640, 360) and np.allclose(c2[1].params, cams[1].params)
    for k in imgs:
        assert i2[k].name == imgs[k].name and i2[k].camera_id == imgs[k].camera_id
        assert np.allclose(i2[k].qvec, imgs[k].qvec) and np.allclose(i2[k].tvec, imgs[k].tvec)
    for k in pts:
        assert np.allclose(p2[k].xyz, pts[k].xyz) and (p2[k].rgb == pts[k].rgb).all()


def test_reader_skips_observations_and_tracks(tmp_path):
    """Real COLMAP files carry 2-D observations / tracks; the reader must skip them correctly."""
    import struct
    with open(tmp_path / "images.bin", "wb") as f:
        f.write(struct.pack("<Q", 1))
        f.write(struct.pack("<idddddddi", 7, 1, 0, 0, 0, 0.1, 0.2, 0.3, 1))
        f.write(b"a.jpg\x00")
        f.write(struct.pack("<Q", 2))
        f.write(struct.pack("<ddq", 1.0, 2.0, 5) + struct.pack("<ddq", 3.0, 4.0, -1))
    with open(tmp_path / "points3D.bin", "wb") as f:
        f.write(struct.pack("<Q", 2))
        for pid in (3, 4):
            f.write(struct.pack("<QdddBBBd", pid, 1.0, 2.0, 3.0, 10, 20, 30, 0.1))
            f.write(struct.pack("<Q", 2) + struct.pack("<iiii", 1, 0, 2, 5))
    im = C.read_images_binary(tmp_path / "images.bin")
    assert im[7].name == "a.jpg" and np.allclose(im[7].tvec, [0.1, 0.2, 0.3])
    pt = C.read_points3d_binary(tmp_path / "points3D.bin")
    assert sorted(pt) == [3, 4] and (pt[4].rgb == [10, 20, 30]).all()


def test_quaternion_rotation_round_trip():
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if q[0] < 0:
            q = -q
        R = C.qvec_to_rotmat(q)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
        assert np.allclose(C.rotmat_to_qvec(R), q, atol=1e-10)
