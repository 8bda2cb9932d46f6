"""COLMAP sparse-model I/O (cameras / images / points3D, ``.bin`` and ``.txt``) + sqlite database.

Capability parity with the vendored ``input_pipelines/colmap_utils.py`` and ``database.py`` of the
reference (only ``read_model`` and ``qvec2rotmat`` are on its training path).  Written
against the COLMAP file-format description, table-driven, with a small writer so tests can build
synthetic scenes on disk.
"""
from __future__ import annotations

import os
import sqlite3
import struct
from dataclasses import dataclass
from typing import BinaryIO, Dict, Tuple

import numpy as np

# model id -> (name, number of params)
CAMERA_MODELS = {
    0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5),
    4: ("OPENCV", 8), 5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5),
    8: ("SIMPLE_RADIAL_FISHEYE", 4), 9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12),
}
CAMERA_MODEL_IDS = {name: mid for mid, (name, _) in CAMERA_MODELS.items()}


@dataclass
class Camera:
    id: int
    model: str
    width: int
    height: int
    params: np.ndarray


@dataclass
class Image:
    id: int
    qvec: np.ndarray           # (w, x, y, z) world -> camera rotation
    tvec: np.ndarray
    camera_id: int
    name: str
    xys: np.ndarray            # N x 2 keypoints
    point3D_ids: np.ndarray    # N, -1 where untriangulated

    def qvec2rotmat(self):
        return qvec2rotmat(self.qvec)


@dataclass
class Point3D:
    id: int
    xyz: np.ndarray
    rgb: np.ndarray
    error: float
    image_ids: np.ndarray
    point2D_idxs: np.ndarray


def qvec2rotmat(q) -> np.ndarray:
    w, x, y, z = [float(v) for v in q]
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rotmat2qvec(r) -> np.ndarray:
    r = np.asarray(r, dtype=np.float64)
    k = np.array([
        [r[0, 0] - r[1, 1] - r[2, 2], 0, 0, 0],
        [r[1, 0] + r[0, 1], r[1, 1] - r[0, 0] - r[2, 2], 0, 0],
        [r[2, 0] + r[0, 2], r[2, 1] + r[1, 2], r[2, 2] - r[0, 0] - r[1, 1], 0],
        [r[2, 1] - r[1, 2], r[0, 2] - r[2, 0], r[1, 0] - r[0, 1], r[0, 0] + r[1, 1] + r[2, 2]]]) / 3.0
    vals, vecs = np.linalg.eigh(k)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    return -q if q[0] < 0 else q


def _rd(f: BinaryIO, fmt: str):
    size = struct.calcsize("<" + fmt)
    return struct.unpack("<" + fmt, f.read(size))


# ---- binary ---------------------------------------------------------------------------------
def read_cameras_binary(path: str) -> Dict[int, Camera]:
    out = {}
    with open(path, "rb") as f:
        (n,) = _rd(f, "Q")
        for _ in range(n):
            cid, mid, w, h = _rd(f, "iiQQ")
            name, npar = CAMERA_MODELS[mid]
            out[cid] = Camera(cid, name, w, h, np.array(_rd(f, "d" * npar)))
    return out


def read_images_binary(path: str) -> Dict[int, Image]:
    out = {}
    with open(path, "rb") as f:
        (n,) = _rd(f, "Q")
        for _ in range(n):
            vals = _rd(f, "idddddddi")
            iid, q, t, cid = vals[0], np.array(vals[1:5]), np.array(vals[5:8]), vals[8]
            name = b""
            while True:
                ch = f.read(1)
                if ch == b"\x00":
                    break
                name += ch
            (m,) = _rd(f, "Q")
            raw = np.frombuffer(f.read(24 * m), dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("id", "<i8")]))
            out[iid] = Image(iid, q, t, cid, name.decode("utf-8"),
                             np.stack([raw["x"], raw["y"]], 1) if m else np.zeros((0, 2)),
                             raw["id"].astype(np.int64))
    return out


def read_points3D_binary(path: str) -> Dict[int, Point3D]:
    out = {}
    with open(path, "rb") as f:
        (n,) = _rd(f, "Q")
        for _ in range(n):
            vals = _rd(f, "QdddBBBd")
            (tl,) = _rd(f, "Q")
            tr = np.frombuffer(f.read(8 * tl), dtype="<i4").reshape(-1, 2) if tl else np.zeros((0, 2), np.int32)
            out[vals[0]] = Point3D(vals[0], np.array(vals[1:4]), np.array(vals[4:7], dtype=np.uint8), vals[7],
                                   tr[:, 0].copy(), tr[:, 1].copy())
    return out


def write_cameras_binary(cams: Dict[int, Camera], path: str) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(cams)))
        for c in cams.values():
            f.write(struct.pack("<iiQQ", c.id, CAMERA_MODEL_IDS[c.model], c.width, c.height))
            f.write(struct.pack("<" + "d" * len(c.params), *[float(p) for p in c.params]))


def write_images_binary(imgs: Dict[int, Image], path: str) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(imgs)))
        for im in imgs.values():
            f.write(struct.pack("<idddddddi", im.id, *[float(v) for v in im.qvec], *[float(v) for v in im.tvec],
                                im.camera_id))
            f.write(im.name.encode("utf-8") + b"\x00")
            f.write(struct.pack("<Q", len(im.point3D_ids)))
            for (x, y), pid in zip(im.xys, im.point3D_ids):
                f.write(struct.pack("<ddq", float(x), float(y), int(pid)))


def write_points3D_binary(pts: Dict[int, Point3D], path: str) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(pts)))
        for p in pts.values():
            f.write(struct.pack("<QdddBBBd", p.id, *[float(v) for v in p.xyz], *[int(v) for v in p.rgb], float(p.error)))
            f.write(struct.pack("<Q", len(p.image_ids)))
            for i, j in zip(p.image_ids, p.point2D_idxs):
                f.write(struct.pack("<ii", int(i), int(j)))


# ---- text -----------------------------------------------------------------------------------
def _lines(path):
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if line and not line.startswith("#"):
                yield line


def read_cameras_text(path: str) -> Dict[int, Camera]:
    out = {}
    for line in _lines(path):
        e = line.split()
        out[int(e[0])] = Camera(int(e[0]), e[1], int(e[2]), int(e[3]), np.array([float(v) for v in e[4:]]))
    return out


def read_images_text(path: str) -> Dict[int, Image]:
    out = {}
    it = iter(_lines_keep_empty(path))
    for head in it:
        e = head.split()
        pts = next(it, "").split()
        xys = np.array([[float(pts[i]), float(pts[i + 1])] for i in range(0, len(pts), 3)]).reshape(-1, 2)
        ids = np.array([int(pts[i + 2]) for i in range(0, len(pts), 3)], dtype=np.int64)
        out[int(e[0])] = Image(int(e[0]), np.array([float(v) for v in e[1:5]]), np.array([float(v) for v in e[5:8]]),
                               int(e[8]), e[9], xys, ids)
    return out


def _lines_keep_empty(path):
    """images.txt alternates header / keypoint lines; the keypoint line may be empty."""
    with open(path, "r") as f:
        for line in f:
            if line.startswith("#"):
                continue
            yield line.rstrip("\n")


def read_points3D_text(path: str) -> Dict[int, Point3D]:
    out = {}
    for line in _lines(path):
        e = line.split()
        tr = np.array([int(v) for v in e[8:]], dtype=np.int32).reshape(-1, 2)
        out[int(e[0])] = Point3D(int(e[0]), np.array([float(v) for v in e[1:4]]),
                                 np.array([int(v) for v in e[4:7]], dtype=np.uint8), float(e[7]), tr[:, 0], tr[:, 1])
    return out


def write_model_text(cams, imgs, pts, path: str) -> None:
    with open(os.path.join(path, "cameras.txt"), "w") as f:
        for c in cams.values():
            f.write(" ".join([str(c.id), c.model, str(c.width), str(c.height)] + [repr(float(p)) for p in c.params]) + "\n")
    with open(os.path.join(path, "images.txt"), "w") as f:
        for im in imgs.values():
            f.write(" ".join([str(im.id)] + [repr(float(v)) for v in im.qvec] + [repr(float(v)) for v in im.tvec]
                             + [str(im.camera_id), im.name]) + "\n")
            f.write(" ".join(f"{float(x)!r} {float(y)!r} {int(pid)}" for (x, y), pid in zip(im.xys, im.point3D_ids)) + "\n")
    with open(os.path.join(path, "points3D.txt"), "w") as f:
        for p in pts.values():
            tr = " ".join(f"{int(i)} {int(j)}" for i, j in zip(p.image_ids, p.point2D_idxs))
            f.write(" ".join([str(p.id)] + [repr(float(v)) for v in p.xyz] + [str(int(v)) for v in p.rgb]
                             + [repr(float(p.error)), tr]).strip() + "\n")


def read_model(path: str, ext: str = ".bin"):
    """``(cameras, images, points3D)`` dicts keyed by id."""
    if ext == ".bin":
        return (read_cameras_binary(os.path.join(path, "cameras.bin")),
                read_images_binary(os.path.join(path, "images.bin")),
                read_points3D_binary(os.path.join(path, "points3D.bin")))
    return (read_cameras_text(os.path.join(path, "cameras.txt")),
            read_images_text(os.path.join(path, "images.txt")),
            read_points3D_text(os.path.join(path, "points3D.txt")))


def write_model(cams, imgs, pts, path: str, ext: str = ".bin") -> None:
    os.makedirs(path, exist_ok=True)
    if ext == ".bin":
        write_cameras_binary(cams, os.path.join(path, "cameras.bin"))
        write_images_binary(imgs, os.path.join(path, "images.bin"))
        write_points3D_binary(pts, os.path.join(path, "points3D.bin"))
    else:
        write_model_text(cams, imgs, pts, path)


# ---- sqlite database (feature/match store; unused by training, kept for tool parity) -----------
MAX_IMAGE_ID = 2 ** 31 - 1


def image_ids_to_pair_id(a: int, b: int) -> int:
    if a > b:
        a, b = b, a
    return a * MAX_IMAGE_ID + b


def pair_id_to_image_ids(pair_id: int) -> Tuple[int, int]:
    b = pair_id % MAX_IMAGE_ID
    return (pair_id - b) // MAX_IMAGE_ID, b


class COLMAPDatabase(sqlite3.Connection):
    """Minimal COLMAP ``database.db`` wrapper: cameras, images, keypoints, matches."""

    SCHEMA = """
    CREATE TABLE IF NOT EXISTS cameras (camera_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, model INTEGER NOT NULL,
        width INTEGER NOT NULL, height INTEGER NOT NULL, params BLOB, prior_focal_length INTEGER NOT NULL);
    CREATE TABLE IF NOT EXISTS images (image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE,
        camera_id INTEGER NOT NULL, prior_qw REAL, prior_qx REAL, prior_qy REAL, prior_qz REAL,
        prior_tx REAL, prior_ty REAL, prior_tz REAL);
    CREATE TABLE IF NOT EXISTS keypoints (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
        cols INTEGER NOT NULL, data BLOB);
    CREATE TABLE IF NOT EXISTS descriptors (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
        cols INTEGER NOT NULL, data BLOB);
    CREATE TABLE IF NOT EXISTS matches (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
        cols INTEGER NOT NULL, data BLOB);
    CREATE TABLE IF NOT EXISTS two_view_geometries (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
        cols INTEGER NOT NULL, data BLOB, config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB);
    """

    @staticmethod
    def connect(path: str) -> "COLMAPDatabase":
        return sqlite3.connect(path, factory=COLMAPDatabase)

    def create_tables(self):
        self.executescript(self.SCHEMA)

    def add_camera(self, model: int, width: int, height: int, params, prior_focal_length: bool = False, camera_id=None) -> int:
        cur = self.execute("INSERT INTO cameras VALUES (?, ?, ?, ?, ?, ?)",
                           (camera_id, model, width, height, np.asarray(params, np.float64).tobytes(), prior_focal_length))
        return cur.lastrowid

    def add_image(self, name: str, camera_id: int, prior_q=(np.nan,) * 4, prior_t=(np.nan,) * 3, image_id=None) -> int:
        cur = self.execute("INSERT INTO images VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                           (image_id, name, camera_id, *[float(v) for v in prior_q], *[float(v) for v in prior_t]))
        return cur.lastrowid

    def add_keypoints(self, image_id: int, keypoints) -> None:
        k = np.asarray(keypoints, np.float32)
        self.execute("INSERT INTO keypoints VALUES (?, ?, ?, ?)", (image_id, k.shape[0], k.shape[1], k.tobytes()))

    def add_descriptors(self, image_id: int, descriptors) -> None:
        d = np.ascontiguousarray(descriptors, np.uint8)
        self.execute("INSERT INTO descriptors VALUES (?, ?, ?, ?)", (image_id, d.shape[0], d.shape[1], d.tobytes()))

    def add_matches(self, image_id1: int, image_id2: int, matches) -> None:
        m = np.asarray(matches, np.uint32)
        if image_id1 > image_id2:
            m = m[:, ::-1]
        self.execute("INSERT INTO matches VALUES (?, ?, ?, ?)",
                     (image_ids_to_pair_id(image_id1, image_id2), m.shape[0], m.shape[1], np.ascontiguousarray(m).tobytes()))

    def read_keypoints(self, image_id: int) -> np.ndarray:
        r, c, data = self.execute("SELECT rows, cols, data FROM keypoints WHERE image_id=?", (image_id,)).fetchone()
        return np.frombuffer(data, np.float32).reshape(r, c)

    def read_matches(self, image_id1: int, image_id2: int) -> np.ndarray:
        row = self.execute("SELECT rows, cols, data FROM matches WHERE pair_id=?",
                           (image_ids_to_pair_id(image_id1, image_id2),)).fetchone()
        return np.frombuffer(row[2], np.uint32).reshape(row[0], row[1])


def main(argv=None):
    """Round-trip CLI: ``python -m mine_b200.data.colmap --input_model DIR --input_format .bin
    --output_model DIR --output_format .txt`` (reference ``colmap_utils.py:481-503``)."""
    import argparse
    ap = argparse.ArgumentParser(description="Read and write COLMAP binary and text models")
    ap.add_argument("--input_model", required=True)
    ap.add_argument("--input_format", choices=[".bin", ".txt"], default=".bin")
    ap.add_argument("--output_model")
    ap.add_argument("--output_format", choices=[".bin", ".txt"], default=".txt")
    a = ap.parse_args(argv)
    cams, imgs, pts = read_model(a.input_model, a.input_format)
    print("num_cameras:", len(cams))
    print("num_images:", len(imgs))
    print("num_points3D:", len(pts))
    if a.output_model:
        write_model(cams, imgs, pts, a.output_model, a.output_format)


if __name__ == "__main__":
    main()
