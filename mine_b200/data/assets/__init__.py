"""Static evaluation-pair definitions shipped with the reference for datasets whose loaders were never
released: RealEstate10K test/validation pairs and the Flowers light-field camera grid.
Stored as compressed ``.npz`` (same content as upstream's JSON-lines / text tables, ~10x smaller).

* ``realestate10k_pairs(split)`` -> list of dicts with ``sequence_id`` and, for each of ``src``,
  ``tgt_5_frames``, ``tgt_10_frames``, ``tgt_random``: ``frame_ts``, normalised intrinsics ``[fx, fy, cx, cy]``
  and the 3x4 world->camera pose.
* ``flowers_lightfield()`` -> 8x8 grid view ids, normalised intrinsics, 3x4 poses, train / test file lists.
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_VIEWS = ("src", "tgt_5_frames", "tgt_10_frames", "tgt_random")


def realestate10k_pairs(split: str = "test") -> List[Dict]:
    if split not in ("test", "validation"):
        raise ValueError("split must be 'test' or 'validation'")
    with np.load(os.path.join(_HERE, f"realestate10k_{split}_pairs.npz")) as z:
        seqs, ts, intr, pose = z["sequence_id"], z["frame_ts"], z["intrinsics"], z["pose"]   # decompress once
    out = []
    for i, seq in enumerate(seqs):
        item = {"sequence_id": str(seq)}
        for j, name in enumerate(_VIEWS):
            item[name] = {"frame_ts": int(ts[i, j]), "intrinsics": intr[i, j], "pose": pose[i, j].reshape(3, 4)}
        out.append(item)
    return out


def intrinsics_matrix(normalised, width: int, height: int) -> np.ndarray:
    """``[fx, fy, cx, cy]`` normalised by image size -> 3x3 K in pixels."""
    fx, fy, cx, cy = [float(v) for v in normalised]
    return np.array([[fx * width, 0, cx * width], [0, fy * height, cy * height], [0, 0, 1]], dtype=np.float32)


def relative_pose(pose_src: np.ndarray, pose_tgt: np.ndarray) -> np.ndarray:
    """``G_src_tgt`` (4x4) from two 3x4 world->camera poses."""
    to4 = lambda p: np.vstack([p, [0, 0, 0, 1]])
    return (to4(pose_src) @ np.linalg.inv(to4(pose_tgt))).astype(np.float32)


def flowers_lightfield() -> Dict:
    z = np.load(os.path.join(_HERE, "flowers_lightfield.npz"))
    return {"view_id": [str(v) for v in z["view_id"]], "intrinsics": z["intrinsics"], "distortion": z["distortion"],
            "pose": z["pose"],
            "train": [str(v) for v in z["train"]], "test": [str(v) for v in z["test"]]}


def export_upstream_layout(dst_root: str) -> List[str]:
    """Write the tables in the directory layout and text formats the reference keeps them in
    (``input_pipelines/realestate10k/test_data_jsons/{test,validation}_pairs.json`` JSON lines,
    ``input_pipelines/flowers/cam_params.txt``, ``input_pipelines/flowers/dataset_list/{train,test}.list``), for
    tools written against those files.  Returns the paths written."""
    import json
    written = []
    d = os.path.join(dst_root, "realestate10k", "test_data_jsons")
    os.makedirs(d, exist_ok=True)
    keys = {"src": "src_img_obj", "tgt_5_frames": "tgt_img_obj_5_frames", "tgt_10_frames": "tgt_img_obj_10_frames",
            "tgt_random": "tgt_img_obj_random"}
    for split in ("test", "validation"):
        path = os.path.join(d, f"{split}_pairs.json")
        with open(path, "w") as f:
            for item in realestate10k_pairs(split):
                row = {"sequence_id": item["sequence_id"]}
                for ours, theirs in keys.items():
                    v = item[ours]
                    row[theirs] = {"sequence_id": item["sequence_id"],
                                   "camera_intrinsics": [round(float(x), 9) for x in v["intrinsics"]],
                                   "camera_pose": [round(float(x), 9) for x in np.asarray(v["pose"]).reshape(-1)],
                                   "frame_ts": str(v["frame_ts"])}
                f.write(json.dumps(row) + "\n")
        written.append(path)
    lf = flowers_lightfield()
    d = os.path.join(dst_root, "flowers")
    os.makedirs(os.path.join(d, "dataset_list"), exist_ok=True)
    path = os.path.join(d, "cam_params.txt")
    with open(path, "w") as f:
        for vid, intr, dist, pose in zip(lf["view_id"], lf["intrinsics"], lf["distortion"], lf["pose"]):
            vals = [float(x) for x in intr] + [float(x) for x in dist] + [float(x) for x in np.asarray(pose).reshape(-1)]
            f.write(vid + " " + " ".join("%.6f" % x for x in vals) + "\n")
    written.append(path)
    for split in ("train", "test"):
        path = os.path.join(d, "dataset_list", f"{split}.list")
        with open(path, "w") as f:
            f.write("\n".join(lf[split]) + "\n")
        written.append(path)
    return written


if __name__ == "__main__":
    import sys
    for p in export_upstream_layout(sys.argv[1] if len(sys.argv) > 1 else "input_pipelines"):
        print(p)
