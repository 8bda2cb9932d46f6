"""Offline helper: write ``images_<ratio>/`` down-sampled copies for every LLFF scene.

    python input_pipelines/llff/misc/resize_nerf_llff_images.py /data/nerf_llff_data --ratio 7.875 [--val_every 8]

``--val_every n`` also writes the ``images_<ratio>_val/`` hold-out folders the dataset reads for validation.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))

from mine_b200.data.llff import resize_llff_images  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("root")
    ap.add_argument("--ratio", type=float, default=7.875)
    ap.add_argument("--val_every", type=int, default=0)
    a = ap.parse_args()
    print("wrote %d images" % resize_llff_images(a.root, a.ratio, a.val_every))
