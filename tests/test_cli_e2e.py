"""Command-line surface end to end on CPU: launcher -> 2-rank training (gloo) -> checkpoint -> resume -> video.

Mirrors the reference's only documented workflows (``README.md`` of the reference: ``sh start_training.sh ...`` and
``python3 visualizations/image_to_video.py ...``)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXTRA = {"data.training_set_path": "synthetic:8", "data.img_w": 64, "data.img_h": 64, "data.per_gpu_batch_size": 1,
         "mpi.num_bins_coarse": 4, "training.epochs": 2, "training.checkpoint_interval": 2, "training.log_interval": 1,
         "data.visible_point_count": 16}


def _launch(ws, port, max_steps):
    extra = dict(EXTRA, **{"training.max_steps": max_steps})
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    cmd = ["sh", os.path.join(REPO, "start_training.sh"), "MASTER_ADDR=127.0.0.1", "MASTER_PORT=%d" % port,
           "GPUS_PER_NODE=2", "WORKSPACE=%s" % ws, "DATASET=llff", "VERSION=v0", "EXTRA_CONFIG=%s" % json.dumps(extra)]
    return subprocess.run(cmd, env=env, cwd=REPO, capture_output=True, text=True, timeout=600)


@pytest.mark.timeout(1500)
def test_launcher_train_resume_and_video(tmp_path):
    ws = str(tmp_path / "ws")
    r = _launch(ws, 29621, 3)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = os.path.join(ws, "v0")
    names = os.listdir(out)
    assert {"params.yaml", "training.log", "checkpoint_latest.pth"} <= set(names)
    assert any(n.startswith("events.out.tfevents") for n in names)
    ck = torch.load(os.path.join(out, "checkpoint_latest.pth"), map_location="cpu", weights_only=False)
    assert {"backbone", "decoder", "optimizer"} <= set(ck)                       # upstream checkpoint layout
    assert any(k.startswith("module.encoder.layer1.0.conv1") for k in ck["backbone"])      # DDP-style upstream names
    step0 = int(ck["meta"]["global_step"])
    assert step0 == 3

    # same command again: picks up checkpoint_latest.pth and continues from step 3
    r = _launch(ws, 29622, 5)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    log = open(os.path.join(out, "training.log")).read()
    assert "Resumed at epoch 1 (batch 3), global_step 3" in log          # 4 batches per rank and epoch: mid-epoch restart
    ck = torch.load(os.path.join(out, "checkpoint_latest.pth"), map_location="cpu", weights_only=False)
    assert int(ck["meta"]["global_step"]) == 5 and int(ck["meta"]["epoch"]) == 2 and int(ck["meta"]["epoch_step"]) == 1

    # inference CLI on the trained checkpoint
    import cv2
    rng = np.random.default_rng(0)
    img_path = str(tmp_path / "photo.png")
    cv2.imwrite(img_path, (rng.random((80, 96, 3)) * 255).astype(np.uint8))
    vid_dir = str(tmp_path / "video")
    r = subprocess.run([sys.executable, os.path.join(REPO, "visualizations", "image_to_video.py"), "--checkpoint_path",
                        os.path.join(out, "checkpoint_latest.pth"), "--data_path", img_path, "--output_dir", vid_dir,
                        "--gpus", "cpu"], cwd=REPO, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    produced = os.listdir(vid_dir)
    assert any(n.endswith((".mp4", ".avi")) or os.path.isdir(os.path.join(vid_dir, n)) for n in produced), produced


@pytest.mark.timeout(900)
def test_train_cli_on_llff_layout_with_heldout_split(tmp_path):
    """COLMAP scene on disk -> resize helper (train / ``_val`` folders) -> ``train.py`` with periodic evaluation."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from test_api_and_data import _write_scene
    from mine_b200.data.llff import resize_llff_images
    root = str(tmp_path / "llff")
    _write_scene(root, n_views=8)
    assert resize_llff_images(root, 2.0, val_every=4) == 8
    assert sorted(os.listdir(os.path.join(root, "scene0", "images_2.0_val"))) == ["v1.png", "v5.png"]
    extra = {"data.training_set_path": root, "data.img_pre_downsample_ratio": 2.0, "data.img_w": 64, "data.img_h": 64,
             "data.per_gpu_batch_size": 2, "mpi.num_bins_coarse": 4, "training.epochs": 1, "training.eval_interval": 2,
             "training.log_interval": 1, "data.visible_point_count": 16}
    cmd = [sys.executable, os.path.join(REPO, "train.py"), "--config_path", os.path.join(REPO, "configs/params_llff.yaml"),
           "--workspace", str(tmp_path / "ws"), "--version", "v0"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    r = subprocess.run(cmd + ["--extra_config", json.dumps(extra)], cwd=REPO, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = str(tmp_path / "ws" / "v0")
    assert os.path.exists(os.path.join(out, "checkpoint_000000000002.pth"))       # written after the evaluation pass
    log = open(os.path.join(out, "training.log")).read()
    assert "number of images: 6" in log and "number of images: 2" in log and "val_psnr_tgt" in log
    # stand-alone evaluation of the checkpoint on the held-out split
    r = subprocess.run([sys.executable, os.path.join(REPO, "evaluate.py"), "--checkpoint_path",
                        os.path.join(out, "checkpoint_latest.pth"), "--device", "cpu", "--output", str(tmp_path / "m.json")],
                       cwd=REPO, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(open(tmp_path / "m.json").read())
    assert res["num_images"] == 2 and 5.0 < res["metrics"]["psnr_tgt"] < 60.0
    assert set(res["metrics"]) >= {"loss_rgb_tgt", "loss_ssim_tgt", "lpips_tgt", "psnr_tgt"}
    # the default ratio points at folders that do not exist here: fail loudly instead of training on nothing
    extra["data.img_pre_downsample_ratio"] = 7.875
    r = subprocess.run(cmd + ["--extra_config", json.dumps(extra)], cwd=REPO, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode != 0 and "no training images under" in (r.stdout + r.stderr)


@pytest.mark.timeout(900)
def test_sigterm_saves_resumable_state_and_exits_cleanly(tmp_path):
    """Preemption: SIGTERM -> the loop stops at the next agreed step boundary, writes checkpoint_latest.pth, exits 0;
    the same command then resumes from that step."""
    import signal
    import time
    extra = dict(EXTRA, **{"data.training_set_path": "synthetic:64", "training.epochs": 50, "training.max_steps": 0})
    cmd = [sys.executable, os.path.join(REPO, "train.py"), "--config_path", os.path.join(REPO, "configs/params_llff.yaml"),
           "--workspace", str(tmp_path / "ws"), "--version", "v0", "--extra_config", json.dumps(extra)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
    log_path = str(tmp_path / "ws" / "v0" / "training.log")
    proc = subprocess.Popen(cmd, cwd=REPO, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        deadline = time.time() + 300
        while time.time() < deadline:
            if os.path.exists(log_path) and "global_step = 2 " in open(log_path).read():
                break
            assert proc.poll() is None, proc.stdout.read()[-3000:]
            time.sleep(0.5)
        else:
            pytest.fail("training did not start")
        proc.send_signal(signal.SIGTERM)
        out, _ = proc.communicate(timeout=120)
    finally:
        if proc.poll() is None:
            proc.kill()
    assert proc.returncode == 0, out[-3000:]
    log = open(log_path).read()
    assert "Stop requested (signal)" in log
    ck = torch.load(str(tmp_path / "ws" / "v0" / "checkpoint_latest.pth"), map_location="cpu", weights_only=False)
    step = int(ck["meta"]["global_step"])
    assert step >= 2 and "optimizer" in ck
    extra["training.max_steps"] = step + 1
    cmd[-1] = json.dumps(extra)
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "global_step %d" % step in open(log_path).read().split("Stop requested (signal)")[-1]
