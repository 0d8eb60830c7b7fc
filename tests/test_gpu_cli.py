"""The drop-in command lines on a GPU: `evaluate.py` with the reference's own flags against the golden checkpoints
and Set5 must log the PSNR the oracle pins (README: 37.15 dB for c-DCSCN x2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PKG

pytestmark = pytest.mark.gpu


def test_evaluate_cli_c_dcscn_set5(tmp_path):
    cmd = [sys.executable, os.path.join(PKG, "evaluate.py"), "--scale=2", "--layers=7", "--filters=32", "--min_filters=8",
           "--filters_decay_gamma=1.2", "--nin_filters=24", "--nin_filters2=8", "--reconstruct_layers=0", "--self_ensemble=1",
           "--batch_image_size=32", "--pixel_shuffler_filters=1", "--test_dataset=set5", "--save_results=false",
           "--data_dir=" + os.path.join(GOLDEN, "data"), "--checkpoint_dir=" + os.path.join(GOLDEN, "models"),
           "--log_filename=" + str(tmp_path / "log.txt"), "--tf_log_dir=" + str(tmp_path / "tf_log"),
           "--graph_dir=" + str(tmp_path / "graphs"), "--output_dir=" + str(tmp_path / "out")]
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    log = open(tmp_path / "log.txt").read()
    m = re.search(r"Model Average \[set5\] PSNR:([0-9.]+), SSIM:([0-9.nan]+), Time \(s\): ([0-9.]+)", log)
    assert m, log
    assert abs(float(m.group(1)) - 37.148) <= 0.01


def test_sr_cli_writes_outputs(tmp_path):
    img = os.path.join(GOLDEN, "data", "set5", "img_003.png")
    cmd = [sys.executable, os.path.join(PKG, "sr.py"), "--file=" + img, "--scale=2", "--layers=7", "--filters=32",
           "--min_filters=8", "--filters_decay_gamma=1.2", "--nin_filters=24", "--nin_filters2=8", "--reconstruct_layers=0",
           "--self_ensemble=1", "--pixel_shuffler_filters=1", "--checkpoint_dir=" + os.path.join(GOLDEN, "models"),
           "--log_filename=" + str(tmp_path / "log.txt"), "--tf_log_dir=" + str(tmp_path / "tf_log"),
           "--graph_dir=" + str(tmp_path / "graphs"), "--output_dir=" + str(tmp_path / "out")]
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = tmp_path / "out" / "dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32"
    names = sorted(os.listdir(out))
    assert names == ["img_003.png", "img_003_bicubic.png", "img_003_bicubic_y.png", "img_003_result.png", "img_003_result_y.png"]
    from PIL import Image
    assert Image.open(out / "img_003_result.png").size == (512, 512)


@pytest.mark.parametrize("build_batch", ["false", "true"])
def test_train_cli_runs_two_epochs_and_writes_a_resumable_checkpoint(tmp_path, build_batch):
    """`train.py` with the reference's flags on a toy schedule (Set5 as the training set, 2 steps per epoch, learning rate
    x0.01 per epoch so the run stops after two): the epoch / lr schedule, per-epoch evaluation, both data-set loaders and
    the checkpoint (trainables + Adam slots + beta powers) through the real command line."""
    from helper import tf_bundle
    ckpt = tmp_path / "ckpt"
    cmd = [sys.executable, os.path.join(PKG, "train.py"), "--scale=2", "--layers=7", "--filters=32", "--min_filters=8",
           "--filters_decay_gamma=1.2", "--nin_filters=24", "--nin_filters2=8", "--reconstruct_layers=0", "--self_ensemble=1",
           "--pixel_shuffler_filters=1", "--dataset=set5", "--test_dataset=set5", "--training_images=16", "--batch_num=8",
           "--batch_image_size=16", "--lr_decay_epoch=1", "--lr_decay=0.01", "--end_lr=1e-5", "--build_batch=" + build_batch,
           "--data_dir=" + os.path.join(GOLDEN, "data"), "--batch_dir=" + str(tmp_path / "batch"),
           "--checkpoint_dir=" + str(ckpt), "--log_filename=" + str(tmp_path / "log.txt"),
           "--tf_log_dir=" + str(tmp_path / "tf_log"), "--graph_dir=" + str(tmp_path / "graphs"),
           "--output_dir=" + str(tmp_path / "out")]
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    log = open(tmp_path / "log.txt").read()
    assert re.search(r"Model Average \[set5\] PSNR:([0-9.]+), SSIM:", log), log
    name = "dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32"
    rd = tf_bundle.BundleReader(str(ckpt / (name + ".ckpt")))
    keys = set(rd.keys())
    assert "CNN1/conv_W" in keys and "CNN1/conv_W/Adam" in keys and "CNN1/conv_W/Adam_1" in keys
    # 2 epochs x 2 steps = 4 Adam updates: beta1_power = 0.9^5
    assert float(rd.get_tensor("beta1_power")) == pytest.approx(0.9 ** 5, rel=1e-5)
    assert np.isfinite(rd.get_tensor("CNN1/conv_W")).all()
